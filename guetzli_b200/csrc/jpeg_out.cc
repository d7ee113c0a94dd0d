// See jpeg_out.h.
#include "jpeg_out.h"

#include <stdlib.h>

#include <algorithm>

#include "tables.h"

namespace gb200 {

namespace {

inline int floor_log2_nz(uint32_t n) { return 31 ^ __builtin_clz(n); }
inline int floor_log2(uint32_t n) { return n == 0 ? -1 : floor_log2_nz(n); }

}  // namespace

// Length-limited code lengths as g/entropy_encode.cc:73 builds them: Huffman tree over the
// counts raised to a floor; whenever a leaf ends up deeper than `limit` the floor doubles and
// the tree is rebuilt.  The reference re-sorts the leaves for every floor.  Raising the floor
// only changes the order among the leaves at or below it (they become equal and then go by
// larger symbol first), and those are a prefix of the list sorted once by (count, larger
// symbol first) -- so each further attempt is a linear pass.  Same trees, same depths.
void huffman_code_lengths(const uint32_t* counts, int n, int limit, uint8_t* depth) {
  const int kMaxN = 512;
  if (n > kMaxN) abort();  // 257 everywhere (SymbolHistogram::kSize)
  uint64_t keys[kMaxN];
  // nodes: [0, leaves) the leaves in merge order, [leaves] a sentinel, (leaves, 2 leaves) the parents
  uint32_t count[2 * kMaxN + 2];
  int16_t left[2 * kMaxN + 2], right[2 * kMaxN + 2];  // leaves: right = symbol
  uint8_t level[2 * kMaxN + 2];
  int leaves = 0;
  for (int i = n - 1; i >= 0; --i)
    if (counts[i]) keys[leaves++] = (static_cast<uint64_t>(counts[i]) << 32) | static_cast<uint32_t>(0x7fffffff - i);
  if (leaves == 0) return;
  if (leaves == 1) {
    depth[0x7fffffff - static_cast<int>(keys[0] & 0xffffffffu)] = 1;
    return;
  }
  // least frequent first; equal counts: larger symbol first
  std::sort(keys, keys + leaves);
  for (int i = 0; i < leaves; ++i) {
    count[i] = static_cast<uint32_t>(keys[i] >> 32);
    right[i] = static_cast<int16_t>(0x7fffffff - static_cast<int>(keys[i] & 0xffffffffu));
  }
  int below = 0;  // leaves whose count is <= the floor: keys[0, below)
  const uint32_t kSentinel = ~static_cast<uint32_t>(0);
  for (uint32_t floor_count = 1;; floor_count *= 2) {
    const int below_before = below;
    while (below < leaves && static_cast<uint32_t>(keys[below] >> 32) <= floor_count) ++below;
    if (floor_count > 1 && below > 0) {
      // the raised leaves, larger symbol first (the list beyond them keeps its order)
      if (below != below_before || floor_count == 2) {
        uint8_t raised[kMaxN / 8] = {0};
        for (int i = 0; i < below; ++i) {
          const int sym = 0x7fffffff - static_cast<int>(keys[i] & 0xffffffffu);
          raised[sym >> 3] |= static_cast<uint8_t>(1u << (sym & 7));
        }
        int m = 0;
        for (int sym = n - 1; sym >= 0; --sym)
          if (raised[sym >> 3] & (1u << (sym & 7))) right[m++] = static_cast<int16_t>(sym);
      }
      for (int i = 0; i < below; ++i) count[i] = floor_count;
    }
    // two-queue merge: leaves in [0,leaves), parents appended from leaves+1
    count[leaves] = kSentinel;
    count[leaves + 1] = kSentinel;
    int i = 0, j = leaves + 1;
    for (int k = leaves - 1; k != 0; --k) {
      int l, r;
      if (count[i] <= count[j]) l = i++; else l = j++;
      if (count[i] <= count[j]) r = i++; else r = j++;
      const int parent = 2 * leaves - k;
      count[parent] = count[l] + count[r];
      left[parent] = static_cast<int16_t>(l);
      right[parent] = static_cast<int16_t>(r);
      count[parent + 1] = kSentinel;
    }
    // a parent has a larger index than its children: levels top-down in one sweep
    const int root = 2 * leaves - 1;
    level[root] = 0;
    for (int p = root; p > leaves; --p) {
      const uint8_t l = static_cast<uint8_t>(level[p] + 1);
      level[left[p]] = l;
      level[right[p]] = l;
    }
    int deepest = 0;
    for (int k = 0; k < leaves; ++k) deepest = level[k] > deepest ? level[k] : deepest;
    if (deepest <= limit) {
      for (int k = 0; k < leaves; ++k) depth[right[k]] = level[k];
      return;
    }
  }
}

size_t histogram_header_bits(const SymbolHistogram& h) {
  size_t bits = 17 * 8;
  for (int i = 0; i + 1 < SymbolHistogram::kSize; ++i)
    if (h.counts[i] > 0) bits += 8;
  return bits;
}

size_t histogram_data_bits(const SymbolHistogram& h, const uint8_t* depth) {
  size_t bits = 0;
  for (int i = 0; i + 1 < SymbolHistogram::kSize; ++i) bits += (h.counts[i] / 2) * (depth[i] + (i & 0xf));
  bits += (bits * 3 + 512) >> 10;  // 0xff stuffing estimate
  return bits;
}

size_t cluster_histograms(SymbolHistogram* h, size_t* num, int* index, uint8_t* depth) {
  const int K = SymbolHistogram::kSize;
  memset(depth, 0, *num * K);
  size_t costs[4];
  for (size_t i = 0; i < *num; ++i) {
    index[i] = static_cast<int>(i);
    huffman_code_lengths(h[i].counts, K, 16, &depth[i * K]);
    costs[i] = histogram_header_bits(h[i]) + histogram_data_bits(h[i], &depth[i * K]);
  }
  const size_t orig_num = *num;
  while (*num > 1) {
    const size_t last = *num - 1, prev = *num - 2;
    SymbolHistogram both(h[last]);
    both.merge(h[prev]);
    uint8_t depth_both[SymbolHistogram::kSize] = {0};
    huffman_code_lengths(both.counts, K, 16, depth_both);
    const size_t cost_both = histogram_header_bits(both) + histogram_data_bits(both, depth_both);
    if (cost_both < costs[last] + costs[prev]) {
      h[prev] = both;
      h[last] = SymbolHistogram();
      costs[prev] = cost_both;
      memcpy(&depth[prev * K], depth_both, sizeof(depth_both));
      for (size_t i = 0; i < orig_num; ++i)
        if (index[i] == static_cast<int>(last)) index[i] = static_cast<int>(prev);
      --(*num);
    } else {
      break;
    }
  }
  size_t total = 0;
  for (size_t i = 0; i < *num; ++i) total += costs[i];
  return (total + 7) / 8;
}

int num_output_components(const CoeffImage& img) {
  if (img.as_encoded || img.as_read) return 3;
  const size_t n = static_cast<size_t>(img.nblocks) * 64;
  const int16_t* p = img.coeffs + n;
  for (size_t i = 0; i < 2 * n; ++i)
    if (p[i] != 0) return 3;
  return 1;
}

void ac_symbols_of_block(const int16_t* dq, const int* q, int weight, SymbolHistogram* h) {
  const int* zz = zigzag_to_natural();
  int run = 0;
  for (int k = 1; k < 64; ++k) {
    const int nat = zz[k];
    const int16_t coeff = dq[nat];
    if (coeff == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      h->add(0xf0, weight);
      run -= 16;
    }
    const int nbits = floor_log2_nz(abs(coeff / q[nat])) + 1;
    h->add((run << 4) + nbits, weight);
    run = 0;
  }
  if (run > 0) h->add(0, weight);
}

// Same symbols as ac_symbols_of_block, restricted to zig-zag positions (a, b] where a is
// a nonzero coefficient (or 0 = the DC slot) and b the next position to stop after
// (a nonzero coefficient, or 64 = run to the end incl. the end-of-block symbol).
// The run counter restarts after every nonzero coefficient, so the symbols of a
// range depend only on the coefficients inside it.
void ac_symbols_of_range(const int16_t* dq, const int* q, int a, int b, int weight, SymbolHistogram* h) {
  const int* zz = zigzag_to_natural();
  int run = 0;
  const int last = b < 64 ? b : 63;
  for (int k = a + 1; k <= last; ++k) {
    const int nat = zz[k];
    const int16_t coeff = dq[nat];
    if (coeff == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      h->add(0xf0, weight);
      run -= 16;
    }
    const int nbits = floor_log2_nz(abs(coeff / q[nat])) + 1;
    h->add((run << 4) + nbits, weight);
    run = 0;
  }
  if (b >= 64 && run > 0) h->add(0, weight);
}

void build_ac_histograms(const CoeffImage& img, SymbolHistogram* h3) {
  const int ncomp = num_output_components(img);
  for (int c = 0; c < ncomp; ++c)
    for (int b = 0; b < img.nblocks; ++b) ac_symbols_of_block(img.block(c, b), img.q[c], 1, &h3[c]);
}

namespace {
void build_dc_histograms(const CoeffImage& img, int ncomp, SymbolHistogram* h) {
  for (int c = 0; c < ncomp; ++c) {
    int last = 0;
    const int q0 = img.q[c][0];
    for (int b = 0; b < img.nblocks; ++b) {
      const int dc = img.block(c, b)[0] / q0;
      const int diff = abs(static_cast<int16_t>(dc) - static_cast<int16_t>(last));
      h[c].add(floor_log2(diff) + 1);
      last = dc;
    }
  }
}

// Distinct quant tables in component order (g/jpeg_data.cc:70).
struct QuantSet {
  int num;
  int table[4][64];
  int precision[4];
  int index[4];     // Tq written into DQT
  int comp_idx[3];  // table position per component
  int comp_id[3];   // component id written into SOF / SOS
};
QuantSet dedup_quant(const CoeffImage& img, int ncomp) {
  QuantSet qs;
  qs.num = 0;
  for (int c = 0; c < 3; ++c) qs.comp_id[c] = c;
  if (img.as_read) {
    const JpegFileLayout& f = *img.as_read;
    qs.num = f.num_tables;
    for (int i = 0; i < f.num_tables; ++i) {
      memcpy(qs.table[i], f.table[i], sizeof(qs.table[i]));
      qs.precision[i] = f.precision[i];
      qs.index[i] = f.index[i];
    }
    for (int c = 0; c < 3; ++c) {
      qs.comp_idx[c] = f.comp_table[c];
      qs.comp_id[c] = f.comp_id[c];
    }
    return qs;
  }
  if (img.as_encoded) {
    qs.num = 3;
    for (int c = 0; c < 3; ++c) {
      memcpy(qs.table[c], img.q[c], sizeof(qs.table[c]));
      qs.precision[c] = 0;
      qs.index[c] = 0;
      qs.comp_idx[c] = c;
    }
    return qs;
  }
  for (int c = 0; c < ncomp; ++c) {
    int found = -1;
    for (int j = 0; j < qs.num; ++j)
      if (memcmp(img.q[c], qs.table[j], sizeof(qs.table[j])) == 0) {
        found = j;
        break;
      }
    if (found < 0) {
      memcpy(qs.table[qs.num], img.q[c], sizeof(qs.table[0]));
      qs.precision[qs.num] = 0;
      for (int k = 0; k < 64; ++k)
        if (img.q[c][k] > 0xff) qs.precision[qs.num] = 1;
      qs.index[qs.num] = qs.num;
      found = qs.num++;
    }
    qs.comp_idx[c] = found;
  }
  return qs;
}

struct CodeTable {
  uint8_t depth[256];
  int code[256];
};

// Canonical JPEG code from code lengths; the phantom symbol 256 sorts last
// within the longest length and is dropped, which reserves the all-ones code
// (g/jpeg_data_writer.cc:130-183,401-427).
void canonical_code(const uint8_t* depth, int* counts /*[17]*/, int* values /*[257]*/, CodeTable* table) {
  const int K = SymbolHistogram::kSize;
  for (int i = 0; i <= 16; ++i) counts[i] = 0;
  for (int i = 0; i < K; ++i)
    if (depth[i] > 0) ++counts[depth[i]];
  int offset[17] = {0};
  for (int i = 1; i <= 16; ++i) offset[i] = offset[i - 1] + counts[i - 1];
  for (int i = 0; i < K; ++i)
    if (depth[i] > 0) values[offset[depth[i]]++] = i;
  for (int i = 0; i < 256; ++i) table->depth[i] = 255;
  int total = 0;
  for (int l = 1; l <= 16; ++l) total += counts[l];
  if (total == 0) return;
  int code = 0, p = 0;
  for (int l = 1; l <= 16; ++l) {
    for (int i = 0; i < counts[l]; ++i, ++p) {
      if (p < total - 1) {  // all but the phantom
        table->depth[values[p]] = static_cast<uint8_t>(l);
        table->code[values[p]] = code;
      }
      ++code;
    }
    code <<= 1;
  }
}

struct BitSink {
  std::string* out;
  uint64_t acc;
  int nbits;  // bits currently held in acc (low end)
  explicit BitSink(std::string* o) : out(o), acc(0), nbits(0) {}
  inline void put_byte(int b) {
    out->push_back(static_cast<char>(b));
    if (b == 0xff) out->push_back(0);
  }
  inline void put(int n, uint32_t bits) {
    acc = (acc << n) | bits;
    nbits += n;
    while (nbits >= 8) {
      nbits -= 8;
      put_byte(static_cast<int>((acc >> nbits) & 0xff));
    }
  }
  void finish() {
    if (nbits > 0) {
      const int pad = 8 - nbits;
      put_byte(static_cast<int>(((acc << pad) | ((1u << pad) - 1)) & 0xff));
      nbits = 0;
    }
  }
};

}  // namespace

size_t estimate_dc_bytes(const CoeffImage& img) {
  // EstimateDCSize builds ncomp histograms (of the saved JPEG) and clusters them.
  const int ncomp = num_output_components(img);
  SymbolHistogram h[3];
  build_dc_histograms(img, ncomp, h);
  size_t num = ncomp;
  int index[4];
  uint8_t depth[3 * SymbolHistogram::kSize];
  return cluster_histograms(h, &num, index, depth);
}

size_t estimate_dc_bytes_of(SymbolHistogram* dc_h, int ncomp) {
  size_t num = ncomp;
  int index[4];
  uint8_t depth[3 * SymbolHistogram::kSize];
  return cluster_histograms(dc_h, &num, index, depth);
}

size_t jpeg_header_bytes(const CoeffImage& img) { return jpeg_header_bytes(img, num_output_components(img)); }

size_t jpeg_header_bytes(const CoeffImage& img, int ncomp) {
  const QuantSet qs = dedup_quant(img, ncomp);
  size_t n = 2;  // SOI
  if (img.meta == nullptr || img.meta->strip) {
    n += 18;  // APP0
  } else {
    for (const std::string& a : img.meta->app_data) n += 1 + a.size();
    for (const std::string& c : img.meta->com_data) n += 2 + c.size();
  }
  if (img.meta) n += img.meta->tail_data.size();  // counted even when stripped (g/jpeg_data_writer.cc:291)
  n += 4;  // DQT marker + length
  for (int i = 0; i < qs.num; ++i) n += 1 + (qs.precision[i] ? 2 : 1) * 64;
  n += 10 + 3 * ncomp;  // SOF
  n += 4;               // DHT marker + length
  n += 8 + 2 * ncomp;   // SOS
  n += 2;               // EOI
  return n;
}

size_t compute_entropy_codes(const SymbolHistogram* h3, uint8_t* depths, int ncomp) {
  const int K = SymbolHistogram::kSize;
  SymbolHistogram clustered[3] = {h3[0], h3[1], h3[2]};
  size_t num = static_cast<size_t>(ncomp);
  int index[4];
  uint8_t cdepth[3 * SymbolHistogram::kSize];
  cluster_histograms(clustered, &num, index, cdepth);
  for (int i = 0; i < ncomp; ++i) memcpy(&depths[i * K], &cdepth[index[i] * K], K);
  size_t bytes = 0;
  for (size_t i = 0; i < num; ++i) bytes += histogram_header_bits(clustered[i]) / 8;
  return bytes;
}

size_t entropy_coded_bytes(const SymbolHistogram* h3, const uint8_t* depths) {
  size_t bits = 0;
  for (int i = 0; i < 3; ++i) bits += histogram_data_bits(h3[i], &depths[i * SymbolHistogram::kSize]);
  return (bits + 7) / 8;
}

JpegPlan plan_jpeg(const CoeffImage& img, int ncomp, SymbolHistogram* dc_h, SymbolHistogram* ac_h) {
  const int K = SymbolHistogram::kSize;
  JpegPlan plan;
  plan.ncomp = ncomp;
  memset(plan.depth, 0, sizeof(plan.depth));
  memset(plan.code, 0, sizeof(plan.code));
  const QuantSet qs = dedup_quant(img, ncomp);
  const int* zz = zigzag_to_natural();
  std::string& out = plan.prefix;
  auto byte = [&out](int b) { out.push_back(static_cast<char>(b)); };

  // SOI + JFIF APP0, or the input's APPn / COM segments (g/jpeg_data_writer.cc:52-74)
  byte(0xff); byte(0xd8);
  if (img.meta == nullptr || img.meta->strip) {
    static const uint8_t kApp0[] = {0xff, 0xe0, 0x00, 0x10, 0x4a, 0x46, 0x49, 0x46, 0x00,
                                    0x01, 0x01, 0x00, 0x00, 0x01, 0x00, 0x01, 0x00, 0x00};
    out.append(reinterpret_cast<const char*>(kApp0), sizeof(kApp0));
  } else {
    for (const std::string& a : img.meta->app_data) {
      byte(0xff);
      out.append(a);
    }
    for (const std::string& c : img.meta->com_data) {
      byte(0xff); byte(0xfe);
      out.append(c);
    }
  }
  plan.trailer = jpeg_trailer(img);
  // DQT
  {
    int len = 2;
    for (int i = 0; i < qs.num; ++i) len += 1 + (qs.precision[i] ? 2 : 1) * 64;
    byte(0xff); byte(0xdb); byte(len >> 8); byte(len & 0xff);
    for (int i = 0; i < qs.num; ++i) {
      byte((qs.precision[i] << 4) + qs.index[i]);
      for (int k = 0; k < 64; ++k) {
        const int v = qs.table[i][zz[k]];
        if (qs.precision[i]) byte(v >> 8);
        byte(v & 0xff);
      }
    }
  }
  // SOF1 (extended sequential, 0xc1)
  {
    const int len = 8 + 3 * ncomp;
    byte(0xff); byte(0xc1); byte(len >> 8); byte(len & 0xff);
    byte(8);
    byte(img.h >> 8); byte(img.h & 0xff);
    byte(img.w >> 8); byte(img.w & 0xff);
    byte(ncomp);
    for (int c = 0; c < ncomp; ++c) {
      byte(qs.comp_id[c]);
      byte(0x11);
      byte(qs.index[qs.comp_idx[c]]);
    }
  }
  // Huffman codes: cluster DC then AC histograms
  size_t num_dc = ncomp, num_ac = ncomp;
  int dc_index[4], ac_index[4];
  std::vector<uint8_t> dc_depth(3 * K), ac_depth(3 * K);
  cluster_histograms(dc_h, &num_dc, dc_index, dc_depth.data());
  cluster_histograms(ac_h, &num_ac, ac_index, ac_depth.data());
  {
    int total_symbols = 0;
    for (size_t i = 0; i < num_dc; ++i) total_symbols += dc_h[i].num_symbols();
    for (size_t i = 0; i < num_ac; ++i) total_symbols += ac_h[i].num_symbols();
    const int num_histo = static_cast<int>(num_dc + num_ac);
    const int dht_len = 2 + num_histo * 17 + total_symbols;
    byte(0xff); byte(0xc4); byte(dht_len >> 8); byte(dht_len & 0xff);
    for (int i = 0; i < num_histo; ++i) {
      const bool is_dc = i < static_cast<int>(num_dc);
      const int idx = is_dc ? i : i - static_cast<int>(num_dc);
      int counts[17], values[SymbolHistogram::kSize] = {0};
      CodeTable table;
      canonical_code(is_dc ? &dc_depth[idx * K] : &ac_depth[idx * K], counts, values, &table);
      for (int c = 0; c < ncomp; ++c) {
        const bool mine = is_dc ? dc_index[c] == idx : ac_index[c] == idx;
        if (!mine) continue;
        const int slot = is_dc ? c : 3 + c;
        for (int s = 0; s < 256; ++s) {
          plan.depth[slot][s] = table.depth[s] == 255 ? 0 : table.depth[s];
          plan.code[slot][s] = table.depth[s] == 255 ? 0 : static_cast<uint16_t>(table.code[s]);
        }
      }
      int max_len = 16;
      while (max_len > 0 && counts[max_len] == 0) --max_len;
      --counts[max_len];  // drop the phantom symbol
      int total = 0;
      for (int j = 0; j <= max_len; ++j) total += counts[j];
      byte(is_dc ? i : idx + 0x10);
      for (int j = 1; j <= 16; ++j) byte(counts[j]);
      for (int j = 0; j < total; ++j) byte(values[j]);
    }
  }
  // SOS
  {
    const int len = 6 + 2 * ncomp;
    byte(0xff); byte(0xda); byte(len >> 8); byte(len & 0xff);
    byte(ncomp);
    for (int c = 0; c < ncomp; ++c) {
      byte(qs.comp_id[c]);
      byte((dc_index[c] << 4) | ac_index[c]);
    }
    byte(0); byte(63); byte(0);
  }
  return plan;
}

void host_symbol_histograms(const CoeffImage& img, int ncomp, SymbolHistogram* dc_h, SymbolHistogram* ac_h) {
  build_dc_histograms(img, ncomp, dc_h);
  for (int c = 0; c < ncomp; ++c)
    for (int b = 0; b < img.nblocks; ++b) ac_symbols_of_block(img.block(c, b), img.q[c], 1, &ac_h[c]);
}

std::string write_jpeg(const CoeffImage& img) {
  const int ncomp = num_output_components(img);
  const int* zz = zigzag_to_natural();
  SymbolHistogram dc_h[3], ac_h[3];
  host_symbol_histograms(img, ncomp, dc_h, ac_h);
  JpegPlan plan = plan_jpeg(img, ncomp, dc_h, ac_h);
  std::string out;
  out.reserve(static_cast<size_t>(img.nblocks) * 48 + 1024);
  out = plan.prefix;
  // entropy-coded scan, one block of each component per MCU (444)
  {
    BitSink bw(&out);
    int last_dc[3] = {0, 0, 0};
    for (int b = 0; b < img.nblocks; ++b) {
      for (int c = 0; c < ncomp; ++c) {
        const int16_t* dq = img.block(c, b);
        const int* q = img.q[c];
        const uint8_t* dcd = plan.depth[c];
        const uint16_t* dcc = plan.code[c];
        const uint8_t* acd = plan.depth[3 + c];
        const uint16_t* acc = plan.code[3 + c];
        // DC difference (g/jpeg_data_writer.cc:460-474), int16 arithmetic
        const int16_t dc = static_cast<int16_t>(dq[0] / q[0]);
        int16_t diff = static_cast<int16_t>(dc - last_dc[c]);
        last_dc[c] = dc;
        int16_t bits = diff;
        if (diff < 0) {
          diff = static_cast<int16_t>(-diff);
          --bits;
        }
        int nbits = floor_log2(static_cast<uint32_t>(static_cast<int>(diff))) + 1;
        bw.put(dcd[nbits], dcc[nbits]);
        if (nbits > 0) bw.put(nbits, bits & ((1 << nbits) - 1));
        int run = 0;
        for (int k = 1; k < 64; ++k) {
          const int nat = zz[k];
          int v = dq[nat];
          if (v == 0) {
            ++run;
            continue;
          }
          v /= q[nat];
          int mag = v, low = v;
          if (v < 0) {
            mag = -v;
            low = ~mag;
          }
          while (run > 15) {
            bw.put(acd[0xf0], acc[0xf0]);
            run -= 16;
          }
          nbits = floor_log2_nz(mag) + 1;
          const int symbol = (run << 4) + nbits;
          bw.put(acd[symbol], acc[symbol]);
          bw.put(nbits, low & ((1 << nbits) - 1));
          run = 0;
        }
        if (run > 0) bw.put(acd[0], acc[0]);
      }
    }
    bw.finish();
  }
  out.append(plan.trailer);
  return out;
}

std::string jpeg_trailer(const CoeffImage& img) {
  std::string t;
  t.push_back(static_cast<char>(0xff));
  t.push_back(static_cast<char>(0xd9));
  if (img.meta && !img.meta->strip) t.append(img.meta->tail_data);
  return t;
}

}  // namespace gb200
