// Device-resident half of the selection walk (a16, g/processor.cc:611-779).
//
// An iteration of SelectFrequencyMasking consumes the first N entries of a global
// order of (block, key) pairs; consuming an entry of block b flips that block's next
// candidate coefficient.  The walk only LOOKS at its state (entropy-code refresh,
// stop test) from entry i0 = the first multiple of 10 with i0 + 9 >= min_coeffs_to_change
// on; before that it is a pure fold over a SET of entries, and the result of a fold
// over a set does not depend on the order inside the set.  So:
//
//   bulk    entries [0, i0) of the sorted order are applied here, on the device, one
//           thread per touched block (a block's own entries are consumed in its
//           candidate order, exactly like the sequential walk would), producing the
//           AC symbol histogram delta, the chroma-nonzero delta and an undo log;
//   window  the entries from i0 on go to the host, which runs the reference's sequential
//           loop (Search::walk) on them with the state of just the blocks involved.
//
// All integer work; every function mirrors its host counterpart in search.cc / jpeg_out.cc
// (plan_edit, ac_symbols_of_range) and is run by the CPU port through the same functors.
#pragma once
#include "hd.h"
#include "jpeg_dev.h"
#include "jpeg_math.h"
#include "kernels.h"

namespace gb200 {

// entries / blocks of the order of one direction, summed by `lanes` lanes
struct WalkStatsPartial {
  const int* last_index;
  const int* z_cnt;
  const float* weight;
  int direction, nblocks, lanes;
  unsigned long long* out;  // [lanes][2]: entries, blocks with entries
  GB_HD void operator()(int i) const {
    unsigned long long n = 0, c = 0;
    for (int b = i; b < nblocks; b += lanes) {
      if (weight[b] == 0) continue;
      const int li = last_index[b], nc = z_cnt[b];
      const int m = direction > 0 ? (li < nc ? nc - li : 0) : (li > 0 ? li : 0);
      n += static_cast<unsigned long long>(m);
      c += m > 0 ? 1u : 0u;
    }
    out[2 * i] = n;
    out[2 * i + 1] = c;
  }
};

// ---------------------------------------------------------------------------
// Two-rank radix select over the order keys (same two-level scheme as OrderSelectState in
// kernels.h, both ranks from the same two histogram passes).  lo = the rank a little before
// the end of the bulk, hi = the rank the host window may reach.  After it:
//   keys in 22-bit bins below lo22           -> in the bulk for sure: counted per block directly
//   keys in bins lo22 .. hi22 (the "middle") -> compacted, sorted; the first entries complete the
//                                               bulk, the rest is the window
//   keys above                               -> not needed this iteration
struct Select2State {
  unsigned int want_lo, want_hi;
  unsigned int bin0_lo, below0_lo, bin0_hi, below0_hi, total;
  unsigned int lo22, before_lo;  // 22-bit bin holding rank want_lo; entries in bins below it
  unsigned int hi22, kept_hi;    // 22-bit bin holding rank want_hi; entries in bins up to and including it
  unsigned int mid_count;        // cursor of the middle list
};

// first bin whose cumulative count reaches `want` (>= 1); none: the last bin.  Serial form.
GB_HD void rank_bin_serial(const unsigned int* hist, int nbins, unsigned int want, unsigned int* bin,
                           unsigned int* before, unsigned int* total) {
  unsigned int cum = 0, b = static_cast<unsigned int>(nbins - 1), at = 0;
  bool found = false;
  for (int i = 0; i < nbins; ++i) {
    if (!found && cum + hist[i] >= want) {
      b = static_cast<unsigned int>(i);
      at = cum;
      found = true;
    }
    cum += hist[i];
  }
  if (!found) at = cum - hist[nbins - 1];
  *bin = b;
  *before = at;
  *total = cum;
}

struct Select2Level0 {  // one invocation
  const unsigned int* hist;  // [2048]: bits 31..21 of every key
  Select2State* st;
  GB_HD void operator()(int) const {
    unsigned int total;
    rank_bin_serial(hist, kOrderBins, st->want_lo, &st->bin0_lo, &st->below0_lo, &total);
    rank_bin_serial(hist, kOrderBins, st->want_hi, &st->bin0_hi, &st->below0_hi, &total);
    st->total = total;
  }
};

struct Select2Hist1 {  // over the entries: bits 20..10 of the keys in the two level-0 bins
  OrderKeyCommon c;
  unsigned int* hist;  // [2][2048]
  const Select2State* st;
  GB_HD void operator()(int entry) const {
    float v;
    int b;
    if (!c.key(entry, &b, &v)) return;
    const unsigned int u = hd_float_sortable(v);
    const unsigned int top = u >> 21, mid = (u >> 10) & 0x7ffu;
    if (top == st->bin0_lo) hd_atomic_add(&hist[mid], 1u);
    if (top == st->bin0_hi) hd_atomic_add(&hist[kOrderBins + mid], 1u);
  }
};

struct Select2Level1 {  // one invocation
  const unsigned int* hist;  // [2][2048]
  Select2State* st;
  GB_HD void operator()(int) const {
    unsigned int bin, before, total;
    rank_bin_serial(hist, kOrderBins, st->want_lo > st->below0_lo ? st->want_lo - st->below0_lo : 1u, &bin, &before,
                    &total);
    st->lo22 = (st->bin0_lo << 11) | bin;
    st->before_lo = st->below0_lo + before;
    rank_bin_serial(hist + kOrderBins, kOrderBins, st->want_hi > st->below0_hi ? st->want_hi - st->below0_hi : 1u, &bin,
                    &before, &total);
    st->hi22 = (st->bin0_hi << 11) | bin;
    st->kept_hi = st->below0_hi + before + hist[kOrderBins + bin];
    st->mid_count = 0;
  }
};

// nonzero coefficients of a coefficient range, summed by `lanes` lanes
struct CountNonzeroPartial {
  const int16_t* coeffs;
  size_t n;
  int lanes;
  unsigned long long* out;
  GB_HD void operator()(int i) const {
    unsigned long long c = 0;
    for (size_t k = static_cast<size_t>(i); k < n; k += static_cast<size_t>(lanes)) c += coeffs[k] != 0 ? 1u : 0u;
    out[i] = c;
  }
};

// compact list of all candidates: entry offset[b] + i  ->  (block b, slot i), i < cnt[b]
struct FillEntries {
  const int* cnt;
  const unsigned int* offset;
  int* entry_block;
  uint8_t* entry_slot;
  GB_HD void operator()(int b) const {
    const int n = cnt[b];
    const unsigned int o = offset[b];
    for (int i = 0; i < n; ++i) {
      entry_block[o + i] = b;
      entry_slot[o + i] = static_cast<uint8_t>(i);
    }
  }
};

// number of order keys below a limit (the partition_point of g/processor.cc:690-698)
struct CountKeysBelow {
  OrderKeyCommon c;
  float limit;
  unsigned int* out;
  GB_HD void operator()(int entry) const {
    float v;
    int b;
    if (!c.key(entry, &b, &v)) return;
    if (v < limit) hd_atomic_add(out, 1u);
  }
};

// classification pass: bulk-for-sure entries are counted per block (BulkCount), the middle is compacted
struct Select2Split {
  OrderKeyCommon c;
  Select2State* st;
  unsigned int* cnt;       // [nblocks]
  int* touched;
  unsigned int* n_touched;
  float* mid_val;
  int* mid_block;
  unsigned int mid_cap;
  GB_HD void operator()(int entry) const {
    float v;
    int b;
    if (!c.key(entry, &b, &v)) return;
    const unsigned int u22 = hd_float_sortable(v) >> 10;
    if (u22 < st->lo22) {
      if (hd_atomic_add(&cnt[b], 1u) == 0u) touched[hd_atomic_add(n_touched, 1u)] = b;
    } else if (u22 <= st->hi22) {
      const unsigned int at = hd_atomic_add(&st->mid_count, 1u);
      if (at < mid_cap) {
        mid_val[at] = v;
        mid_block[at] = b;
      }
    }
  }
};

struct ResetCounts {
  const int* touched;
  unsigned int* cnt;
  GB_HD void operator()(int j) const { cnt[touched[j]] = 0u; }
};

// how many of the first `n` sorted entries belong to each block; first toucher lists the block
struct BulkCount {
  const int* sel_block;
  unsigned int* cnt;       // [nblocks], zero on entry
  int* touched;            // [nblocks]
  unsigned int* n_touched;
  GB_HD void operator()(int i) const {
    const int b = sel_block[i];
    if (hd_atomic_add(&cnt[b], 1u) == 0u) touched[hd_atomic_add(n_touched, 1u)] = b;
  }
};

struct WalkState {
  const int16_t* orig;     // [3][nblocks][64]
  int16_t* cand;
  const int* q;            // [192]
  const int* zz2nat;       // [64]
  const int* nat2zz;       // [64]
  const uint8_t* z_idx;    // [nblocks][192]
  int* last_index;         // [nblocks]
  int nblocks;
};

// AC symbols of zig-zag positions (a, b] of a block, weight +-1 into hist[256] (ac_symbols_of_range)
GB_HD void dev_ac_symbols_of_range(const int16_t* dq, const int* q, const int* zz, int a, int b, unsigned int weight,
                                   unsigned int* hist) {
  int run = 0;
  const int last = b < 64 ? b : 63;
  for (int k = a + 1; k <= last; ++k) {
    const int nat = zz[k];
    const int coeff = dq[nat];
    if (coeff == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      hd_atomic_add(&hist[0xf0], weight);
      run -= 16;
    }
    const int v = coeff / q[nat];
    const int nbits = hd_floor_log2_nz(static_cast<unsigned int>(v < 0 ? -v : v)) + 1;
    hd_atomic_add(&hist[(run << 4) + nbits], weight);
    run = 0;
  }
  if (b >= 64 && run > 0) hd_atomic_add(&hist[0], weight);
}

// One thread per touched block: consumes cnt[b] candidates of the block (plan_edit + the
// body of Search::walk), logs what it overwrote.
struct BulkApply {
  WalkState s;
  const int* touched;
  unsigned int* cnt;        // reset to 0 here
  int* done;                // [nblocks] entries consumed from the block by this bulk (undo, host bookkeeping)
  int* stamp;               // [nblocks] = iter for touched blocks
  int iter;
  int direction;
  unsigned int* delta_hist; // [3][256] symbol count deltas (two's complement)
  unsigned int* chroma_nz;  // [1] delta of the number of nonzero chroma coefficients (two's complement)
  int* log_index;           // undo log: flat coefficient index
  int16_t* log_old;         //           value before
  unsigned int* n_log;
  GB_HD void operator()(int j) const {
    const int b = touched[j];
    const int n = static_cast<int>(cnt[b]);
    cnt[b] = 0u;
    done[b] = n;
    stamp[b] = iter;
    const size_t per = static_cast<size_t>(s.nblocks) * 64;
    for (int t = 0; t < n; ++t) {
      const int li = s.last_index[b];
      const int idx = s.z_idx[static_cast<size_t>(b) * 192 + li + (direction < 0 ? -1 : 0)];
      const int c = idx >> 6, k = idx & 63;
      const int* qc = s.q + 64 * c;
      const int16_t* ob = s.orig + c * per + static_cast<size_t>(b) * 64;
      int16_t* blk = s.cand + c * per + static_cast<size_t>(b) * 64;
      const int newval = direction > 0 ? 0 : quantize_coeff(ob[k], qc[k]);
      const int zp = s.nat2zz[k];
      int za = zp - 1, zb = zp + 1;
      while (za > 0 && blk[s.zz2nat[za]] == 0) --za;
      while (zb < 64 && blk[s.zz2nat[zb]] == 0) ++zb;
      bool precious = false;
      if (k == 1 || k == 8) {
        int sum_of_hf = 0;
        for (int ii = 3; ii < 64; ++ii) {
          if ((ii & 7) < 3 && ii < 3 * 8) continue;
          const int v = ob[ii];
          sum_of_hf += v < 0 ? -v : v;
        }
        const int limit = sum_of_hf < 60 ? 4 : 8;
        const int a = ob[k] < 0 ? -ob[k] : ob[k];
        precious = a >= limit;
      }
      const bool store = !precious || newval != 0;
      unsigned int* h = delta_hist + 256 * c;
      dev_ac_symbols_of_range(blk, qc, s.zz2nat, za, zb, 0xffffffffu, h);
      if (store) {
        const int16_t old = blk[k];
        const unsigned int at = hd_atomic_add(n_log, 1u);
        log_index[at] = static_cast<int>(c * per + static_cast<size_t>(b) * 64 + k);
        log_old[at] = old;
        blk[k] = static_cast<int16_t>(newval);
        if (c > 0) {
          const int d = (newval != 0 ? 1 : 0) - (old != 0 ? 1 : 0);
          if (d != 0) hd_atomic_add(chroma_nz, static_cast<unsigned int>(d));
        }
      }
      dev_ac_symbols_of_range(blk, qc, s.zz2nat, za, zb, 1u, h);
      s.last_index[b] = li + direction;
    }
  }
};

// Rolls a bulk back: coefficient values from the log, candidate cursors from `done`.
struct BulkUndoCoeffs {
  const int* log_index;
  const int16_t* log_old;
  int16_t* cand;
  GB_HD void operator()(int i) const { cand[log_index[i]] = log_old[i]; }
};
struct BulkUndoCursors {
  const int* touched;
  const int* done;
  int* last_index;
  int direction;
  GB_HD void operator()(int j) const {
    const int b = touched[j];
    last_index[b] -= direction * done[b];
  }
};

// State of the blocks the host window walk will touch: coefficients [n][3][64], candidate
// cursor, and whether the bulk of iteration `iter` touched the block.
struct GatherBlockState {
  const int* blocks;
  const int16_t* cand;
  const int* last_index;
  const int* stamp;
  int iter, nblocks;
  int16_t* out_coeffs;  // [n][3][64]
  int* out_cursor;      // [n]
  int* out_in_bulk;     // [n]
  GB_HD void operator()(int i) const {  // i over n * 3 * 8: one row of 8 coefficients each
    const int row = i & 7, ec = i >> 3;
    const int e = ec / 3, c = ec - 3 * e;
    const int b = blocks[e];
    const int16_t* src = cand + (static_cast<size_t>(c) * nblocks + b) * 64 + 8 * row;
    int16_t* dst = out_coeffs + static_cast<size_t>(ec) * 64 + 8 * row;
    for (int k = 0; k < 8; ++k) dst[k] = src[k];
    if (c == 0 && row == 0) {
      out_cursor[e] = last_index[b];
      out_in_bulk[e] = stamp[b] == iter ? 1 : 0;
    }
  }
};

// the host window consumed one more candidate of each listed block (a block may repeat)
struct AdvanceCursors {
  const int* blocks;
  int* last_index;
  int direction;
  GB_HD void operator()(int i) const {
    hd_atomic_add(reinterpret_cast<unsigned int*>(&last_index[blocks[i]]), static_cast<unsigned int>(direction));
  }
};

// max_block_error[b] += block_weight[b] * val_threshold * direction (g/processor.cc:752-755)
struct AddMaxErr {
  float* max_err;
  const float* weight;
  float val_threshold;
  int direction;
  GB_HD void operator()(int b) const { max_err[b] += weight[b] * val_threshold * direction; }
};

}  // namespace gb200
