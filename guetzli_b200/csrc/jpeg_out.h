// Host side of the size/bitstream step (a11) and of the entropy-size model used
// by the selection walk (a16).  Exact integer work; the output bytes must equal
// the reference's WriteJpeg (g/jpeg_data_writer.cc:540) byte for byte.
// g/ = /root/reference/guetzli/.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace gb200 {

// Symbol histogram with every real symbol counted twice and one phantom symbol
// (index 256, count 1) that ends up with the all-ones code
// (g/jpeg_data_writer.h:67-99).
struct SymbolHistogram {
  static const int kSize = 257;
  uint32_t counts[kSize];
  SymbolHistogram() { clear(); }
  void clear() {
    memset(counts, 0, sizeof(counts));
    counts[kSize - 1] = 1;
  }
  void add(int symbol, int weight = 1) { counts[symbol] += 2 * weight; }
  void merge(const SymbolHistogram& o) {
    for (int i = 0; i + 1 < kSize; ++i) counts[i] += o.counts[i];
    counts[kSize - 1] = 1;
  }
  int num_symbols() const {
    int n = 0;
    for (int i = 0; i + 1 < kSize; ++i) n += counts[i] > 0 ? 1 : 0;
    return n;
  }
};

// What surrounds the coded image in the output file (EncodeMetadata and the tail,
// g/jpeg_data_writer.cc:52-74,552): Params::clear_metadata keeps only a JFIF APP0.
struct JpegMeta {
  bool strip = true;
  std::vector<std::string> app_data;  // marker low byte + length + payload, as read
  std::vector<std::string> com_data;  // length + payload
  std::string tail_data;              // bytes after EOI
};

// Component / quant-table structure of a JPEG exactly as it was read (used for the
// "Original Out" file of JPEG input, g/processor.cc:826: the input re-serialised).
struct JpegFileLayout {
  int comp_id[3];
  int comp_table[3];  // position of the component's table in the list below
  int num_tables;
  int table[4][64];
  int precision[4];
  int index[4];  // Tq
};

// Candidate image as the host sees it: dequantised coefficients (multiples of q)
// in [3][nblocks][64] layout plus the quant tables.
struct CoeffImage {
  int w, h, bw, bh, nblocks;
  const int16_t* coeffs;
  int q[3][64];
  // true for the JPEGData produced by EncodeRGBToJpeg itself (g/jpeg_data_encoder.cc:74-83,
  // g/jpeg_data.cc:48): always three components and three un-deduplicated quant tables
  // that all carry table index 0.
  bool as_encoded = false;
  // JPEG input: the file's own structure (3 components, its tables and ids) instead of
  // SaveToJpegData's; null otherwise.
  const JpegFileLayout* as_read = nullptr;
  const JpegMeta* meta = nullptr;  // null = strip (JFIF APP0 only)
  const int16_t* block(int c, int b) const { return coeffs + (static_cast<size_t>(c) * nblocks + b) * 64; }
};

// Length-limited Huffman code lengths (g/entropy_encode.cc:73).
void huffman_code_lengths(const uint32_t* counts, int n, int limit, uint8_t* depth);

size_t histogram_header_bits(const SymbolHistogram& h);                         // :211
size_t histogram_data_bits(const SymbolHistogram& h, const uint8_t* depth);     // :221
// Greedy merge of the trailing histograms while it saves bits (:295). Returns bytes.
size_t cluster_histograms(SymbolHistogram* h, size_t* num, int* index, uint8_t* depth);

int num_output_components(const CoeffImage& img);                     // g/output_image.cc:357
void ac_symbols_of_block(const int16_t* dq_block, const int* q, int weight, SymbolHistogram* h);  // g/processor.cc:471
void ac_symbols_of_range(const int16_t* dq_block, const int* q, int a, int b, int weight, SymbolHistogram* h);
void build_ac_histograms(const CoeffImage& img, SymbolHistogram* h3);  // g/jpeg_data_writer.cc:258
size_t estimate_dc_bytes(const CoeffImage& img);                        // g/processor.cc:528
size_t jpeg_header_bytes(const CoeffImage& img);                        // g/jpeg_data_writer.cc:269
// the same two when the caller already knows the component count / holds the DC histograms
// (device symbol counts): no pass over the coefficients
size_t jpeg_header_bytes(const CoeffImage& img, int ncomp);
size_t estimate_dc_bytes_of(SymbolHistogram* dc_h, int ncomp);
// g/processor.cc:497: per-component depths [3][257] + header bytes of the clustered codes.
// ncomp = number of histograms the reference holds there (jpg.components.size(), :583): 3, or 1 for the
// one-component JPEGData of the YUV420 pass over a grayscale image
size_t compute_entropy_codes(const SymbolHistogram* h3, uint8_t* depths, int ncomp = 3);
size_t entropy_coded_bytes(const SymbolHistogram* h3, const uint8_t* depths);  // g/processor.cc:518

// Everything of the file that precedes the entropy-coded scan (SOI, APP0, DQT,
// SOF1, DHT, SOS) plus the canonical codes of the scan, derived from symbol
// histograms (which the clustering modifies).  Slots: dc0 dc1 dc2 ac0 ac1 ac2.
struct JpegPlan {
  std::string prefix;
  std::string trailer;
  int ncomp;
  uint8_t depth[6][256];
  uint16_t code[6][256];
};
JpegPlan plan_jpeg(const CoeffImage& img, int ncomp, SymbolHistogram* dc_h, SymbolHistogram* ac_h);
void host_symbol_histograms(const CoeffImage& img, int ncomp, SymbolHistogram* dc_h, SymbolHistogram* ac_h);
// EOI plus whatever follows it (the input's tail bytes unless metadata is stripped).
std::string jpeg_trailer(const CoeffImage& img);

// SaveToJpegData + WriteJpeg (strip_metadata path): the complete JPEG file.
std::string write_jpeg(const CoeffImage& img);

}  // namespace gb200
