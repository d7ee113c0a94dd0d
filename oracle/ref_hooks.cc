// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-ABI hooks over the UNMODIFIED google/guetzli reference, compiled from the
// sources where they lie under /root/reference (never copied into this repo).
// The product (guetzli_b200/) must never link, import or call anything here;
// only tests/, __graft_entry__.smoke() and bench.py's CPU arm do.
//
// The two reference translation units that keep their interesting functions
// file-local (butteraugli.cc: SeparateFrequencies, MaltaDiffMapImpl, ...;
// processor.cc: Processor::ComputeBlockZeroingOrder, ...) are #included into
// this TU so that the hooks can reach them; the arithmetic is the reference's
// own object code either way.  Build recipe: oracle/Makefile.
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

// reference sources, in place
#include "butteraugli/butteraugli.cc"  // third_party/butteraugli/butteraugli/butteraugli.cc

#define private public
#include "guetzli/processor.cc"
#undef private

#include "guetzli/entropy_encode.h"
#include "guetzli/fdct.h"
#include "guetzli/idct.h"
#include "guetzli/jpeg_data_encoder.h"
#include "guetzli/quality.h"
#include "guetzli/color_transform.h"
#include "guetzli/gamma_correct.h"

using butteraugli::ImageF;

namespace {

std::vector<ImageF> PlanesFromFlat(const float* p, int w, int h, int n) {
  std::vector<ImageF> planes = butteraugli::CreatePlanes<float>(w, h, n);
  for (int c = 0; c < n; ++c)
    for (int y = 0; y < h; ++y)
      memcpy(planes[c].Row(y), p + ((size_t)c * h + y) * w, sizeof(float) * w);
  return planes;
}

void FlatFromPlane(const ImageF& im, float* out) {
  for (size_t y = 0; y < im.ysize(); ++y)
    memcpy(out + y * im.xsize(), im.Row(y), sizeof(float) * im.xsize());
}

// JPEGData holding the given YUV444 coefficients with all-ones quant tables
// (what Processor sees after RemoveOriginalQuantization).
void JpegFromCoeffs(const int16_t* coeffs, int w, int h, guetzli::JPEGData* jpg) {
  guetzli::InitJPEGDataForYUV444(w, h, jpg);
  guetzli::AddApp0Data(jpg);
  for (int c = 0; c < 3; ++c) {
    guetzli::JPEGComponent& comp = jpg->components[c];
    memcpy(comp.coeffs.data(), coeffs + (size_t)c * comp.num_blocks * 64,
           sizeof(int16_t) * comp.num_blocks * 64);
    for (int k = 0; k < 64; ++k) jpg->quant[c].values[k] = 1;
  }
}

}  // namespace

extern "C" {

double gref_target_for_quality(double q) {
  return guetzli::ButteraugliScoreForQuality(q);
}

void gref_free(void* p) { free(p); }

// guetzli::Process(RGB) (processor.cc:926). counters = {iterations, up, down}.
// Params::clear_metadata for the gref_process_rgb* hooks (default: the reference's, true).
static int g_clear_metadata = 1;
void gref_set_clear_metadata(int on) { g_clear_metadata = on; }

// Params::try_420 / force_420 for the gref_process_rgb* hooks (processor.h:32-33; default false).
static int g_try_420 = 0, g_force_420 = 0;
void gref_set_420(int try_420, int force_420) {
  g_try_420 = try_420;
  g_force_420 = force_420;
}

// _ex: also sets Params::zeroing_greedy_lookahead / new_zeroing_model (processor.h:35-36).
int gref_process_rgb_ex(const uint8_t* rgb, int w, int h, float butteraugli_target, int lookahead, int new_model,
                        uint8_t** out, size_t* out_len, char** trace, size_t* trace_len, int* counters,
                        double* seconds) {
  guetzli::Params params;
  params.butteraugli_target = butteraugli_target;
  params.zeroing_greedy_lookahead = lookahead;
  params.new_zeroing_model = new_model != 0;
  params.clear_metadata = g_clear_metadata != 0;
  params.try_420 = g_try_420 != 0;
  params.force_420 = g_force_420 != 0;
  guetzli::ProcessStats stats;
  std::string dbg;
  if (trace) stats.debug_output = &dbg;
  std::vector<uint8_t> v(rgb, rgb + (size_t)3 * w * h);
  std::string jpg;
  auto t0 = std::chrono::steady_clock::now();
  bool ok = guetzli::Process(params, &stats, v, w, h, &jpg);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  *out = (uint8_t*)malloc(jpg.size() + 1);
  memcpy(*out, jpg.data(), jpg.size());
  *out_len = jpg.size();
  if (trace) {
    *trace = (char*)malloc(dbg.size() + 1);
    memcpy(*trace, dbg.data(), dbg.size());
    (*trace)[dbg.size()] = 0;
    *trace_len = dbg.size();
  }
  if (counters) {
    counters[0] = stats.counters[guetzli::kNumItersCnt];
    counters[1] = stats.counters[guetzli::kNumItersUpCnt];
    counters[2] = stats.counters[guetzli::kNumItersDownCnt];
  }
  return ok ? 1 : 0;
}

// butteraugli::ButteraugliInterface (butteraugli.cc:1858): planar linear RGB [3][h][w].
int gref_butteraugli_interface(const float* rgb0, const float* rgb1, int w, int h, float* diffmap, double* score) {
  std::vector<butteraugli::ImageF> a, b;
  for (int c = 0; c < 3; ++c) {
    a.emplace_back(w, h);
    b.emplace_back(w, h);
    for (int y = 0; y < h; ++y) {
      memcpy(a[c].Row(y), rgb0 + ((size_t)c * h + y) * w, sizeof(float) * w);
      memcpy(b[c].Row(y), rgb1 + ((size_t)c * h + y) * w, sizeof(float) * w);
    }
  }
  butteraugli::ImageF dm;
  double v = 0;
  if (!butteraugli::ButteraugliInterface(a, b, dm, v)) return 0;
  for (int y = 0; y < h; ++y) memcpy(diffmap + (size_t)y * w, dm.Row(y), sizeof(float) * w);
  *score = v;
  return 1;
}

// The tool's heat map (butteraugli.cc:1979 CreateHeatMapImage with the thresholds of
// butteraugli_main.cc:423-424).
void gref_heatmap(const float* distmap, int w, int h, uint8_t* rgb) {
  std::vector<float> dm(distmap, distmap + (size_t)w * h);
  std::vector<uint8_t> out;
  butteraugli::CreateHeatMapImage(dm, butteraugli::ButteraugliFuzzyInverse(1.5), butteraugli::ButteraugliFuzzyInverse(0.5),
                                  w, h, &out);
  memcpy(rgb, out.data(), out.size());
}

// guetzli::Process(jpeg bytes) (processor.cc:890).  clear_metadata: Params::clear_metadata.
int gref_process_jpeg(const uint8_t* jpeg, size_t len, float butteraugli_target, int clear_metadata,
                      uint8_t** out, size_t* out_len, char** trace, size_t* trace_len, int* counters) {
  guetzli::Params params;
  params.butteraugli_target = butteraugli_target;
  params.clear_metadata = clear_metadata != 0;
  params.try_420 = g_try_420 != 0;
  params.force_420 = g_force_420 != 0;
  guetzli::ProcessStats stats;
  std::string dbg;
  if (trace) stats.debug_output = &dbg;
  std::string in(reinterpret_cast<const char*>(jpeg), len), jpg;
  bool ok = guetzli::Process(params, &stats, in, &jpg);
  *out = (uint8_t*)malloc(jpg.size() + 1);
  memcpy(*out, jpg.data(), jpg.size());
  *out_len = jpg.size();
  if (trace) {
    *trace = (char*)malloc(dbg.size() + 1);
    memcpy(*trace, dbg.data(), dbg.size());
    (*trace)[dbg.size()] = 0;
    *trace_len = dbg.size();
  }
  if (counters) {
    counters[0] = stats.counters[guetzli::kNumItersCnt];
    counters[1] = stats.counters[guetzli::kNumItersUpCnt];
    counters[2] = stats.counters[guetzli::kNumItersDownCnt];
  }
  return ok ? 1 : 0;
}

// ReadJpeg(JPEG_READ_ALL) (jpeg_data_reader.cc:931): quantised coefficients of component c
// into out (caller sizes it from dims[]).  dims = {w, h, ncomp, wb0, hb0, wb1, hb1, ...}.
int gref_read_jpeg(const uint8_t* jpeg, size_t len, int* dims, int16_t* out, size_t out_cap) {
  guetzli::JPEGData jpg;
  std::string in(reinterpret_cast<const char*>(jpeg), len);
  if (!guetzli::ReadJpeg(in, guetzli::JPEG_READ_ALL, &jpg)) return 0;
  dims[0] = jpg.width;
  dims[1] = jpg.height;
  dims[2] = (int)jpg.components.size();
  size_t pos = 0;
  for (size_t c = 0; c < jpg.components.size(); ++c) {
    dims[3 + 2 * c] = jpg.components[c].width_in_blocks;
    dims[4 + 2 * c] = jpg.components[c].height_in_blocks;
    for (size_t i = 0; i < jpg.components[c].coeffs.size(); ++i) {
      if (pos < out_cap) out[pos] = jpg.components[c].coeffs[i];
      ++pos;
    }
  }
  return pos <= out_cap ? 1 : 0;
}

int gref_process_rgb(const uint8_t* rgb, int w, int h, float butteraugli_target,
                     uint8_t** out, size_t* out_len, char** trace,
                     size_t* trace_len, int* counters, double* seconds) {
  const guetzli::Params d;
  return gref_process_rgb_ex(rgb, w, h, butteraugli_target, d.zeroing_greedy_lookahead, d.new_zeroing_model ? 1 : 0, out,
                             out_len, trace, trace_len, counters, seconds);
}

// EncodeRGBToJpeg (jpeg_data_encoder.cc:66): coeffs = 3 planes of B*64 int16.
int gref_rgb_to_coeffs(const uint8_t* rgb, int w, int h, int16_t* coeffs) {
  guetzli::JPEGData jpg;
  std::vector<uint8_t> v(rgb, rgb + (size_t)3 * w * h);
  if (!guetzli::EncodeRGBToJpeg(v, w, h, &jpg)) return 0;
  for (int c = 0; c < 3; ++c) {
    const guetzli::JPEGComponent& comp = jpg.components[c];
    memcpy(coeffs + (size_t)c * comp.num_blocks * 64, comp.coeffs.data(),
           sizeof(int16_t) * comp.num_blocks * 64);
  }
  return 1;
}

void gref_fdct_block(int16_t* block) { guetzli::ComputeBlockDCT(block); }
void gref_idct_block(const int16_t* in, uint8_t* out) {
  guetzli::ComputeBlockIDCT(in, out);
}
int gref_quantize(int coeff, int q) { return guetzli::Quantize((int16_t)coeff, q); }
void gref_ycbcr_to_rgb(uint8_t* px) { guetzli::ColorTransformYCbCrToRGB(px); }
void gref_srgb_lut(double* out256) {
  memcpy(out256, guetzli::Srgb8ToLinearTable(), 256 * sizeof(double));
}

// OutputImage render of dequantised coefficients (output_image.cc:250,411,427).
// srgb: interleaved 3N u8; linear: 3 planes of N float.
void gref_render(const int16_t* coeffs, int w, int h, uint8_t* srgb, float* linear) {
  guetzli::JPEGData jpg;
  JpegFromCoeffs(coeffs, w, h, &jpg);
  guetzli::OutputImage img(w, h);
  img.CopyFromJpegData(jpg);
  if (srgb) {
    std::vector<uint8_t> s = img.ToSRGB();
    memcpy(srgb, s.data(), s.size());
  }
  if (linear) {
    std::vector<std::vector<float> > lin(3, std::vector<float>((size_t)w * h));
    img.ToLinearRGB(&lin);
    for (int c = 0; c < 3; ++c)
      memcpy(linear + (size_t)c * w * h, lin[c].data(), sizeof(float) * w * h);
  }
}

// OutputImage::ApplyGlobalQuantization (output_image.cc:342) on raw coeffs.
void gref_apply_global_quant(const int16_t* coeffs, int w, int h, const int* q,
                             int16_t* out) {
  guetzli::JPEGData jpg;
  JpegFromCoeffs(coeffs, w, h, &jpg);
  guetzli::OutputImage img(w, h);
  img.CopyFromJpegData(jpg);
  int qq[3][64];
  memcpy(qq, q, sizeof(qq));
  img.ApplyGlobalQuantization(qq);
  for (int c = 0; c < 3; ++c) {
    size_t n = (size_t)img.component(c).width_in_blocks() *
               img.component(c).height_in_blocks() * 64;
    memcpy(out + c * n, img.component(c).coeffs(), n * sizeof(int16_t));
  }
}

// butteraugli::Blur (butteraugli.cc:229)
void gref_blur(const float* in, int w, int h, float sigma, float border_ratio,
               float* out) {
  std::vector<ImageF> p = PlanesFromFlat(in, w, h, 1);
  ImageF b = butteraugli::Blur(p[0], sigma, border_ratio);
  FlatFromPlane(b, out);
}

// butteraugli::OpsinDynamicsImage (butteraugli.cc:324)
void gref_opsin(const float* rgb, int w, int h, float* xyb) {
  std::vector<ImageF> p = PlanesFromFlat(rgb, w, h, 3);
  std::vector<ImageF> x = butteraugli::OpsinDynamicsImage(p);
  for (int c = 0; c < 3; ++c) FlatFromPlane(x[c], xyb + (size_t)c * w * h);
}

// SeparateFrequencies (butteraugli.cc:489). out: uhf0 uhf1 hf0 hf1 mf0 mf1 mf2 lf0 lf1 lf2
void gref_separate(const float* xyb, int w, int h, float* out) {
  std::vector<ImageF> p = PlanesFromFlat(xyb, w, h, 3);
  butteraugli::PsychoImage ps;
  butteraugli::SeparateFrequencies(w, h, p, ps);
  size_t n = (size_t)w * h;
  FlatFromPlane(ps.uhf[0], out + 0 * n);
  FlatFromPlane(ps.uhf[1], out + 1 * n);
  FlatFromPlane(ps.hf[0], out + 2 * n);
  FlatFromPlane(ps.hf[1], out + 3 * n);
  for (int c = 0; c < 3; ++c) FlatFromPlane(ps.mf[c], out + (4 + c) * n);
  for (int c = 0; c < 3; ++c) FlatFromPlane(ps.lf[c], out + (7 + c) * n);
}

// MaltaDiffMapImpl (butteraugli.cc:1461): out += malta(lum0, lum1)
void gref_malta(const float* lum0, const float* lum1, int w, int h, double w_0gt1,
                double w_0lt1, double norm1, int lf, float* inout) {
  std::vector<ImageF> a = PlanesFromFlat(lum0, w, h, 1);
  std::vector<ImageF> b = PlanesFromFlat(lum1, w, h, 1);
  std::vector<ImageF> o = PlanesFromFlat(inout, w, h, 1);
  if (lf)
    butteraugli::MaltaDiffMapImpl<butteraugli::MaltaTagLF>(
        a[0], b[0], w, h, w_0gt1, w_0lt1, norm1, 3.75, 0.405371989604, &o[0]);
  else
    butteraugli::MaltaDiffMapImpl<butteraugli::MaltaTag>(
        a[0], b[0], w, h, w_0gt1, w_0lt1, norm1, 3.75, 0.354191303559, &o[0]);
  FlatFromPlane(o[0], inout);
}

// butteraugli::Mask (butteraugli.cc:1741)
void gref_mask(const float* xyb0, const float* xyb1, int w, int h, float* mask,
               float* mask_dc) {
  std::vector<ImageF> a = PlanesFromFlat(xyb0, w, h, 3);
  std::vector<ImageF> b = PlanesFromFlat(xyb1, w, h, 3);
  std::vector<ImageF> m, mdc;
  butteraugli::Mask(a, b, &m, &mdc);
  for (int c = 0; c < 3; ++c) {
    FlatFromPlane(m[c], mask + (size_t)c * w * h);
    FlatFromPlane(mdc[c], mask_dc + (size_t)c * w * h);
  }
}

// butteraugli::ButteraugliComparator(rgb0).Diffmap(rgb1) (butteraugli.cc:784,799)
void gref_diffmap(const float* rgb0, const float* rgb1, int w, int h, float* distmap) {
  std::vector<ImageF> a = PlanesFromFlat(rgb0, w, h, 3);
  std::vector<ImageF> b = PlanesFromFlat(rgb1, w, h, 3);
  butteraugli::ButteraugliComparator cmp(a);
  ImageF d;
  cmp.Diffmap(b, d);
  FlatFromPlane(d, distmap);
}

// guetzli::ButteraugliComparator::Compare (butteraugli_comparator.cc:63) of a
// candidate given as dequantised coefficients.
void gref_compare_coeffs(const uint8_t* rgb_orig, int w, int h, float target,
                         const int16_t* coeffs, float* distmap, float* distance) {
  std::vector<uint8_t> v(rgb_orig, rgb_orig + (size_t)3 * w * h);
  guetzli::ProcessStats stats;
  guetzli::ButteraugliComparator cmp(w, h, &v, target, &stats);
  guetzli::JPEGData jpg;
  JpegFromCoeffs(coeffs, w, h, &jpg);
  guetzli::OutputImage img(w, h);
  img.CopyFromJpegData(jpg);
  cmp.Compare(img);
  std::vector<float> d = cmp.distmap();
  if (distmap) memcpy(distmap, d.data(), d.size() * sizeof(float));
  *distance = cmp.distmap_aggregate();
}

// StartBlockComparisons (butteraugli_comparator.cc:415): mask_xyz_ planes.
void gref_block_mask(const uint8_t* rgb_orig, int w, int h, float* mask_xyz) {
  std::vector<uint8_t> v(rgb_orig, rgb_orig + (size_t)3 * w * h);
  guetzli::ProcessStats stats;
  guetzli::ButteraugliComparator cmp(w, h, &v, 1.0f, &stats);
  cmp.StartBlockComparisons();
  for (int c = 0; c < 3; ++c)
    FlatFromPlane(cmp.mask_xyz_[c], mask_xyz + (size_t)c * w * h);
}

// The per-block loop of SelectFrequencyMasking (processor.cc:554-590) for
// comp_mask=7 on a YUV444 image quantised with q.  Returns total candidates.
// offsets: B+1 ints; idx/err: capacity 189*B.
int gref_zeroing_orders(const uint8_t* rgb_orig, int w, int h, float target,
                        const int16_t* orig_coeffs, const int* q, int* offsets,
                        uint8_t* idx, float* err) {
  std::vector<uint8_t> v(rgb_orig, rgb_orig + (size_t)3 * w * h);
  guetzli::ProcessStats stats;
  guetzli::ButteraugliComparator cmp(w, h, &v, target, &stats);
  guetzli::JPEGData jpg;
  JpegFromCoeffs(orig_coeffs, w, h, &jpg);
  guetzli::OutputImage img(w, h);
  img.CopyFromJpegData(jpg);
  int qq[3][64];
  memcpy(qq, q, sizeof(qq));
  img.ApplyGlobalQuantization(qq);
  guetzli::Processor proc;
  proc.params_ = guetzli::Params();
  proc.params_.butteraugli_target = target;
  proc.comparator_ = &cmp;
  proc.stats_ = &stats;
  const int bw = (w + 7) / 8, bh = (h + 7) / 8;
  cmp.StartBlockComparisons();
  int total = 0;
  std::vector<guetzli::CoeffData> order;
  for (int by = 0, b = 0; by < bh; ++by) {
    for (int bx = 0; bx < bw; ++bx, ++b) {
      guetzli::coeff_t block[192] = {0}, orig_block[192] = {0};
      for (int c = 0; c < 3; ++c) {
        img.component(c).GetCoeffBlock(bx, by, &block[c * 64]);
        memcpy(&orig_block[c * 64], &jpg.components[c].coeffs[(size_t)b * 64],
               64 * sizeof(int16_t));
      }
      order.clear();
      proc.ComputeBlockZeroingOrder(block, orig_block, bx, by, 1, 1, 7, &img, &order);
      offsets[b] = total;
      for (size_t i = 0; i < order.size(); ++i) {
        idx[total] = (uint8_t)order[i].idx;
        err[total] = order[i].block_err;
        ++total;
      }
    }
  }
  offsets[bw * bh] = total;
  cmp.FinishBlockComparisons();
  return total;
}

// ComputeBlockErrorAdjustmentWeights (butteraugli_comparator.cc:494)
void gref_block_weights(int w, int h, float target, int direction, int max_block_dist,
                        double target_mul, const float* distmap, float* weights) {
  std::vector<uint8_t> v((size_t)3 * w * h);
  guetzli::ProcessStats stats;
  guetzli::ButteraugliComparator cmp(w, h, &v, target, &stats);
  std::vector<float> d(distmap, distmap + (size_t)w * h);
  const int bw = (w + 7) / 8, bh = (h + 7) / 8;
  std::vector<float> wt(bw * bh);
  cmp.ComputeBlockErrorAdjustmentWeights(direction, max_block_dist, target_mul, 1, 1,
                                         d, &wt);
  memcpy(weights, wt.data(), wt.size() * sizeof(float));
}

// OutputImage::SaveToJpegData + WriteJpeg (output_image.cc:348,
// jpeg_data_writer.cc:540) of dequantised coeffs that are multiples of q.
int gref_write_jpeg(const int16_t* coeffs, int w, int h, const int* q, uint8_t** out,
                    size_t* out_len) {
  guetzli::JPEGData jpg;
  JpegFromCoeffs(coeffs, w, h, &jpg);
  guetzli::OutputImage img(w, h);
  img.CopyFromJpegData(jpg);
  int qq[3][64];
  memcpy(qq, q, sizeof(qq));
  img.ApplyGlobalQuantization(qq);
  guetzli::JPEGData jpg_out = jpg;
  img.SaveToJpegData(&jpg_out);
  std::string s;
  guetzli::JPEGOutput o(guetzli::GuetzliStringOut, &s);
  if (!guetzli::WriteJpeg(jpg_out, true, o)) return 0;
  *out = (uint8_t*)malloc(s.size() + 1);
  memcpy(*out, s.data(), s.size());
  *out_len = s.size();
  return 1;
}

// Table dumps used to pin the generated tables of the product.
void gref_color_tables(int* cr_r, int* cb_b, int* cr_g, int* cb_g, uint8_t* range_limit) {
  memcpy(cr_r, guetzli::kCrToRedTable, 256 * sizeof(int));
  memcpy(cb_b, guetzli::kCbToBlueTable, 256 * sizeof(int));
  memcpy(cr_g, guetzli::kCrToGreenTable, 256 * sizeof(int));
  memcpy(cb_g, guetzli::kCbToGreenTable, 256 * sizeof(int));
  memcpy(range_limit, guetzli::kRangeLimitLut, 4 * 256);
}

void gref_blur_kernel(float sigma, float* out, int* len) {
  std::vector<float> k = butteraugli::ComputeKernel(sigma);
  *len = (int)k.size();
  memcpy(out, k.data(), k.size() * sizeof(float));
}

double gref_mask_lut(int which, double delta) {
  switch (which) {
    case 0: return butteraugli::MaskX(delta);
    case 1: return butteraugli::MaskY(delta);
    case 2: return butteraugli::MaskDcX(delta);
    default: return butteraugli::MaskDcY(delta);
  }
}

double gref_gamma(double v) { return butteraugli::Gamma(v); }

// CreateHuffmanTree (guetzli/entropy_encode.cc:73): depth[n] zeroed by the caller
void gref_huffman_depths(const uint32_t* counts, int n, int limit, uint8_t* depth) {
  std::vector<guetzli::HuffmanTree> tree(2 * static_cast<size_t>(n) + 1);
  guetzli::CreateHuffmanTree(counts, static_cast<size_t>(n), limit, tree.data(), depth);
}

}  // extern "C"
