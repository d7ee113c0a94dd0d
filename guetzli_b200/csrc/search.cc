// See search.h.  Everything O(pixels) or O(blocks) runs in ImageContext kernels;
// this file is scalar control flow plus the (for now host-side) sequential
// selection walk and JPEG serialisation.
#include "search.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <set>
#include <vector>

#include "jpeg_out.h"
#include "pipeline.h"
#include "tables.h"

namespace gb200 {

long total_launches();
long long h2d_bytes_total();
long long d2h_bytes_total();

double score_jpeg(double distance, int size, double target) {
  const double kScale = 50, kMaxExponent = 10, kLargeSize = 1e30;
  const double diff = distance - target;
  if (diff <= 0.0) return size;
  const double exponent = kScale * diff;
  if (exponent > kMaxExponent) return kLargeSize * std::exp(kMaxExponent) * diff + size;
  return std::exp(exponent) * size;
}

namespace {

typedef std::chrono::steady_clock Clock;
double ms_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

struct QuantTrial {
  int q[3][64];
  size_t jpg_size;
  bool dist_ok;
};

// -1 / 0 / 1 when a <= b / a == b / a >= b coordinate-wise, 2 when incomparable
// (g/processor.cc:161).
int compare_quant(const int* a, const int* b) {
  int i = 0;
  while (i < 192 && a[i] == b[i]) ++i;
  if (i == 192) return 0;
  if (a[i] < b[i]) {
    for (++i; i < 192; ++i)
      if (a[i] > b[i]) return 2;
    return -1;
  }
  for (++i; i < 192; ++i)
    if (a[i] < b[i]) return 2;
  return 1;
}

double contrast_sensitivity(int k) { return 1.0 / (1.0 + natural_to_zigzag()[k] / 2.0); }

double quant_heuristic_score(const int q[3][64]) {
  double score = 0.0;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 64; ++k) score += 0.5 * (q[c][k] - 1.0) * contrast_sensitivity(k);
  return score;
}

// Bisection over a scalar "heuristic score" that indexes a one-parameter family
// of quant matrices (g/processor.cc:194-296).
class QuantBisection {
 public:
  QuantBisection() : lo_(-1.0), hi_(-1.0), total_csf_(0.0) {
    for (int k = 0; k < 64; ++k) total_csf_ += 3.0 * contrast_sensitivity(k);
  }

  bool next(int q[3][64]) {
    for (int iter = 0; iter < 1000; ++iter) {
      double hscore;
      if (hi_ == -1.0) {
        if (lo_ == -1.0) {
          hscore = total_csf_;
        } else if (lo_ < 5.0 * total_csf_) {
          hscore = lo_ + total_csf_;
        } else {
          hscore = 2 * (lo_ + total_csf_);
        }
        if (hscore > 100 * total_csf_) return false;
      } else if (hi_ == 0.0) {
        return false;
      } else if (lo_ == -1.0) {
        hscore = 0.0;
      } else {
        int lower_q[3][64], upper_q[3][64];
        const double kEps = 0.05;
        matrix_for_score((1 - kEps) * lo_ + kEps * 0.5 * (lo_ + hi_), lower_q);
        matrix_for_score((1 - kEps) * hi_ + kEps * 0.5 * (lo_ + hi_), upper_q);
        if (compare_quant(&lower_q[0][0], &upper_q[0][0]) == 0) return false;
        hscore = (lo_ + hi_) * 0.5;
      }
      matrix_for_score(hscore, q);
      bool retry = false;
      for (size_t i = 0; i < tried_.size(); ++i) {
        if (compare_quant(&q[0][0], &tried_[i].q[0][0]) == 0) {
          if (tried_[i].dist_ok) lo_ = hscore; else hi_ = hscore;
          retry = true;
          break;
        }
      }
      if (!retry) return true;
    }
    return false;
  }

  void add(const QuantTrial& t) {
    tried_.push_back(t);
    const double hscore = quant_heuristic_score(t.q);
    if (t.dist_ok) {
      lo_ = std::max(lo_, hscore);
    } else {
      hi_ = hi_ == -1.0 ? hscore : std::min(hi_, hscore);
    }
  }

 private:
  void matrix_for_score(double score, int q[3][64]) const {
    const int level = static_cast<int>(score / total_csf_);
    score -= level * total_csf_;
    const int* zz = zigzag_to_natural();
    for (int k = 63; k >= 0; --k) {
      for (int c = 0; c < 3; ++c) q[c][zz[k]] = 2 * level + (score > 0.0 ? 3 : 1);
      score -= 3.0 * contrast_sensitivity(zz[k]);
    }
  }
  double lo_, hi_, total_csf_;
  std::vector<QuantTrial> tried_;
};

class Search {
 public:
  Search(const SearchParams& p, ImageContext* ctx, LogSink log, void* log_user, SearchStats* st)
      : params_(p), ctx_(ctx), log_(log), log_user_(log_user), st_(st), best_score_(-1.0), distance_(0.0f) {
    const Geom& g = ctx->geom();
    img_.w = g.w;
    img_.h = g.h;
    img_.bw = g.bw;
    img_.bh = g.bh;
    img_.nblocks = g.nblocks;
    cand_.assign(ctx->orig_coeffs().begin(), ctx->orig_coeffs().end());
    img_.coeffs = cand_.data();
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 64; ++k) img_.q[c][k] = 1;
  }

  void run(std::string* best_out) {
    best_ = best_out;
    const float target = params_.butteraugli_target;
    // the q=1 JPEG is the fallback output (g/processor.cc:826-846)
    img_.as_encoded = true;
    std::string encoded = timed_write();
    img_.as_encoded = false;
    logf("Original Out[%7zd]", encoded.size());
    compare();
    maybe_output(encoded);
    int best_q[3][64];
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 64; ++k) best_q[c][k] = 1;
    if (!select_quant_matrix(best_q)) {
      for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 64; ++k) best_q[c][k] = 1;
    }
    set_global_quant(best_q);
    select_frequency_masking(1.0);
    (void)target;
  }

 private:
  void logf(const char* fmt, ...) {
    if (!log_) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    log_(log_user_, buf);
  }
  void log_quant(const int q[3][64]) {
    for (int y = 0; y < 8; ++y) {
      for (int c = 0; c < 3; ++c) {
        for (int x = 0; x < 8; ++x) logf(" %2d", q[c][8 * y + x]);
        logf("   ");
      }
      logf("\n");
    }
  }

  std::string timed_write() {
    Clock::time_point t0 = Clock::now();
    std::string s = write_jpeg(img_);
    st_->ms_jpeg += ms_since(t0);
    return s;
  }

  void compare() {
    Clock::time_point t0 = Clock::now();
    distance_ = ctx_->compare();
    st_->ms_compare += ms_since(t0);
    ++st_->compares;
    logf(" BA[100.00%%] D[%6.4f]", distance_);
  }

  bool distance_ok(double target_mul) const {
    return distance_ <= target_mul * params_.butteraugli_target;
  }

  void maybe_output(const std::string& encoded) {
    const double score = score_jpeg(distance_, static_cast<int>(encoded.size()), params_.butteraugli_target);
    logf(" Score[%.4f]", score);
    if (score < best_score_ || best_score_ < 0) {
      *best_ = encoded;
      best_score_ = score;
      logf(" (*)");
    }
    logf("\n");
  }

  // candidate := Quantize(original, q) on both sides of the bus
  void set_global_quant(const int q[3][64]) {
    const std::vector<int16_t>& orig = ctx_->orig_coeffs();
    const size_t per = static_cast<size_t>(img_.nblocks) * 64;
    for (int c = 0; c < 3; ++c) {
      const int16_t* src = &orig[c * per];
      int16_t* dst = &cand_[c * per];
      for (size_t i = 0; i < per; ++i) dst[i] = static_cast<int16_t>(quantize_coeff(src[i], q[c][i & 63]));
    }
    memcpy(img_.q, q, sizeof(img_.q));
    ctx_->apply_global_quant(&q[0][0]);
  }

  QuantTrial try_quant_matrix(const float target_mul, int q[3][64]) {
    QuantTrial data;
    memcpy(data.q, q, sizeof(data.q));
    set_global_quant(q);
    std::string encoded = timed_write();
    logf("Iter %2d: %s quantization matrix:\n", st_->iterations + 1, "f111111");
    log_quant(q);
    logf("Iter %2d: %s GQ[%5.2f] Out[%7zd]", st_->iterations + 1, "f111111", quant_heuristic_score(q),
         encoded.size());
    ++st_->iterations;
    compare();
    data.dist_ok = distance_ok(target_mul);
    data.jpg_size = encoded.size();
    maybe_output(encoded);
    return data;
  }

  static bool better(const QuantTrial& a, const QuantTrial& b) {
    if (a.dist_ok && !b.dist_ok) return true;
    if (!a.dist_ok && b.dist_ok) return false;
    return a.jpg_size < b.jpg_size;
  }

  bool select_quant_matrix(int best_q[3][64]) {
    QuantBisection gen;
    const float target_mul_high = 0.97f, target_mul_low = 0.95f;
    QuantTrial best = try_quant_matrix(target_mul_high, best_q);
    for (;;) {
      int q_next[3][64];
      if (!gen.next(q_next)) break;
      QuantTrial data = try_quant_matrix(target_mul_high, q_next);
      gen.add(data);
      if (better(data, best)) {
        best = data;
        if (data.dist_ok && !distance_ok(target_mul_low)) break;
      }
    }
    memcpy(&best_q[0][0], &best.q[0][0], sizeof(best.q));
    logf("\n%s selected quantization matrix:\n", "YUV444");
    log_quant(best_q);
    return best.dist_ok;
  }

  void select_frequency_masking(const double target_mul) {
    const int num_blocks = img_.nblocks;
    const int block_width = img_.bw;
    const std::vector<int16_t>& orig = ctx_->orig_coeffs();
    const size_t per = static_cast<size_t>(num_blocks) * 64;

    // a13 + a14 on the device: per-block candidate lists
    std::vector<int> offsets(num_blocks + 1);
    std::vector<uint8_t> cand_idx;
    std::vector<float> cand_err;
    {
      Clock::time_point t0 = Clock::now();
      std::vector<uint8_t> idx;
      std::vector<float> err;
      std::vector<int> count;
      ctx_->zeroing_orders(params_.butteraugli_target, params_.zeroing_greedy_lookahead, &idx, &err, &count);
      size_t total = 0;
      for (int b = 0; b < num_blocks; ++b) total += count[b];
      cand_idx.reserve(total);
      cand_err.reserve(total);
      for (int b = 0; b < num_blocks; ++b) {
        offsets[b] = static_cast<int>(cand_idx.size());
        cand_idx.insert(cand_idx.end(), &idx[static_cast<size_t>(b) * 192], &idx[static_cast<size_t>(b) * 192] + count[b]);
        cand_err.insert(cand_err.end(), &err[static_cast<size_t>(b) * 192], &err[static_cast<size_t>(b) * 192] + count[b]);
      }
      offsets[num_blocks] = static_cast<int>(cand_idx.size());
      st_->ms_zeroing += ms_since(t0);
    }

    SymbolHistogram ac_h[3];
    const int header_size = static_cast<int>(jpeg_header_bytes(img_));
    const int dc_size = static_cast<int>(estimate_dc_bytes(img_));
    build_ac_histograms(img_, ac_h);
    std::vector<uint8_t> ac_depths(3 * SymbolHistogram::kSize);
    int ac_histogram_size = static_cast<int>(compute_entropy_codes(ac_h, ac_depths.data()));
    const int base_size =
        header_size + dc_size + ac_histogram_size + static_cast<int>(entropy_coded_bytes(ac_h, ac_depths.data()));
    int prev_size = base_size;

    std::vector<float> max_block_error(num_blocks);
    std::vector<int> last_indexes(num_blocks);
    std::vector<float> block_weight(num_blocks);
    std::vector<int> edit_index;
    std::vector<int16_t> edit_value;

    bool first_up_iter = true;
    const int directions[2] = {1, -1};
    for (int di = 0; di < 2; ++di) {
      const int direction = directions[di];
      for (;;) {
        std::vector<std::pair<int, float> > global_order;
        int blocks_to_change = 0;
        for (int rblock = 1; rblock <= 4; ++rblock) {
          ctx_->block_weights(direction, rblock, params_.butteraugli_target * target_mul, first_up_iter,
                              block_weight.data());
          global_order.clear();
          blocks_to_change = 0;
          for (int block_ix = 0; block_ix < num_blocks; ++block_ix) {
            const int last_index = last_indexes[block_ix];
            const int offset = offsets[block_ix];
            const int num_candidates = offsets[block_ix + 1] - offset;
            const float* candidate_errors = &cand_err[offset];
            const float max_err = max_block_error[block_ix];
            if (block_weight[block_ix] == 0) continue;
            if (direction > 0) {
              for (int i = last_index; i < num_candidates; ++i) {
                const float val = (candidate_errors[i] - max_err) / block_weight[block_ix];
                global_order.push_back(std::make_pair(block_ix, val));
              }
              blocks_to_change += (last_index < num_candidates ? 1 : 0);
            } else {
              for (int i = last_index - 1; i >= 0; --i) {
                const float val = (max_err - candidate_errors[i]) / block_weight[block_ix];
                global_order.push_back(std::make_pair(block_ix, val));
              }
              blocks_to_change += (last_index > 0 ? 1 : 0);
            }
          }
          if (!global_order.empty()) break;
        }
        if (global_order.empty()) break;

        {
          Clock::time_point t0 = Clock::now();
          std::sort(global_order.begin(), global_order.end(),
                    [](const std::pair<int, float>& a, const std::pair<int, float>& b) {
                      return a.second < b.second;
                    });
          st_->ms_sort += ms_since(t0);
        }

        double rel_size_delta = direction > 0 ? 0.01 : 0.0005;
        if (direction > 0 && distance_ok(1.0)) rel_size_delta = 0.05;
        const double min_size_delta = base_size * rel_size_delta;
        const float coeffs_to_change_per_block = direction > 0 ? 2.0f : 1 * 1 * 0.2f;
        int min_coeffs_to_change = coeffs_to_change_per_block * blocks_to_change;

        if (first_up_iter) {
          const float limit = 0.75f * params_.butteraugli_target;
          std::vector<std::pair<int, float> >::iterator it = std::partition_point(
              global_order.begin(), global_order.end(),
              [=](const std::pair<int, float>& a) { return a.second < limit; });
          min_coeffs_to_change = std::max<int>(min_coeffs_to_change, it - global_order.begin());
          first_up_iter = false;
        }

        Clock::time_point tw = Clock::now();
        std::set<int> changed_blocks;
        float val_threshold = 0.0;
        int changed_coeffs = 0;
        int est_jpg_size = prev_size;
        edit_index.clear();
        edit_value.clear();
        for (size_t i = 0; i < global_order.size(); ++i) {
          const int block_ix = global_order[i].first;
          const int last_idx = last_indexes[block_ix];
          const uint8_t* candidates = &cand_idx[offsets[block_ix]];
          const int idx = candidates[last_idx + std::min(direction, 0)];
          const int c = idx / 64;
          const int k = idx % 64;
          const int* quant = img_.q[c];
          const int16_t* orig_block = &orig[c * per + static_cast<size_t>(block_ix) * 64];
          const int newval = direction > 0 ? 0 : quantize_coeff(orig_block[k], quant[k]);
          int16_t* block = &cand_[c * per + static_cast<size_t>(block_ix) * 64];
          ac_symbols_of_block(block, quant, -1, &ac_h[c]);
          double sum_of_hf = 0;
          for (int ii = 3; ii < 64; ++ii) {
            if ((ii & 7) < 3 && ii < 3 * 8) continue;
            sum_of_hf += std::abs(orig_block[ii]);
          }
          const int limit = sum_of_hf < 60 ? 4 : 8;
          const bool precious = (k == 1 || k == 8) && std::abs(orig_block[k]) >= limit;
          if (!precious || newval != 0) {
            block[k] = static_cast<int16_t>(newval);
            edit_index.push_back(static_cast<int>(c * per + static_cast<size_t>(block_ix) * 64 + k));
            edit_value.push_back(static_cast<int16_t>(newval));
          }
          ac_symbols_of_block(block, quant, 1, &ac_h[c]);
          last_indexes[block_ix] += direction;
          changed_blocks.insert(block_ix);
          val_threshold = global_order[i].second;
          ++changed_coeffs;
          if (i % 10 == 0) ac_histogram_size = static_cast<int>(compute_entropy_codes(ac_h, ac_depths.data()));
          est_jpg_size = header_size + dc_size + ac_histogram_size +
                         static_cast<int>(entropy_coded_bytes(ac_h, ac_depths.data()));
          if (changed_coeffs > min_coeffs_to_change && std::abs(est_jpg_size - prev_size) > min_size_delta) break;
        }
        const size_t global_order_size = global_order.size();
        std::vector<std::pair<int, float> >().swap(global_order);
        st_->ms_walk += ms_since(tw);
        (void)block_width;

        for (int i = 0; i < num_blocks; ++i) max_block_error[i] += block_weight[i] * val_threshold * direction;

        ++st_->iterations;
        if (direction > 0) ++st_->iterations_up; else ++st_->iterations_down;
        ctx_->scatter_coeffs(edit_index, edit_value);
        std::string encoded = timed_write();
        logf("Iter %2d: %s(%d) %s Coeffs[%d/%zd] Blocks[%zd/%d/%d] ValThres[%.4f] Out[%7zd] EstErr[%.2f%%]",
             st_->iterations, "f111111", 7, direction > 0 ? "up" : "down", changed_coeffs, global_order_size,
             changed_blocks.size(), blocks_to_change, num_blocks, val_threshold, encoded.size(),
             100.0 - (100.0 * est_jpg_size) / encoded.size());
        compare();
        maybe_output(encoded);
        prev_size = est_jpg_size;
      }
    }
  }

  SearchParams params_;
  ImageContext* ctx_;
  LogSink log_;
  void* log_user_;
  SearchStats* st_;
  std::string* best_;
  double best_score_;
  float distance_;
  CoeffImage img_;
  std::vector<int16_t> cand_;
};

}  // namespace

static bool check_params(const SearchParams& params, std::string* err) {
  if (params.butteraugli_target > 2.0f) {
    *err =
        "Guetzli should be called with quality >= 84, otherwise the\n"
        "output will have noticeable artifacts. If you want to\n"
        "proceed anyway, please edit the source code.\n";
    fputs(err->c_str(), stderr);
    return false;
  }
  if (params.try_420 || params.force_420 || !params.new_zeroing_model) {
    *err = "guetzli_b200: YUV420 and the legacy zeroing model are outside the B200 hot path (DESIGN.md)\n";
    fputs(err->c_str(), stderr);
    return false;
  }
  return true;
}

bool process_resident(const SearchParams& params, ImageContext* ctx, LogSink log, void* log_user,
                      std::string* jpeg_out, SearchStats* stats, std::string* err) {
  SearchStats local;
  SearchStats* st = stats ? stats : &local;
  const double setup_ms = st->ms_device_setup;
  *st = SearchStats();
  st->ms_device_setup = setup_ms;
  jpeg_out->clear();
  Clock::time_point t_all = Clock::now();
  if (!check_params(params, err)) return false;
  const long launches0 = total_launches();
  const long long h2d0 = h2d_bytes_total(), d2h0 = d2h_bytes_total();
  ctx->prepare();
  const int w = ctx->width(), h = ctx->height();
  if (w < 32 || h < 32) {
    // Butteraugli is skipped for tiny images (g/processor.cc:832-838,940)
    CoeffImage img;
    img.w = w;
    img.h = h;
    img.bw = ctx->geom().bw;
    img.bh = ctx->geom().bh;
    img.nblocks = ctx->geom().nblocks;
    img.coeffs = ctx->orig_coeffs().data();
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 64; ++k) img.q[c][k] = 1;
    img.as_encoded = true;
    *jpeg_out = write_jpeg(img);
    if (log) {
      char buf[128];
      snprintf(buf, sizeof(buf), "Original Out[%7zd] <image too small for Butteraugli>\n", jpeg_out->size());
      log(log_user, buf);
    }
  } else {
    Search search(params, ctx, log, log_user, st);
    search.run(jpeg_out);
  }
  st->gpu_launches = total_launches() - launches0;
  st->h2d_bytes = h2d_bytes_total() - h2d0;
  st->d2h_bytes = d2h_bytes_total() - d2h0;
  st->ms_total = ms_since(t_all);
  return true;
}

bool process_rgb(const SearchParams& params, const uint8_t* rgb, int w, int h, int device, LogSink log,
                 void* log_user, std::string* jpeg_out, SearchStats* stats, std::string* err) {
  SearchStats local;
  SearchStats* st = stats ? stats : &local;
  *st = SearchStats();
  jpeg_out->clear();
  if (rgb == nullptr || w < 0 || w >= 1 << 16 || h < 0 || h >= 1 << 16) {
    *err = "Could not create jpg data from rgb pixels\n";
    fputs(err->c_str(), stderr);
    return false;
  }
  if (!check_params(params, err)) return false;
  if (w == 0 || h == 0) {
    *err = "guetzli_b200: empty image\n";
    fputs(err->c_str(), stderr);
    return false;
  }
  Clock::time_point t0 = Clock::now();
  const long long h2d0 = h2d_bytes_total();
  ImageContext ctx(rgb, w, h, device, false);
  st->ms_device_setup = ms_since(t0);
  const bool ok = process_resident(params, &ctx, log, log_user, jpeg_out, st, err);
  st->h2d_bytes = h2d_bytes_total() - h2d0;
  st->ms_total = ms_since(t0);
  return ok;
}

}  // namespace gb200
