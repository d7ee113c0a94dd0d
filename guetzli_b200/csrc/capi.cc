// C ABI (include/guetzli_b200.h) over the C++ implementation.
#include "guetzli_b200.h"

#include <stdlib.h>
#include <string.h>

#include <exception>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include <algorithm>

#include "comm.h"
#include "exact_sort.h"
#include "jpeg_out.h"
#include "pipeline.h"
#include "jpeg_in.h"
#include "search.h"
#include "tables.h"

namespace gb200 {
struct ThreadGroup;
ThreadGroup* thread_group_create(int world);
void thread_group_destroy(ThreadGroup* g);
void thread_group_abort(ThreadGroup* g);
Comm* thread_comm_create(ThreadGroup* g, int rank);
#if !defined(GB200_HOSTSIM)
void nccl_unique_id(uint8_t out[128]);
Comm* nccl_comm_create(const uint8_t id[128], int rank, int world);
void select_device(int device);
void dev_trim();
#endif
long total_launches();
long long h2d_bytes_total();
long long d2h_bytes_total();
void profiling_enable(bool on);
std::vector<KernelStat> profiling_snapshot();
void profiling_reset();
#if !defined(GB200_HOSTSIM)
int cuda_device_count();
#endif
}  // namespace gb200

namespace {
thread_local std::string g_err;

template <class F>
int guarded(F f) {
  try {
    f();
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
  } catch (...) {
    g_err = "unknown error";
  }
  return 0;
}
}  // namespace

struct gb200_image {
  gb200::ImageContext* ctx;
};

namespace {
gb200::SearchParams to_search_params(const gb200_params* params) {
  gb200::SearchParams sp;
  if (params) {
    sp.butteraugli_target = params->butteraugli_target;
    sp.clear_metadata = params->clear_metadata != 0;
    sp.try_420 = params->try_420 != 0;
    sp.force_420 = params->force_420 != 0;
    sp.use_silver_screen = params->use_silver_screen != 0;
    sp.zeroing_greedy_lookahead = params->zeroing_greedy_lookahead;
    sp.new_zeroing_model = params->new_zeroing_model != 0;
  }
  return sp;
}
void fill_stats(const gb200::SearchStats& st, gb200_stats* stats) {
  if (!stats) return;
  stats->iterations = st.iterations;
  stats->iterations_up = st.iterations_up;
  stats->iterations_down = st.iterations_down;
  stats->compares = st.compares;
  stats->gpu_launches = st.gpu_launches;
  stats->h2d_bytes = st.h2d_bytes;
  stats->d2h_bytes = st.d2h_bytes;
  stats->ms_total = st.ms_total;
  stats->ms_device_setup = st.ms_device_setup;
  stats->ms_compare = st.ms_compare;
  stats->ms_zeroing = st.ms_zeroing;
  stats->ms_jpeg = st.ms_jpeg;
  stats->ms_sort = st.ms_sort;
  stats->ms_walk = st.ms_walk;
  stats->order_partial = st.order_partial;
  stats->order_exact = st.order_exact;
}
}  // namespace

extern "C" {

void gb200_params_default(gb200_params* p) {
  p->butteraugli_target = 1.0f;
  p->clear_metadata = 1;
  p->try_420 = 0;
  p->force_420 = 0;
  p->use_silver_screen = 0;
  p->zeroing_greedy_lookahead = 3;
  p->new_zeroing_model = 1;
}

double gb200_butteraugli_score_for_quality(double quality) { return gb200::distance_for_quality(quality); }

int gb200_process_rgb(const gb200_params* params, const uint8_t* rgb, int w, int h, int device,
                      gb200_log_fn log, void* log_user, uint8_t** out, size_t* out_len,
                      gb200_stats* stats) {
  *out = nullptr;
  *out_len = 0;
  bool ok = false;
  int guarded_ok = guarded([&]() {
    gb200::SearchParams sp = to_search_params(params);
    gb200::SearchStats st;
    std::string jpeg, err;
    ok = gb200::process_rgb(sp, rgb, w, h, device, log, log_user, &jpeg, &st, &err);
    if (!ok) g_err = err;
    if (!jpeg.empty()) {
      *out = static_cast<uint8_t*>(malloc(jpeg.size()));
      memcpy(*out, jpeg.data(), jpeg.size());
      *out_len = jpeg.size();
    }
    fill_stats(st, stats);
  });
  return (guarded_ok && ok) ? 1 : 0;
}

int gb200_process_jpeg(const gb200_params* params, const uint8_t* jpeg_in, size_t jpeg_len, int device,
                       gb200_log_fn log, void* log_user, uint8_t** out, size_t* out_len, gb200_stats* stats) {
  *out = nullptr;
  *out_len = 0;
  bool ok = false;
  int guarded_ok = guarded([&]() {
    gb200::SearchParams sp = to_search_params(params);
    gb200::SearchStats st;
    std::string jpeg, err;
    ok = gb200::process_jpeg(sp, jpeg_in, jpeg_len, device, log, log_user, &jpeg, &st, &err);
    if (!ok) g_err = err;
    if (!jpeg.empty()) {
      *out = static_cast<uint8_t*>(malloc(jpeg.size()));
      memcpy(*out, jpeg.data(), jpeg.size());
      *out_len = jpeg.size();
    }
    fill_stats(st, stats);
  });
  return (guarded_ok && ok) ? 1 : 0;
}

// butteraugli::ButteraugliInterface (b/butteraugli.cc:1858) on the device kernels.
int gb200_butteraugli_diffmap(const float* rgb0, const float* rgb1, int w, int h, int device, float* diffmap,
                              double* score) {
  if (rgb0 == nullptr || rgb1 == nullptr || w < 1 || h < 1) {
    g_err = "butteraugli: no image";
    return 0;
  }
  return guarded([&]() {
    // images below 8 pixels in a dimension are edge-replicated to 8 (b/butteraugli.cc:1825-1853)
    const int kMin = 8;
    const int xb = w < kMin ? (kMin - w) / 2 : 0, yb = h < kMin ? (kMin - h) / 2 : 0;
    const int ws = std::max(w, kMin), hs = std::max(h, kMin);
    std::vector<float> s0, s1;
    const float* p0 = rgb0;
    const float* p1 = rgb1;
    if (ws != w || hs != h) {
      s0.resize(static_cast<size_t>(3) * ws * hs);
      s1.resize(s0.size());
      for (int c = 0; c < 3; ++c)
        for (int y = 0; y < hs; ++y)
          for (int x = 0; x < ws; ++x) {
            const int x2 = std::min(w - 1, std::max(0, x - xb)), y2 = std::min(h - 1, std::max(0, y - yb));
            const size_t d = (static_cast<size_t>(c) * hs + y) * ws + x, s = (static_cast<size_t>(c) * h + y2) * w + x2;
            s0[d] = rgb0[s];
            s1[d] = rgb1[s];
          }
      p0 = s0.data();
      p1 = s1.data();
    }
    gb200::ImageContext ctx(p0, ws, hs, device);
    const float dmax = ctx.compare_linear(p1);
    std::vector<float> dm(static_cast<size_t>(ws) * hs);
    if (diffmap != nullptr || ws != w || hs != h) ctx.download_distmap(dm.data());
    float m = 0.0f;
    if (ws != w || hs != h) {
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          const float v = dm[static_cast<size_t>(y + yb) * ws + x + xb];
          if (diffmap) diffmap[static_cast<size_t>(y) * w + x] = v;
          m = std::max(m, v);
        }
    } else {
      if (diffmap) memcpy(diffmap, dm.data(), dm.size() * sizeof(float));
      m = dmax;
    }
    if (score) *score = m;
  });
}

int gb200_jpeg_dimensions(const uint8_t* jpeg_in, size_t jpeg_len, int* width, int* height) {
  return gb200::read_jpeg_dimensions(jpeg_in, jpeg_len, width, height) ? 1 : 0;
}

int gb200_debug_read_jpeg(const uint8_t* jpeg_in, size_t jpeg_len, int* dims, int16_t* out, size_t out_cap) {
  gb200::JpegInput jpg;
  std::string err;
  if (!gb200::read_jpeg(jpeg_in, jpeg_len, &jpg, &err)) {
    g_err = err;
    return 0;
  }
  dims[0] = jpg.width;
  dims[1] = jpg.height;
  dims[2] = static_cast<int>(jpg.components.size());
  size_t pos = 0;
  for (size_t c = 0; c < jpg.components.size(); ++c) {
    dims[3 + 2 * c] = jpg.components[c].width_in_blocks;
    dims[4 + 2 * c] = jpg.components[c].height_in_blocks;
    for (size_t i = 0; i < jpg.components[c].coeffs.size(); ++i, ++pos)
      if (pos < out_cap) out[pos] = jpg.components[c].coeffs[i];
  }
  return pos <= out_cap ? 1 : 0;
}

int gb200_image_process(gb200_image* img, const gb200_params* params, gb200_log_fn log, void* log_user,
                        uint8_t** out, size_t* out_len, gb200_stats* stats) {
  *out = nullptr;
  *out_len = 0;
  bool ok = false;
  int guarded_ok = guarded([&]() {
    gb200::SearchParams sp = to_search_params(params);
    gb200::SearchStats st;
    std::string jpeg, err;
    ok = gb200::process_resident(sp, img->ctx, log, log_user, &jpeg, &st, &err);
    if (!ok) g_err = err;
    if (!jpeg.empty()) {
      *out = static_cast<uint8_t*>(malloc(jpeg.size()));
      memcpy(*out, jpeg.data(), jpeg.size());
      *out_len = jpeg.size();
    }
    fill_stats(st, stats);
  });
  return (guarded_ok && ok) ? 1 : 0;
}

// ---- row-strip mode -------------------------------------------------------
namespace {
gb200::Comm* g_comm = nullptr;
int g_comm_device = 0;
}  // namespace

int gb200_dist_unique_id(uint8_t* out128) {
#if defined(GB200_HOSTSIM)
  (void)out128;
  g_err = "the CPU port has no NCCL";
  return 0;
#else
  return guarded([&]() { gb200::nccl_unique_id(out128); });
#endif
}

int gb200_dist_init(const uint8_t* id128, int rank, int world, int device) {
#if defined(GB200_HOSTSIM)
  (void)id128; (void)rank; (void)world; (void)device;
  g_err = "the CPU port has no NCCL";
  return 0;
#else
  return guarded([&]() {
    gb200::select_device(device);
    delete g_comm;
    g_comm = gb200::nccl_comm_create(id128, rank, world);
    g_comm_device = device;
  });
#endif
}

void gb200_dist_shutdown(void) {
  guarded([&]() {
    delete g_comm;
    g_comm = nullptr;
  });
}

int gb200_process_rgb_tiled(const gb200_params* params, const uint8_t* rgb, int w, int h, gb200_log_fn log,
                            void* log_user, uint8_t** out, size_t* out_len, gb200_stats* stats) {
  *out = nullptr;
  *out_len = 0;
  bool ok = false;
  int guarded_ok = guarded([&]() {
    if (!g_comm) throw std::runtime_error("gb200_dist_init has not been called");
    gb200::SearchParams sp = to_search_params(params);
    gb200::SearchStats st;
    std::string jpeg, err;
    ok = gb200::process_rgb_tiled(sp, rgb, w, h, g_comm_device, g_comm, log, log_user, &jpeg, &st, &err);
    if (!ok) g_err = err;
    if (!jpeg.empty()) {
      *out = static_cast<uint8_t*>(malloc(jpeg.size()));
      memcpy(*out, jpeg.data(), jpeg.size());
      *out_len = jpeg.size();
    }
    fill_stats(st, stats);
  });
  return (guarded_ok && ok) ? 1 : 0;
}

// The same strip decomposition with `world` host threads of this process sharing one
// device (test entry: exercises the strip kernels and the exchange without NCCL).
int gb200_process_rgb_tiled_threads(const gb200_params* params, const uint8_t* rgb, int w, int h, int device,
                                    int world, uint8_t** out, size_t* out_len, gb200_stats* stats) {
  *out = nullptr;
  *out_len = 0;
  bool all_ok = true;
  int guarded_ok = guarded([&]() {
    if (world < 1 || world > 64) throw std::runtime_error("bad world size");
    gb200::ThreadGroup* group = gb200::thread_group_create(world);
    std::vector<std::string> jpegs(world), errs(world);
    std::vector<gb200::SearchStats> sts(world);
    std::vector<int> oks(world, 0);
    std::vector<std::string> what(world);
    std::vector<std::thread> threads;
    gb200::SearchParams sp = to_search_params(params);
    for (int r = 0; r < world; ++r) {
      threads.emplace_back([&, r]() {
        gb200::Comm* comm = gb200::thread_comm_create(group, r);
        try {
          oks[r] = gb200::process_rgb_tiled(sp, rgb, w, h, device, comm, nullptr, nullptr, &jpegs[r], &sts[r], &errs[r]);
        } catch (const std::exception& e) {
          what[r] = e.what();
          oks[r] = -1;
          gb200::thread_group_abort(group);  // the other ranks leave their barriers with an error
        }
        delete comm;
      });
    }
    for (auto& t : threads) t.join();
    gb200::thread_group_destroy(group);
    for (int r = 0; r < world; ++r) {
      if (oks[r] < 0) throw std::runtime_error("rank " + std::to_string(r) + ": " + what[r]);
      if (!oks[r]) {
        all_ok = false;
        g_err = errs[r];
      }
      if (jpegs[r] != jpegs[0]) throw std::runtime_error("ranks disagree on the output");
    }
    if (!jpegs[0].empty()) {
      *out = static_cast<uint8_t*>(malloc(jpegs[0].size()));
      memcpy(*out, jpegs[0].data(), jpegs[0].size());
      *out_len = jpegs[0].size();
    }
    fill_stats(sts[0], stats);
  });
  return (guarded_ok && all_ok) ? 1 : 0;
}

void gb200_free(void* p) { free(p); }
const char* gb200_last_error(void) { return g_err.c_str(); }
const char* gb200_backend_name(void) { return gb200::backend_name(); }

int gb200_device_count(void) {
#if defined(GB200_HOSTSIM)
  return 1;
#else
  return gb200::cuda_device_count();
#endif
}

gb200_image* gb200_image_create(const uint8_t* rgb, int w, int h, int device) {
  return gb200_image_create2(rgb, w, h, device, 1);
}

gb200_image* gb200_image_create2(const uint8_t* rgb, int w, int h, int device, int prepare) {
  gb200_image* img = nullptr;
  guarded([&]() {
    if (!rgb || w <= 0 || h <= 0 || w >= 65536 || h >= 65536) throw std::runtime_error("bad image size");
    std::string why;
    if (!gb200::image_size_supported(w, h, &why)) throw std::runtime_error(why);
    gb200::ImageContext* ctx = new gb200::ImageContext(rgb, w, h, device, prepare != 0);
    img = new gb200_image;
    img->ctx = ctx;
  });
  return img;
}

void gb200_image_destroy(gb200_image* img) {
  if (!img) return;
  guarded([&]() { img->ctx->bind(); delete img->ctx; });
  delete img;
}

int gb200_image_reset(gb200_image* img) {
  return guarded([&]() { img->ctx->reset_prepared(); });
}

int gb200_image_num_blocks(const gb200_image* img) { return img->ctx->geom().nblocks; }

int gb200_image_orig_coeffs(gb200_image* img, int16_t* out) {
  return guarded([&]() {
    img->ctx->bind();
    const std::vector<int16_t>& c = img->ctx->orig_coeffs();
    memcpy(out, c.data(), c.size() * sizeof(int16_t));
  });
}

int gb200_image_apply_global_quant(gb200_image* img, const int* q) {
  return guarded([&]() { img->ctx->bind(); img->ctx->apply_global_quant(q); });
}
int gb200_image_upload_candidate(gb200_image* img, const int16_t* coeffs) {
  return guarded([&]() { img->ctx->bind(); img->ctx->upload_candidate(coeffs); });
}
int gb200_image_download_candidate(gb200_image* img, int16_t* coeffs) {
  return guarded([&]() { img->ctx->bind(); img->ctx->download_candidate(coeffs); });
}
int gb200_image_scatter(gb200_image* img, const int* index, const int16_t* value, int n) {
  return guarded([&]() {
    img->ctx->bind();
    img->ctx->scatter_coeffs(std::vector<int>(index, index + n), std::vector<int16_t>(value, value + n));
  });
}
int gb200_image_compare(gb200_image* img, float* distance) {
  return guarded([&]() { *distance = img->ctx->compare(); });
}
int gb200_image_distmap(gb200_image* img, float* out) {
  return guarded([&]() { img->ctx->bind(); img->ctx->download_distmap(out); });
}
int gb200_image_block_weights(gb200_image* img, int direction, int radius, double target_distance,
                              int zero_distmap, float* out) {
  return guarded([&]() { img->ctx->bind(); img->ctx->block_weights(direction, radius, target_distance, zero_distmap != 0, out); });
}
int gb200_image_zeroing_orders(gb200_image* img, float block_error_limit, int lookahead, uint8_t* idx,
                               float* err, int* count) {
  return guarded([&]() {
    std::vector<uint8_t> vi;
    std::vector<float> ve;
    std::vector<int> vc;
    img->ctx->bind();
    img->ctx->zeroing_orders(block_error_limit, lookahead, true, &vi, &ve, &vc);
    memcpy(idx, vi.data(), vi.size());
    memcpy(err, ve.data(), ve.size() * sizeof(float));
    memcpy(count, vc.data(), vc.size() * sizeof(int));
  });
}
int gb200_image_debug_blur(gb200_image* img, const float* in, float* out, int blur_id) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_blur(in, out, blur_id); });
}
int gb200_image_debug_opsin(gb200_image* img, const float* rgb_linear, float* xyb) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_opsin(rgb_linear, xyb); });
}
int gb200_image_debug_separate(gb200_image* img, const float* xyb, float* psycho10) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_separate(xyb, psycho10); });
}
int gb200_image_debug_render(gb200_image* img, float* linear_rgb) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_render(linear_rgb); });
}
int gb200_image_debug_psycho0(gb200_image* img, float* psycho10) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_psycho0(psycho10); });
}
int gb200_image_debug_corner_mask(gb200_image* img, float* out) {
  return guarded([&]() { img->ctx->bind(); img->ctx->debug_corner_mask(out); });
}

int gb200_write_jpeg(const int16_t* coeffs, int w, int h, const int* q, uint8_t** out, size_t* out_len) {
  return guarded([&]() {
    gb200::CoeffImage ci;
    ci.w = w;
    ci.h = h;
    ci.bw = (w + 7) / 8;
    ci.bh = (h + 7) / 8;
    ci.nblocks = ci.bw * ci.bh;
    ci.coeffs = coeffs;
    memcpy(ci.q, q, sizeof(ci.q));
    std::string s = gb200::write_jpeg(ci);
    *out = static_cast<uint8_t*>(malloc(s.size() + 1));
    memcpy(*out, s.data(), s.size());
    *out_len = s.size();
  });
}

int gb200_image_save_jpeg(gb200_image* img, const int* q, uint8_t** out, size_t* out_len) {
  return guarded([&]() {
    img->ctx->bind();
    std::string s;
    gb200::device_save_jpeg(img->ctx, q, &s);
    *out = static_cast<uint8_t*>(malloc(s.size() + 1));
    memcpy(*out, s.data(), s.size());
    *out_len = s.size();
  });
}

// test hooks for the prefix-exact std::sort replay (exact_sort.h)
size_t gb200_debug_partial_sort(int* block, float* key, size_t n, size_t want) {
  std::vector<gb200::exact_sort::Item> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(block[i], key[i]);
  const size_t k = gb200::exact_sort::partial_std_sort(v.data(), n, want);
  for (size_t i = 0; i < n; ++i) {
    block[i] = v[i].first;
    key[i] = v[i].second;
  }
  return k;
}
// the same replay with the large partition passes on the device (order_exact.h); needs an image
// context only for its stream and scratch
size_t gb200_debug_device_partial_sort(gb200_image* img, int* block, float* key, size_t n, size_t want) {
  size_t k = 0;
  guarded([&]() {
    std::vector<gb200::exact_sort::Item> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(block[i], key[i]);
    k = img->ctx->debug_device_partial_sort(v.data(), n, want);
    for (size_t i = 0; i < k; ++i) {
      block[i] = v[i].first;
      key[i] = v[i].second;
    }
  });
  return k;
}
void gb200_debug_std_sort(int* block, float* key, size_t n) {
  std::vector<std::pair<int, float> > v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(block[i], key[i]);
  std::sort(v.begin(), v.end(),
            [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second < b.second; });
  for (size_t i = 0; i < n; ++i) {
    block[i] = v[i].first;
    key[i] = v[i].second;
  }
}

void gb200_debug_huffman_depths(const uint32_t* counts, int n, int limit, uint8_t* depth) {
  gb200::huffman_code_lengths(counts, n, limit, depth);
}

void gb200_trim_memory(void) {
#if !defined(GB200_HOSTSIM)
  guarded([&]() { gb200::dev_trim(); });
#endif
}

void gb200_counters(long* launches, long long* h2d_bytes, long long* d2h_bytes) {
  if (launches) *launches = gb200::total_launches();
  if (h2d_bytes) *h2d_bytes = gb200::h2d_bytes_total();
  if (d2h_bytes) *d2h_bytes = gb200::d2h_bytes_total();
}

void gb200_profile_enable(int on) { gb200::profiling_enable(on != 0); }
void gb200_profile_reset(void) { gb200::profiling_reset(); }
int gb200_profile_get(char (*names)[48], long* launches, double* ms, double* elements, int cap) {
  std::vector<gb200::KernelStat> s = gb200::profiling_snapshot();
  const int n = static_cast<int>(s.size());
  for (int i = 0; i < n && i < cap; ++i) {
    strncpy(names[i], s[i].name.c_str(), 47);
    names[i][47] = 0;
    launches[i] = s[i].launches;
    ms[i] = s[i].ms;
    elements[i] = s[i].elements;
  }
  return n;
}

}  // extern "C"
