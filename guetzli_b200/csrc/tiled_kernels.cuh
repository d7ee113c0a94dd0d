// Hand-tiled sm_100a kernels for the stages that dominate Compare (a10).  They
// compute exactly what the generic per-pixel functors in kernels.h compute (same
// helper arithmetic from ba_math.h, same tap order) but stage tiles in shared
// memory and keep partial results in registers instead of round-tripping planes
// through HBM.  CUDA only; the CPU port keeps the generic functors.
#pragma once
#include <cuda_runtime.h>

#include "ba_math.h"
#include "kernels.h"
#include "malta_unrolled.inc"

namespace gb200 {

// ---------------------------------------------------------------------------
// S7 Malta, one colour channel per launch (b/butteraugli.cc:829-871,1461-1568).
// For each of the three bands (uhf: 9-tap lines; hf, mf: 5-tap lines) the CTA
// builds the "diffs" tile (pre-pass, malta_diff) with a 4-pixel zero-padded halo
// in shared memory, every thread evaluates the 16 line sums of its two pixels from
// shared memory, and the three results are accumulated in registers in the
// reference's call order: ac = ((0 + uhf) + hf) + mf.   Replaces 3x(malta_pre +
// malta_acc): 6 plane reads + 1 write per pixel instead of 9 reads + 6 writes.
struct MaltaChannelArgs {
  const float* lum0[3];  // original:  uhf, hf, mf plane of this channel
  const float* lum1[3];  // candidate
  MaltaParams mp[3];
  float* acc;            // block_diff_ac plane of this channel (overwritten)
  Geom g;
};

#define GB_MALTA_TILE_W 32
#define GB_MALTA_TILE_H 16
#define GB_MALTA_SW (GB_MALTA_TILE_W + 8)
#define GB_MALTA_SH (GB_MALTA_TILE_H + 8)

__global__ void __launch_bounds__(256) k_malta_channel(MaltaChannelArgs a) {
  __shared__ float tile[GB_MALTA_SH * GB_MALTA_SW];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int x0 = blockIdx.x * GB_MALTA_TILE_W, y0 = blockIdx.y * GB_MALTA_TILE_H;
  const int tid = ty * 32 + tx;
  float r0 = 0.0f, r1 = 0.0f;
#pragma unroll 1
  for (int band = 0; band < 3; ++band) {
    const float* l0 = a.lum0[band];
    const float* l1 = a.lum1[band];
    const MaltaParams mp = a.mp[band];
    __syncthreads();
    for (int i = tid; i < GB_MALTA_SH * GB_MALTA_SW; i += 256) {
      const int sy = i / GB_MALTA_SW, sx = i - sy * GB_MALTA_SW;
      const int x = x0 + sx - 4, y = y0 + sy - 4;
      float v = 0.0f;
      if (x >= 0 && x < a.g.w && y >= 0 && y < a.g.h) {
        const size_t o = static_cast<size_t>(y) * a.g.pitch + x;
        v = malta_diff(l0[o], l1[o], mp);
      }
      tile[i] = v;
    }
    __syncthreads();
    float u0 = 0.0f, u1 = 0.0f;
    const float* c0 = tile + (ty + 4) * GB_MALTA_SW + tx + 4;
    const float* c1 = c0 + 8 * GB_MALTA_SW;
#define GB_T0(dx, dy) c0[(dy) * GB_MALTA_SW + (dx)]
#define GB_T1(dx, dy) c1[(dy) * GB_MALTA_SW + (dx)]
    if (band == 0) {
      GB_MALTA_HF_SUMS(GB_T0, u0)
      GB_MALTA_HF_SUMS(GB_T1, u1)
    } else {
      GB_MALTA_LF_SUMS(GB_T0, u0)
      GB_MALTA_LF_SUMS(GB_T1, u1)
    }
#undef GB_T0
#undef GB_T1
    r0 = r0 + u0;
    r1 = r1 + u1;
  }
  const int x = x0 + tx;
  if (x < a.g.w) {
    const int ya = y0 + ty, yb = y0 + ty + 8;
    if (ya < a.g.h) a.acc[static_cast<size_t>(ya) * a.g.pitch + x] = r0;
    if (yb < a.g.h) a.acc[static_cast<size_t>(yb) * a.g.pitch + x] = r1;
  }
}

inline void launch_malta_channel(Stream s, const MaltaChannelArgs& a) {
  dim3 block(32, 8), grid((a.g.w + GB_MALTA_TILE_W - 1) / GB_MALTA_TILE_W, (a.g.h + GB_MALTA_TILE_H - 1) / GB_MALTA_TILE_H);
  note_launch("malta_channel", s, static_cast<double>(a.g.w) * a.g.h);
  k_malta_channel<<<grid, block, 0, s>>>(a);
  note_launch_end("malta_channel", s);
}

// ---------------------------------------------------------------------------
// Separable blur (b/butteraugli.cc:184-233) with register blocking.
//
// x pass: a CTA stages [8 rows][128 + 2r] input samples in shared memory; each
// thread produces 4 adjacent outputs of one row, streaming the 2r+4 samples it
// needs once and adding every product to the right accumulator in ascending tap
// order (per output the sequence of float additions is the reference's).
// Border outputs (x < r or x + r >= w) take the raw-tap / scale path.
#define GB_BLUR_MAX_R 24
#define GB_BLURX_TW 128

struct BlurArgs {
  const float* in;
  float* out;
  BlurTab tab;
  Geom g;
  int rows;  // total rows = nplanes * h
};

__global__ void __launch_bounds__(256) k_blur_x(BlurArgs a) {
  __shared__ float tile[8][GB_BLURX_TW + 2 * GB_BLUR_MAX_R];
  __shared__ float taps_n[2 * GB_BLUR_MAX_R + 1];
  __shared__ float taps[2 * GB_BLUR_MAX_R + 1];
  const int r = a.tab.r;
  const int len = 2 * r + 1;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < len) {
    taps_n[tid] = a.tab.taps_n[tid];
    taps[tid] = a.tab.taps[tid];
  }
  const int x0 = blockIdx.x * GB_BLURX_TW;
  const int row0 = blockIdx.y * 8;
  const int span = GB_BLURX_TW + 2 * r;
  for (int i = tid; i < 8 * span; i += 256) {
    const int ry = i / span, sx = i - ry * span;
    const int x = x0 - r + sx, row = row0 + ry;
    float v = 0.0f;
    if (row < a.rows && x >= 0 && x < a.g.w) v = a.in[static_cast<size_t>(row) * a.g.pitch + x];
    tile[ry][sx] = v;
  }
  __syncthreads();
  const int ry = threadIdx.y;
  const int row = row0 + ry;
  if (row >= a.rows) return;
  const int xb = x0 + threadIdx.x * 4;  // first of 4 outputs
  if (xb >= a.g.w) return;
  const float* srow = tile[ry];
  float* orow = a.out + static_cast<size_t>(row) * a.g.pitch;
  const int w = a.g.w;
  const bool interior = (xb >= r) && (xb + 3 + r < w);
  if (interior) {
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
    const int base = threadIdx.x * 4;  // tile index of tap 0 of output 0
    // sample t (0 .. 2r+3) contributes tap (t - o) to output o
    for (int t = 0; t < len + 3; ++t) {
      const float v = srow[base + t];
      if (t < len) acc0 += v * taps_n[t];
      if (t >= 1 && t - 1 < len) acc1 += v * taps_n[t - 1];
      if (t >= 2 && t - 2 < len) acc2 += v * taps_n[t - 2];
      if (t >= 3) acc3 += v * taps_n[t - 3];
    }
    orow[xb] = acc0;
    orow[xb + 1] = acc1;
    orow[xb + 2] = acc2;
    orow[xb + 3] = acc3;
  } else {
    for (int o = 0; o < 4; ++o) {
      const int x = xb + o;
      if (x >= w) break;
      float sum = 0.0f;
      if (x < r || x + r >= w) {
        const int lo = x < r ? 0 : x - r;
        const int hi = (x + r < w - 1) ? x + r : w - 1;
        for (int j = lo; j <= hi; ++j) sum += srow[j - x0 + r] * taps[j - x + r];
        sum = sum * a.tab.scale_x[x];
      } else {
        for (int j = 0; j < len; ++j) sum += srow[x - x0 + j] * taps_n[j];
      }
      orow[x] = sum;
    }
  }
}

// y pass: each thread owns one column x and produces 8 consecutive rows, streaming
// the 8 + 2r input rows once (coalesced across the warp) into 8 accumulators.
#define GB_BLURY_R 8
__global__ void __launch_bounds__(128) k_blur_y(BlurArgs a) {
  __shared__ float taps_n[2 * GB_BLUR_MAX_R + 1];
  __shared__ float taps[2 * GB_BLUR_MAX_R + 1];
  const int r = a.tab.r;
  const int len = 2 * r + 1;
  if (threadIdx.x < len) {
    taps_n[threadIdx.x] = a.tab.taps_n[threadIdx.x];
    taps[threadIdx.x] = a.tab.taps[threadIdx.x];
  }
  __syncthreads();
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= a.g.w) return;
  const int h = a.g.h;
  const int strips = (h + GB_BLURY_R - 1) / GB_BLURY_R;
  const int pl = blockIdx.y / strips;
  const int yb = (blockIdx.y - pl * strips) * GB_BLURY_R;
  const float* col = a.in + static_cast<size_t>(pl) * a.g.plane + x;
  float* ocol = a.out + static_cast<size_t>(pl) * a.g.plane + x;
  const size_t pitch = a.g.pitch;
  const bool interior = (yb >= r) && (yb + GB_BLURY_R - 1 + r < h);
  if (interior) {
    float acc[GB_BLURY_R];
#pragma unroll
    for (int o = 0; o < GB_BLURY_R; ++o) acc[o] = 0.0f;
    const int ystart = yb - r;
    for (int t = 0; t < len + GB_BLURY_R - 1; ++t) {
      const float v = col[static_cast<size_t>(ystart + t) * pitch];
#pragma unroll
      for (int o = 0; o < GB_BLURY_R; ++o) {
        const int j = t - o;
        if (j >= 0 && j < len) acc[o] += v * taps_n[j];
      }
    }
#pragma unroll
    for (int o = 0; o < GB_BLURY_R; ++o) ocol[static_cast<size_t>(yb + o) * pitch] = acc[o];
  } else {
    for (int o = 0; o < GB_BLURY_R; ++o) {
      const int y = yb + o;
      if (y >= h) break;
      float sum = 0.0f;
      if (y < r || y + r >= h) {
        const int lo = y < r ? 0 : y - r;
        const int hi = (y + r < h - 1) ? y + r : h - 1;
        for (int j = lo; j <= hi; ++j) sum += col[static_cast<size_t>(j) * pitch] * taps[j - y + r];
        sum = sum * a.tab.scale_y[y];
      } else {
        for (int j = 0; j < len; ++j) sum += col[static_cast<size_t>(y - r + j) * pitch] * taps_n[j];
      }
      ocol[static_cast<size_t>(y) * pitch] = sum;
    }
  }
}

inline void launch_blur_tiled(Stream s, const float* in, float* tmp, float* out, int nplanes, const BlurTab& tab,
                              const Geom& g) {
  BlurArgs ax{in, tmp, tab, g, nplanes * g.h};
  dim3 bx(32, 8), gx((g.w + GB_BLURX_TW - 1) / GB_BLURX_TW, (nplanes * g.h + 7) / 8);
  note_launch("blur_x", s, static_cast<double>(g.w) * g.h * nplanes);
  k_blur_x<<<gx, bx, 0, s>>>(ax);
  note_launch_end("blur_x", s);
  BlurArgs ay{tmp, out, tab, g, nplanes * g.h};
  const int strips = (g.h + GB_BLURY_R - 1) / GB_BLURY_R;
  dim3 gy((g.w + 127) / 128, strips * nplanes);
  note_launch("blur_y", s, static_cast<double>(g.w) * g.h * nplanes);
  k_blur_y<<<gy, 128, 0, s>>>(ay);
  note_launch_end("blur_y", s);
}

// ---------------------------------------------------------------------------
// Parallel form of OrderSelectBin (kernels.h): one CTA of 1024 threads, 64 bins
// per thread, block-wide scan of the partial sums, then the owning thread walks its
// 64 bins.  Same result as the serial functor.
__global__ void __launch_bounds__(1024) k_order_select_bin(const unsigned int* hist, OrderSelectState* st, int level) {
  __shared__ unsigned int part[1024];
  __shared__ unsigned int warp_tot[32];
  const int t = threadIdx.x;
  unsigned int local = 0;
  for (int i = 0; i < 64; ++i) local += hist[t * 64 + i];
  unsigned int incl = local;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned int base = 0;
  for (int k = 0; k < warp; ++k) base += warp_tot[k];
  const unsigned int excl = base + incl - local;  // entries in bins before this thread's range
  part[t] = excl;
  __syncthreads();
  unsigned int total = 0;
  for (int k = 0; k < 32; ++k) total += warp_tot[k];
  unsigned int want;
  if (level == 0) {
    want = st->want;
  } else {
    want = st->want > st->below_hi ? st->want - st->below_hi : 0;
  }
  // the owner is the thread whose range first reaches `want` (cum >= want)
  const bool reaches = excl + local >= want;
  const bool prev_reaches = t > 0 ? (excl >= want) : false;
  // level 1 with want == 0 selects bin 0 of thread 0 like the serial loop (cum >= 0 at i = 0)
  bool owner = reaches && !prev_reaches;
  if (level == 0 && want == 0) owner = false;  // serial loop: found only if cum + h >= want at i=0 -> bin 0
  __syncthreads();
  if (level == 0) {
    if (want == 0) {
      if (t == 0) {
        st->hi_bin = 0;
        st->below_hi = 0;
        st->total = total;
      }
      return;
    }
    if (total < want) {
      if (t == 0) {
        st->hi_bin = 65535;
        st->below_hi = total - hist[65535];
        st->total = total;
      }
      return;
    }
    if (owner) {
      unsigned int cum = excl;
      for (int i = 0; i < 64; ++i) {
        const unsigned int h = hist[t * 64 + i];
        if (cum + h >= want) {
          st->hi_bin = t * 64 + i;
          st->below_hi = cum;
          break;
        }
        cum += h;
      }
      st->total = total;
    }
  } else {
    if (total < want) {
      if (t == 0) {
        st->threshold = (st->hi_bin << 16) | 65535u;
        st->kept = st->below_hi + total;
        st->counter = 0;
      }
      return;
    }
    if (want == 0) {
      if (t == 0) {
        st->threshold = (st->hi_bin << 16) | 0u;
        st->kept = st->below_hi + hist[0];
        st->counter = 0;
      }
      return;
    }
    if (owner) {
      unsigned int cum = excl;
      for (int i = 0; i < 64; ++i) {
        cum += hist[t * 64 + i];
        if (cum >= want) {
          st->threshold = (st->hi_bin << 16) | static_cast<unsigned int>(t * 64 + i);
          st->kept = st->below_hi + cum;
          st->counter = 0;
          break;
        }
      }
    }
  }
}

inline void launch_order_select_bin(Stream s, const unsigned int* hist, OrderSelectState* st, int level) {
  note_launch("order_select_bin", s, 65536);
  k_order_select_bin<<<1, 1024, 0, s>>>(hist, st, level);
  note_launch_end("order_select_bin", s);
}

}  // namespace gb200
