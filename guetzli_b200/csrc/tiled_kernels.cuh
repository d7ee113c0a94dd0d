// Hand-tiled sm_100a kernels for the stages that dominate Compare (a10).  They
// compute exactly what the generic per-pixel functors in kernels.h compute (same
// helper arithmetic from ba_math.h, same tap order) but stage tiles in shared
// memory and keep partial results in registers instead of round-tripping planes
// through HBM.  CUDA only; the CPU port keeps the generic functors.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>

#include "ba_math.h"
#include "jpeg_dev.h"
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
  int y0, nrows;         // rows [y0, y0 + nrows) are produced
};

#define GB_MALTA_TILE_W 32
#define GB_MALTA_TILE_H 16
#define GB_MALTA_SW (GB_MALTA_TILE_W + 8)
#define GB_MALTA_SH (GB_MALTA_TILE_H + 8)

__global__ void __launch_bounds__(256) k_malta_channel(MaltaChannelArgs a) {
  __shared__ float tile[GB_MALTA_SH * GB_MALTA_SW];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int x0 = blockIdx.x * GB_MALTA_TILE_W, y0 = a.y0 + blockIdx.y * GB_MALTA_TILE_H;
  const int y_end = a.y0 + a.nrows < a.g.h ? a.y0 + a.nrows : a.g.h;
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
    if (ya < y_end) a.acc[static_cast<size_t>(ya) * a.g.pitch + x] = r0;
    if (yb < y_end) a.acc[static_cast<size_t>(yb) * a.g.pitch + x] = r1;
  }
}

inline void launch_malta_channel(Stream s, const MaltaChannelArgs& a) {
  dim3 block(32, 8), grid((a.g.w + GB_MALTA_TILE_W - 1) / GB_MALTA_TILE_W, (a.nrows + GB_MALTA_TILE_H - 1) / GB_MALTA_TILE_H);
  note_launch("malta_channel", s, static_cast<double>(a.g.w) * a.nrows);
  k_malta_channel<<<grid, block, 0, s>>>(a);
  note_launch_end("malta_channel", s);
}

// ---------------------------------------------------------------------------
// Separable blur (b/butteraugli.cc:184-233), register-blocked, radius known at
// compile time.  The interior taps travel as a kernel argument, i.e. they sit in
// the constant bank and every multiply takes its tap as an immediate constant
// operand (fully unrolled loops): per multiply-add the SM issues one FMUL and one
// FADD (no FMA: bit-exactness) and per input sample one load.
// For every output the products are added in ascending tap order, exactly the
// reference's sequence.  Border outputs (p < r or p + r >= n) take the raw-tap /
// per-position-scale path of ConvolveBorderColumn.
#define GB_BLUR_MAX_R 24

struct BlurArgs {
  const float* in;
  float* out;
  BlurTab tab;
  Geom g;
  int rows;   // rows to produce = nplanes * nrows
  int y0;     // first row of every plane
  int nrows;  // rows per plane
};

template <int R>
struct BlurTaps {
  float n[2 * R + 1];  // taps * (1/sum)
};

// x pass: a CTA stages [8 rows][256 + 2R] samples in shared memory; a thread makes
// 8 adjacent outputs of one row.
#define GB_BLURX_TW 256
#define GB_BLURX_PT 8
template <int R>
__global__ void __launch_bounds__(256) k_blur_x(BlurArgs a, BlurTaps<R> taps) {
  constexpr int LEN = 2 * R + 1;
  constexpr int SPAN = GB_BLURX_TW + 2 * R;
  __shared__ float tile[8][SPAN + 1];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * GB_BLURX_TW;
  const int row0 = blockIdx.y * 8;
  for (int i = tid; i < 8 * SPAN; i += 256) {
    const int ry = i / SPAN, sx = i - ry * SPAN;
    const int x = x0 - R + sx, row = row0 + ry;
    float v = 0.0f;
    if (row < a.rows && x >= 0 && x < a.g.w) {
      const int pl = row / a.nrows;
      const size_t grow = static_cast<size_t>(pl) * a.g.h + a.y0 + (row - pl * a.nrows);
      v = a.in[grow * a.g.pitch + x];
    }
    tile[ry][sx] = v;
  }
  __syncthreads();
  const int ry = threadIdx.y;
  const int row = row0 + ry;
  if (row >= a.rows) return;
  const int w = a.g.w;
  const float* srow = tile[ry];
  const int opl = row / a.nrows;
  float* orow = a.out + (static_cast<size_t>(opl) * a.g.h + a.y0 + (row - opl * a.nrows)) * a.g.pitch;
  // thread handles x = x0 + threadIdx.x + 32*o (o < 8): a warp reads consecutive
  // shared-memory words (no bank conflicts).  Eight independent accumulation chains
  // per thread, each adding its products in ascending tap order; outputs that need the
  // border rule (only in the first / last tile of a row) are recomputed afterwards.
  float acc[GB_BLURX_PT];
#pragma unroll
  for (int o = 0; o < GB_BLURX_PT; ++o) acc[o] = 0.0f;
#pragma unroll
  for (int j = 0; j < LEN; ++j) {
    const float tap = taps.n[j];
#pragma unroll
    for (int o = 0; o < GB_BLURX_PT; ++o) acc[o] += srow[threadIdx.x + 32 * o + j] * tap;
  }
  if (x0 >= R && x0 + GB_BLURX_TW - 1 + R < w) {  // whole tile interior (all but the first / last tile of a row)
#pragma unroll
    for (int o = 0; o < GB_BLURX_PT; ++o) orow[x0 + threadIdx.x + 32 * o] = acc[o];
    return;
  }
#pragma unroll
  for (int o = 0; o < GB_BLURX_PT; ++o) {
    const int x = x0 + threadIdx.x + 32 * o;
    if (x < w && x >= R && x + R < w) orow[x] = acc[o];
  }
  // border rule (blur_tap_sum's clamped window and per-column scale) for the few outputs that need it
#pragma unroll 1
  for (int o = 0; o < GB_BLURX_PT; ++o) {
    const int x = x0 + threadIdx.x + 32 * o;
    if (x >= w || (x >= R && x + R < w)) continue;
    const int lo = x < R ? 0 : x - R;
    const int hi = (x + R < w - 1) ? x + R : w - 1;
    float sum = 0.0f;
#pragma unroll 1
    for (int j = lo; j <= hi; ++j) sum += srow[j - x0 + R] * a.tab.taps[j - x + R];
    orow[x] = sum * a.tab.scale_x[x];
  }
}

// y pass: a thread owns one column and makes GB_BLURY_R consecutive rows, streaming the
// GB_BLURY_R + 2R input rows once (coalesced across the warp) into that many accumulators.
#define GB_BLURY_R 8
template <int R>
__global__ void __launch_bounds__(128) k_blur_y(BlurArgs a, BlurTaps<R> taps) {
  constexpr int LEN = 2 * R + 1;
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= a.g.w) return;
  const int h = a.g.h;
  const int y_end = a.y0 + a.nrows < h ? a.y0 + a.nrows : h;
  const int strips = (a.nrows + GB_BLURY_R - 1) / GB_BLURY_R;
  const int pl = blockIdx.y / strips;
  const int yb = a.y0 + (blockIdx.y - pl * strips) * GB_BLURY_R;
  const float* col = a.in + static_cast<size_t>(pl) * a.g.plane + x;
  float* ocol = a.out + static_cast<size_t>(pl) * a.g.plane + x;
  const size_t pitch = a.g.pitch;
  const bool interior = (yb >= R) && (yb + GB_BLURY_R - 1 + R < h) && (yb + GB_BLURY_R <= y_end);
  if (interior) {
    float acc[GB_BLURY_R];
#pragma unroll
    for (int o = 0; o < GB_BLURY_R; ++o) acc[o] = 0.0f;
    const float* p = col + static_cast<size_t>(yb - R) * pitch;
#pragma unroll
    for (int t = 0; t < LEN + GB_BLURY_R - 1; ++t) {
      const float v = *p;
      p += pitch;
#pragma unroll
      for (int o = 0; o < GB_BLURY_R; ++o) {
        const int j = t - o;  // compile-time after unrolling
        if (j >= 0 && j < LEN) acc[o] += v * taps.n[j];
      }
    }
#pragma unroll
    for (int o = 0; o < GB_BLURY_R; ++o) ocol[static_cast<size_t>(yb + o) * pitch] = acc[o];
  } else {
    for (int o = 0; o < GB_BLURY_R; ++o) {
      const int y = yb + o;
      if (y >= y_end) break;
      float sum = 0.0f;
      if (y < R || y + R >= h) {
        const int lo = y < R ? 0 : y - R;
        const int hi = (y + R < h - 1) ? y + R : h - 1;
        for (int j = lo; j <= hi; ++j) sum += col[static_cast<size_t>(j) * pitch] * a.tab.taps[j - y + R];
        sum = sum * a.tab.scale_y[y];
      } else {
        for (int j = 0; j < LEN; ++j) sum += col[static_cast<size_t>(y - R + j) * pitch] * a.tab.taps_n[j];
      }
      ocol[static_cast<size_t>(y) * pitch] = sum;
    }
  }
}

template <int R>
inline void launch_blur_r(Stream s, const float* in, float* tmp, float* out, int nplanes, const BlurTab& tab,
                          const float* host_taps_n, const Geom& g, int y0, int nrows) {
  BlurTaps<R> taps;
  for (int j = 0; j < 2 * R + 1; ++j) taps.n[j] = host_taps_n[j];
  // strip mode: the y pass reads x-pass rows outside [y0, y0+nrows) that are stale; the
  // outputs they reach lie in the halo margin that the next stage no longer needs
  const int xy0 = y0, xy1 = y0 + nrows;
  BlurArgs ax{in, tmp, tab, g, nplanes * (xy1 - xy0), xy0, xy1 - xy0};
  dim3 bx(32, 8), gx((g.w + GB_BLURX_TW - 1) / GB_BLURX_TW, (nplanes * (xy1 - xy0) + 7) / 8);
  note_launch("blur_x", s, static_cast<double>(g.w) * (xy1 - xy0) * nplanes);
  k_blur_x<R><<<gx, bx, 0, s>>>(ax, taps);
  note_launch_end("blur_x", s);
  BlurArgs ay{tmp, out, tab, g, nplanes * nrows, y0, nrows};
  const int strips = (nrows + GB_BLURY_R - 1) / GB_BLURY_R;
  dim3 gy((g.w + 127) / 128, strips * nplanes);
  note_launch("blur_y", s, static_cast<double>(g.w) * nrows * nplanes);
  k_blur_y<R><<<gy, 128, 0, s>>>(ay, taps);
  note_launch_end("blur_y", s);
}

// host_taps_n: the interior kernel (taps * 1/sum) as built by tables.cc.
inline void launch_blur_tiled(Stream s, const float* in, float* tmp, float* out, int nplanes, const BlurTab& tab,
                              const float* host_taps_n, const Geom& g, int y0, int nrows) {
  switch (tab.r) {
    case 2: launch_blur_r<2>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 3: launch_blur_r<3>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 4: launch_blur_r<4>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 5: launch_blur_r<5>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 8: launch_blur_r<8>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 16: launch_blur_r<16>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 20: launch_blur_r<20>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 23: launch_blur_r<23>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    case 24: launch_blur_r<24>(s, in, tmp, out, nplanes, tab, host_taps_n, g, y0, nrows); break;
    default: throw std::runtime_error("blur radius without a compiled kernel");
  }
}

// ---------------------------------------------------------------------------
// a11 symbol histograms with a per-CTA shared-memory histogram (same counts as
// JpegHistAcc + JpegHistSum in jpeg_dev.h): out[6][257] = dc0 dc1 dc2 ac0 ac1 ac2.
__global__ void __launch_bounds__(256) k_jpeg_hist(const int16_t* cand, const int* q, const int* zigzag,
                                                   unsigned int* out, unsigned int* chroma_nonzero, int nblocks) {
  __shared__ unsigned int sh[kHistStride];
  __shared__ int sq[192];
  __shared__ int szz[64];
  for (int i = threadIdx.x; i < kHistStride; i += 256) sh[i] = 0;
  if (threadIdx.x < 192) sq[threadIdx.x] = q[threadIdx.x];
  if (threadIdx.x < 64) szz[threadIdx.x] = zigzag[threadIdx.x];
  __syncthreads();
  const int units = 3 * nblocks;
  bool chroma = false;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < units; i += gridDim.x * 256) {
    const int c = i / nblocks, b = i - c * nblocks;
    const int16_t* blk = cand + static_cast<size_t>(i) * 64;
    const int* qc = sq + 64 * c;
    const int prev = b > 0 ? (blk - 64)[0] / qc[0] : 0;
    JpegHistAcc::Visitor v{sh + c * 257, sh + (3 + c) * 257};
    visit_block_symbols(blk, qc, prev, szz, v);
    if (c > 0) {
      for (int k = 0; k < 64; ++k) chroma = chroma || (blk[k] != 0);
    }
  }
  if (chroma) *chroma_nonzero = 1u;
  __syncthreads();
  for (int i = threadIdx.x; i < kHistStride; i += 256) {
    const unsigned int n = sh[i];
    if (n) atomicAdd(&out[i], n);
  }
}

inline void launch_jpeg_hist(Stream s, const int16_t* cand, const int* q, const int* zigzag, unsigned int* out,
                             unsigned int* chroma_nonzero, int nblocks) {
  int ctas = (3 * nblocks + 255) / 256;
  if (ctas > 592) ctas = 592;  // 148 SMs x 4
  note_launch("jpeg_hist", s, 3.0 * nblocks);
  k_jpeg_hist<<<ctas, 256, 0, s>>>(cand, q, zigzag, out, chroma_nonzero, nblocks);
  note_launch_end("jpeg_hist", s);
}

// ---------------------------------------------------------------------------
// a16 key histograms with a per-CTA shared-memory histogram (keys cluster in a few
// bins: global atomics would serialise), flushed once per CTA; and the matching
// single-CTA bin search.  Same results as OrderKeyHist / OrderSelectBin (kernels.h).
__global__ void __launch_bounds__(256) k_order_hist(OrderKeyCommon c, unsigned int* hist, const OrderSelectState* st,
                                                    int level, int entries) {
  __shared__ unsigned int sh[kOrderBins];
  for (int i = threadIdx.x; i < kOrderBins; i += 256) sh[i] = 0;
  __syncthreads();
  for (int e = blockIdx.x * 256 + threadIdx.x; e < entries; e += gridDim.x * 256) {
    float v;
    int b;
    if (!c.key(e, &b, &v)) continue;
    unsigned int bin;
    if (order_bin(st, level, hd_float_sortable(v), &bin)) atomicAdd(&sh[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kOrderBins; i += 256) {
    const unsigned int n = sh[i];
    if (n) atomicAdd(&hist[i], n);
  }
}

__global__ void __launch_bounds__(1024) k_order_select_bin(const unsigned int* hist, OrderSelectState* st, int level) {
  __shared__ unsigned int warp_tot[32];
  const int t = threadIdx.x;
  const unsigned int h0 = hist[2 * t], h1 = hist[2 * t + 1];
  const unsigned int local = h0 + h1;
  unsigned int incl = local;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned int base = 0, total = 0;
  for (int k = 0; k < 32; ++k) {
    if (k < warp) base += warp_tot[k];
    total += warp_tot[k];
  }
  const unsigned int excl = base + incl - local;
  const unsigned int want = level == 0 ? st->want : (st->want > st->below0 ? st->want - st->below0 : 0u);
  // serial semantics: first bin i with cum_before(i) + hist[i] >= want; none -> last bin
  unsigned int bin = 0xffffffffu, at = 0;
  if (excl + h0 >= want) {
    if (t == 0 || excl < want || want == 0) {
      // candidate: bin 2t, valid only if no earlier bin qualifies
      bin = 2 * t;
      at = excl;
    }
  } else if (excl + local >= want) {
    bin = 2 * t + 1;
    at = excl + h0;
  }
  // the first qualifying bin overall = minimum candidate
  __shared__ unsigned int best_bin;
  if (t == 0) best_bin = 0xffffffffu;
  __syncthreads();
  if (bin != 0xffffffffu) atomicMin(&best_bin, bin);
  __syncthreads();
  const unsigned int chosen = best_bin;
  if (chosen == 0xffffffffu) {
    if (t == 0) {
      const unsigned int last = kOrderBins - 1;
      const unsigned int before = total - hist[last];
      if (level == 0) {
        st->bin0 = last;
        st->below0 = before;
        st->total = total;
      } else {
        st->threshold = (st->bin0 << 21) | (last << 10) | 0x3ffu;
        st->kept = st->below0 + before + hist[last];
        st->counter = 0;
      }
    }
    return;
  }
  if (bin == chosen) {
    if (level == 0) {
      st->bin0 = chosen;
      st->below0 = at;
      st->total = total;
    } else {
      st->threshold = (st->bin0 << 21) | (chosen << 10) | 0x3ffu;
      st->kept = st->below0 + at + hist[chosen];
      st->counter = 0;
    }
  }
}

inline void launch_order_hist(Stream s, const OrderKeyCommon& c, unsigned int* hist, const OrderSelectState* st,
                              int level, int entries) {
  int ctas = (entries + 256 * 8 - 1) / (256 * 8);
  if (ctas < 1) ctas = 1;
  if (ctas > 1184) ctas = 1184;  // 148 SMs x 8 resident CTAs
  note_launch("order_key_hist", s, entries);
  k_order_hist<<<ctas, 256, 0, s>>>(c, hist, st, level, entries);
  note_launch_end("order_key_hist", s);
}

inline void launch_order_select_bin(Stream s, const unsigned int* hist, OrderSelectState* st, int level) {
  note_launch("order_select_bin", s, kOrderBins);
  k_order_select_bin<<<1, 1024, 0, s>>>(hist, st, level);
  note_launch_end("order_select_bin", s);
}

}  // namespace gb200
