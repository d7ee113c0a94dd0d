// 8x8 block metric and greedy zeroing order (a13/a14): the body of
// Processor::ComputeBlockZeroingOrder (g/processor.cc:364-467) with
// ButteraugliComparator::SwitchBlock / CompareBlock
// (g/butteraugli_comparator.cc:427-488) inlined, one 8x8 block per invocation.
// g/ = /root/reference/guetzli/.
#pragma once
#include "ba_math.h"
#include "jpeg_math.h"
#include "kernels.h"

namespace gb200 {

// ---------------------------------------------------------------------------
// 8-point DFTs with the operation order of the reference's DJB-derived
// butterflies (g/butteraugli_comparator.cc:154-353), written as explicit data
// flow.  h = sqrt(1/2).  Results are bit-identical: IEEE addition is
// commutative and sign-symmetric, so only the association is preserved.
struct Cplx {
  double re, im;
};

GB_HD void real_dft8(const double in[8], Cplx out[8]) {
  const double h = 0.70710678118654752440084436210484903;
  const double s04 = in[4] + in[0], d04 = in[0] - in[4];
  const double s26 = in[6] + in[2], d26 = in[2] - in[6];
  const double s15 = in[5] + in[1], d15 = in[1] - in[5];
  const double s37 = in[7] + in[3], d37 = in[3] - in[7];
  const double a = (d15 - d37) * h;
  const double b = (d15 + d37) * h;
  const double even0 = s26 + s04, even1 = s37 + s15;
  out[0].re = even0 + even1;
  out[0].im = 0;
  out[4].re = even0 - even1;
  out[4].im = 0;
  out[2].re = s04 - s26;
  out[2].im = -(s15 - s37);
  out[6].re = s04 - s26;
  out[6].im = s15 - s37;
  out[1].re = a + d04;
  out[1].im = -(b + d26);
  out[7].re = a + d04;
  out[7].im = b + d26;
  out[3].re = d04 - a;
  out[3].im = d26 - b;
  out[5].re = d04 - a;
  out[5].im = -(d26 - b);
}

GB_HD void cplx_dft8(Cplx a[8]) {
  const double h = 0.70710678118654752440084436210484903;
  const double sr04 = a[4].re + a[0].re, dr04 = a[0].re - a[4].re;
  const double si04 = a[4].im + a[0].im, di04 = a[0].im - a[4].im;
  const double sr26 = a[6].re + a[2].re, dr26 = a[2].re - a[6].re;
  const double si26 = a[6].im + a[2].im, di26 = a[2].im - a[6].im;
  const double sr15 = a[5].re + a[1].re, dr15 = a[1].re - a[5].re;
  const double si15 = a[5].im + a[1].im, di15 = a[1].im - a[5].im;
  const double sr37 = a[7].re + a[3].re, dr37 = a[3].re - a[7].re;
  const double si37 = a[7].im + a[3].im, di37 = a[3].im - a[7].im;
  // odd half
  const double o4re = dr04 - di26, o4im = di04 + dr26;
  const double o6re = dr04 + di26, o6im = di04 - dr26;
  const double u = dr15 - di37, v = di15 + dr37;
  const double p = di15 - dr37, q = dr15 + di37;
  const double A = (u - v) * h, B = (u + v) * h;
  const double C = (p - q) * h, D = (p + q) * h;
  Cplx r[8];
  r[3].re = o4re - A;  // a5
  r[3].im = o4im - B;
  r[7].re = A + o4re;  // a4
  r[7].im = B + o4im;
  r[5].re = o6re - D;  // a7
  r[5].im = o6im - C;
  r[1].re = D + o6re;  // a6
  r[1].im = C + o6im;
  // even half (4-point)
  const double er = sr26 + sr04, fr = sr37 + sr15;
  const double ei = si26 + si04, fi = si37 + si15;
  r[0].re = er + fr;
  r[0].im = ei + fi;
  r[4].re = er - fr;
  r[4].im = ei - fi;
  const double gr = sr04 - sr26, gi = si04 - si26;
  const double hr = sr15 - sr37, hi = si15 - si37;
  r[6].re = gr - hi;  // a2'
  r[6].im = gi + hr;
  r[2].re = gr + hi;  // a3'
  r[2].im = gi - hr;
  for (int i = 0; i < 8; ++i) a[i] = r[i];
}

// ButteraugliFFTSquared (g/butteraugli_comparator.cc:357): power spectrum bins
// 4..36 of the packed half-spectrum, times 0.000064; weighted sum with csf.
// Returns sum_{i=4}^{36} csf[i] * |F[i]|^2 * 0.000064 accumulated onto acc in
// index order (ButteraugliBlockDiff :405-410).
GB_HD double block_spectrum_cost(const double diff[64], const double* csf, double acc) {
  Cplx f[64];  // [freq of row transform][row index]
  for (int y = 0; y < 8; ++y) {
    Cplx row[8];
    real_dft8(diff + 8 * y, row);
    for (int k = 0; k < 8; ++k) f[8 * k + y] = row[k];
  }
  double r0[8], r1[8];
  for (int x = 0; x < 8; ++x) {
    r0[x] = f[x].re;
    r1[x] = f[32 + x].re;
  }
  real_dft8(r0, f);
  real_dft8(r1, f + 32);
  for (int y = 1; y < 4; ++y) cplx_dft8(f + 8 * y);
  const double global_mul = 0.000064;
  for (int i = 4; i < 37; ++i) {
    double p = f[i].re * f[i].re + f[i].im * f[i].im;
    p *= global_mul;
    acc += csf[i] * p;
  }
  return acc;
}

// OpsinDynamicsImage on an 8x8 tile (b/butteraugli.cc:324 via
// g/butteraugli_comparator.cc:450,469): sigma 1.2 blur (r=2) with the border
// rule on rows/cols 0,1,6,7, then the per-pixel opsin.
GB_HD void opsin_8x8(const float lin[3][64], const BlurTab& tab, const float* scale8,
                     float xyb[3][64]) {
  float blr[3][64];
  for (int c = 0; c < 3; ++c) {
    float tmp[64];
    for (int y = 0; y < 8; ++y) {
      BlurRowAt at{lin[c] + 8 * y};
      for (int x = 0; x < 8; ++x)
        tmp[8 * y + x] = blur_tap_sum(at, tab.taps, tab.taps_n, scale8, tab.r, x, 8);
    }
    for (int x = 0; x < 8; ++x) {
      BlurColAt at{tmp + x, 8};
      for (int y = 0; y < 8; ++y)
        blr[c][8 * y + x] = blur_tap_sum(at, tab.taps, tab.taps_n, scale8, tab.r, y, 8);
    }
  }
  for (int i = 0; i < 64; ++i)
    opsin_pixel(lin[0][i], lin[1][i], lin[2][i], blr[0][i], blr[1][i], blr[2][i], &xyb[0][i],
                &xyb[1][i], &xyb[2][i]);
}

// ---------------------------------------------------------------------------
// libstdc++ std::sort (introsort) replayed for small arrays of (key, id) with
// comparator key_a < key_b, so that equal keys land exactly where the
// reference's std::sort (g/processor.cc:398) puts them.
struct SortItem {
  float key;
  int id;
};
GB_HD bool sort_less(const SortItem& a, const SortItem& b) { return a.key < b.key; }
GB_HD void sort_swap(SortItem& a, SortItem& b) {
  SortItem t = a;
  a = b;
  b = t;
}

GB_HD void sort_adjust_heap(SortItem* first, int hole, int len, SortItem value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (sort_less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  // __push_heap
  int parent = (hole - 1) / 2;
  while (hole > top && sort_less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

GB_HD void sort_heapsort(SortItem* first, int n) {
  // __heap_select(first, last, last) == make_heap; then __sort_heap
  if (n >= 2) {
    int parent = (n - 2) / 2;
    while (true) {
      SortItem v = first[parent];
      sort_adjust_heap(first, parent, n, v);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int last = n; last > 1;) {
    --last;
    SortItem v = first[last];
    first[last] = first[0];
    sort_adjust_heap(first, 0, last, v);
  }
}

GB_HD void sort_unguarded_linear_insert(SortItem* a, int last) {
  SortItem val = a[last];
  int next = last - 1;
  while (sort_less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

GB_HD void sort_insertion(SortItem* a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (sort_less(a[i], a[first])) {
      SortItem val = a[i];
      for (int j = i; j > first; --j) a[j] = a[j - 1];
      a[first] = val;
    } else {
      sort_unguarded_linear_insert(a, i);
    }
  }
}

GB_HD void std_sort_replay(SortItem* a, int n) {
  if (n < 2) return;
  // __introsort_loop with an explicit stack of pending [first,last) ranges.
  int depth_limit = 0;
  for (int m = n; m > 1; m >>= 1) ++depth_limit;  // floor(log2 n)
  depth_limit *= 2;
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 0;
  stack_first[sp] = 0;
  stack_last[sp] = n;
  stack_depth[sp] = depth_limit;
  ++sp;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > 16) {
      if (depth == 0) {
        sort_heapsort(a + first, last - first);
        break;
      }
      --depth;
      // __move_median_to_first(first, first+1, mid, last-1)
      const int mid = first + (last - first) / 2;
      const int ia = first + 1, ib = mid, ic = last - 1;
      if (sort_less(a[ia], a[ib])) {
        if (sort_less(a[ib], a[ic])) sort_swap(a[first], a[ib]);
        else if (sort_less(a[ia], a[ic])) sort_swap(a[first], a[ic]);
        else sort_swap(a[first], a[ia]);
      } else if (sort_less(a[ia], a[ic])) {
        sort_swap(a[first], a[ia]);
      } else if (sort_less(a[ib], a[ic])) {
        sort_swap(a[first], a[ic]);
      } else {
        sort_swap(a[first], a[ib]);
      }
      // __unguarded_partition(first+1, last, pivot=first)
      int lo = first + 1, hi = last;
      while (true) {
        while (sort_less(a[lo], a[first])) ++lo;
        --hi;
        while (sort_less(a[first], a[hi])) --hi;
        if (!(lo < hi)) break;
        sort_swap(a[lo], a[hi]);
        ++lo;
      }
      const int cut = lo;
      // recurse on [cut,last) (pushed), loop on [first,cut)
      stack_first[sp] = cut;
      stack_last[sp] = last;
      stack_depth[sp] = depth;
      ++sp;
      last = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    sort_insertion(a, 0, 16);
    for (int i = 16; i < n; ++i) sort_unguarded_linear_insert(a, i);
  } else {
    sort_insertion(a, 0, n);
  }
}

// Score that orders the candidate coefficients of a block (g/processor.cc:369-393):
// the regression tables of order.inc (new_zeroing_model, the default) or the legacy
// formula (|orig| - zigzag_pos/64) * weight[c] / oldCsf[k].
GB_HD float zeroing_score(int abs_orig, int idx, bool new_model, const Tables& t) {
  if (new_model) return abs_orig * t.order_csf[idx] + t.order_bias[idx];
  const int c = idx >> 6, k = idx & 63;
  const double weight = c == 0 ? 1.0 : (c == 1 ? 0.22 : 0.20);
  return static_cast<float>((abs_orig - t.nat2zz[k] / 64.0) * weight / t.order_old_csf[k]);
}

// ---------------------------------------------------------------------------
struct ZeroingOrders {
  const int16_t* cand;   // [3][nblocks][64] after global quantisation
  const int16_t* orig;   // [3][nblocks][64] original (q = 1) coefficients
  const uint8_t* rgb;    // original sRGB, interleaved
  const float* corner_mask;  // [nblocks][3] mask_xyz_ at (8bx, 8by)
  uint8_t* out_idx;      // [nblocks][192]
  float* out_err;        // [nblocks][192]
  int* out_count;        // [nblocks]
  Geom g;
  Tables t;
  const float* scale8;   // border scales of the sigma-1.2 blur on an 8-long axis
  int lookahead;         // Params::zeroing_greedy_lookahead (3)
  float block_error_limit;
  int new_model;         // Params::new_zeroing_model

  // CompareBlock: pixels (YCbCr u8, full 8x8 IDCT output) -> error.
  GB_HD float compare_block(const uint8_t px[3][64], int xlast, int ylast, const float xyb0[3][64],
                            const float mask[3]) const {
    float lin[3][64];
    for (int iy = 0; iy < 8; ++iy) {
      const int sy = iy < ylast ? iy : ylast;
      for (int ix = 0; ix < 8; ++ix) {
        const int sx = ix < xlast ? ix : xlast;
        const int s = 8 * sy + sx;
        int r, gg, bb;
        ycc_to_rgb(t.cr_r, t.cb_b, t.cr_g, t.cb_g, px[0][s], px[1][s], px[2][s], &r, &gg, &bb);
        lin[0][8 * iy + ix] = t.srgb_lin[r];
        lin[1][8 * iy + ix] = t.srgb_lin[gg];
        lin[2][8 * iy + ix] = t.srgb_lin[bb];
      }
    }
    float xyb1[3][64];
    opsin_8x8(lin, t.blur[kBlurOpsin], scale8, xyb1);
    double diff = 0.0;
    for (int c = 0; c < 3; ++c) {
      double d[64];
      double avg = 0.0;
      for (int i = 0; i < 64; ++i) {
        d[i] = static_cast<double>(xyb0[c][i]) - static_cast<double>(xyb1[c][i]);
        avg += d[i];
      }
      const double avgdiff = avg / 64;
      double dc = 0.0;
      dc += 4.0 * avgdiff * avgdiff;
      dc = block_spectrum_cost(d, t.block_csf, dc);
      diff += dc * mask[c];
    }
    return static_cast<float>(sqrt(diff));
  }

  GB_HD void operator()(int b) const {
    const int bx = b % g.bw, by = b / g.bw;
    const int xlast = hd_min(7, g.w - 1 - 8 * bx), ylast = hd_min(7, g.h - 1 - 8 * by);
    int16_t blk[192];
    SortItem order[189];
    int n = 0;
    for (int c = 0; c < 3; ++c) {
      const int16_t* cb = cand + (static_cast<size_t>(c) * g.nblocks + b) * 64;
      const int16_t* ob = orig + (static_cast<size_t>(c) * g.nblocks + b) * 64;
      for (int k = 0; k < 64; ++k) blk[64 * c + k] = cb[k];
      for (int k = 1; k < 64; ++k) {
        const int idx = 64 * c + k;
        if (cb[k] != 0) {
          const int a = ob[k] < 0 ? -ob[k] : ob[k];
          order[n].key = zeroing_score(a, idx, new_model != 0, t);
          order[n].id = idx;
          ++n;
        }
      }
    }
    std_sort_replay(order, n);

    // SwitchBlock: original tile (edge-replicated) -> linear -> opsin.
    float xyb0[3][64];
    {
      float lin[3][64];
      for (int iy = 0; iy < 8; ++iy) {
        const int y = hd_min(8 * by + iy, g.h - 1);
        for (int ix = 0; ix < 8; ++ix) {
          const int x = hd_min(8 * bx + ix, g.w - 1);
          const uint8_t* p = rgb + 3 * (static_cast<size_t>(y) * g.w + x);
          for (int c = 0; c < 3; ++c) lin[c][8 * iy + ix] = t.srgb_lin[p[c]];
        }
      }
      opsin_8x8(lin, t.blur[kBlurOpsin], scale8, xyb0);
    }
    float mask[3] = {corner_mask[3 * b], corner_mask[3 * b + 1], corner_mask[3 * b + 2]};

    uint8_t px[3][64];
    for (int c = 0; c < 3; ++c) idct_8x8(t.idct, blk + 64 * c, px[c]);

    uint8_t* oi = out_idx + static_cast<size_t>(b) * 192;
    float* oe = out_err + static_cast<size_t>(b) * 192;
    int nout = 0;
    while (n > 0) {
      float best_err = 1e17f;
      int best_i = 0;
      const int tries = lookahead < n ? lookahead : n;
      for (int i = 0; i < tries; ++i) {
        const int idx = order[i].id;
        const int c = idx >> 6;
        const int16_t saved = blk[idx];
        uint8_t saved_px[64];
        for (int k = 0; k < 64; ++k) saved_px[k] = px[c][k];
        blk[idx] = 0;
        idct_8x8(t.idct, blk + 64 * c, px[c]);
        const float err = compare_block(px, xlast, ylast, xyb0, mask);
        float max_err = 0;
        max_err = hd_max(max_err, err);
        if (max_err < best_err) {
          best_err = max_err;
          best_i = i;
        }
        blk[idx] = saved;
        for (int k = 0; k < 64; ++k) px[c][k] = saved_px[k];
      }
      const int idx = order[best_i].id;
      blk[idx] = 0;
      idct_8x8(t.idct, blk + 64 * (idx >> 6), px[idx >> 6]);
      for (int i = best_i; i + 1 < n; ++i) order[i] = order[i + 1];
      --n;
      oi[nout] = static_cast<uint8_t>(idx);
      oe[nout] = best_err;
      ++nout;
    }
    // monotone suffix minimum, then cut at the block error limit (:447-459)
    float min_err = 1e10f;
    for (int i = nout - 1; i >= 0; --i) {
      min_err = hd_min(min_err, oe[i]);
      oe[i] = min_err;
    }
    int num = 0;
    while (num < nout && oe[num] <= block_error_limit) ++num;
    out_count[b] = num;
  }
};

}  // namespace gb200
