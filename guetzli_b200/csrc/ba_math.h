// Per-pixel arithmetic of the butteraugli distance (2017 "PsychoImage/Malta"
// version vendored by the reference at third_party/butteraugli/), single source
// for the CUDA kernels and the CPU port.  File:line citations are into
// /root/reference/third_party/butteraugli/butteraugli/.
//
// Every expression keeps the reference's evaluation order and its implicit
// float<->double promotions; they are written out explicitly here because the
// search loop needs bit-identical scores, not merely close ones.
#pragma once
#include "hd.h"
#include "tables.h"

namespace gb200 {

// ---------------------------------------------------------------------------
// OpsinAbsorbance<float> (butteraugli.h:499): 3x3 mix + bias, constants narrowed
// to float, evaluated left to right.
GB_HD void opsin_absorbance(float r, float g, float b, float* o0, float* o1, float* o2) {
  const float m0 = static_cast<float>(0.254462330846);
  const float m1 = static_cast<float>(0.488238255095);
  const float m2 = static_cast<float>(0.0635278003854);
  const float m3 = static_cast<float>(1.01681026909);
  const float m4 = static_cast<float>(0.195214015766);
  const float m5 = static_cast<float>(0.568019861857);
  const float m6 = static_cast<float>(0.0860755536007);
  const float m7 = static_cast<float>(1.1510118369);
  const float m8 = static_cast<float>(0.07374607900105684);
  const float m9 = static_cast<float>(0.06142425304154509);
  const float m10 = static_cast<float>(0.24416850520714256);
  const float m11 = static_cast<float>(1.20481945273);
  *o0 = m0 * r + m1 * g + m2 * b + m3;
  *o1 = m4 * r + m5 * g + m6 * b + m7;
  *o2 = m8 * r + m9 * g + m10 * b + m11;
}

// Chebyshev series by Clenshaw's recurrence, degree 5 (butteraugli.h:548-568).
GB_HD double clenshaw5(double x, const double c[6]) {
  double b1 = 0.0, b2 = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 5; i >= 1; --i) {
    const double x_b1 = x * b1;
    const double t = (x_b1 + x_b1) - b2 + c[i];
    b2 = b1;
    b1 = t;
  }
  const double x_b1 = x * b1;
  return x_b1 - b2 + c[0];
}

// GammaPolynomial (butteraugli.h:599): rational 5/5 on [0.971783, 590.188894],
// quotient narrowed to float and returned as double.
GB_HD double gamma_poly(double v) {
  const double p[6] = {98.7821300963361, 164.273222212631, 92.948112871376,
                       33.8165311212688, 6.91626704983562, 0.556380877028234};
  const double q[6] = {1, 1.64339473427892, 0.89392405219969,
                       0.298947051776379, 0.0507146002577288, 0.00226495093949756};
  const double min_value = 0.971783, max_value = 590.188894;
  const double x01 = (v - min_value) / (max_value - min_value);
  const double xc = 2.0 * x01 - 1.0;
  const double yp = clenshaw5(xc, p);
  const double yq = clenshaw5(xc, q);
  if (yq == 0.0) return 0.0;
  return static_cast<float>(yp / yq);
}

// One pixel of OpsinDynamicsImage (butteraugli.cc:337-362): sensitivity from the
// blurred image, applied to the sharp one, then RgbToXyb.
GB_HD void opsin_pixel(float r, float g, float b, float br, float bg, float bb, float* ox,
                       float* oy, float* ob) {
  float pre0, pre1, pre2;
  opsin_absorbance(br, bg, bb, &pre0, &pre1, &pre2);
  const float s0 = static_cast<float>(gamma_poly(pre0) / static_cast<double>(pre0));
  const float s1 = static_cast<float>(gamma_poly(pre1) / static_cast<double>(pre1));
  const float s2 = static_cast<float>(gamma_poly(pre2) / static_cast<double>(pre2));
  float c0, c1, c2;
  opsin_absorbance(r, g, b, &c0, &c1, &c2);
  c0 *= s0;
  c1 *= s1;
  c2 *= s2;
  *ox = c0 - c1;
  *oy = c0 + c1;
  *ob = c2;
}

// ---------------------------------------------------------------------------
// Separable blur taps (butteraugli.cc:156-233).  `at(j)` returns the sample at
// axis position j.  p = output position, n = axis length.
template <class At>
GB_HD float blur_tap_sum(const At& at, const float* taps, const float* taps_n, const float* scale,
                         int r, int p, int n) {
  if (p < r || p + r >= n) {
    // ConvolveBorderColumn: raw taps over the clipped support, then * 1/weight.
    const int lo = p < r ? 0 : p - r;
    const int hi = (p + r < n - 1) ? p + r : n - 1;
    float sum = 0.0f;
    for (int j = lo; j <= hi; ++j) sum += at(j) * taps[j - p + r];
    return sum * scale[p];
  }
  float sum = 0.0f;
  const int d = p - r;
  const int len = 2 * r + 1;
  for (int j = 0; j < len; ++j) sum += at(d + j) * taps_n[j];
  return sum;
}

// ---------------------------------------------------------------------------
// SeparateFrequencies pointwise pieces (butteraugli.cc:369-487).
GB_HD float remove_range_around_zero(float w, float x) {
  return x > w ? x - w : x < -w ? x + w : 0.0f;
}
GB_HD float amplify_range_around_zero(float w, float x) {
  return x > w ? x + w : x < -w ? x - w : 2.0f * x;
}
GB_HD float maximum_clamp(float v, float maxval) {
  const double kMul = 0.688059627878;
  if (v >= maxval) {
    v -= maxval;
    v = static_cast<float>(v * kMul);
    v += maxval;
  } else if (v < -maxval) {
    v += maxval;
    v = static_cast<float>(v * kMul);
    v -= maxval;
  }
  return v;
}
GB_HD float suppress_in_bright_areas(float hf, float brightness, float mul, float reg) {
  const float scaler = mul * reg / (reg + brightness);
  return scaler * hf;
}
GB_HD float suppress_x_by_y(float xv, float yv) {
  const double s = 0.745954517135;
  const double yw = 2.96534974403;
  const double xval = xv, yval = yv;
  const double scaler = s + (yw * (1.0 - s)) / (yw + yval * yval);
  return static_cast<float>(scaler * xval);
}

// ---------------------------------------------------------------------------
// Malta pre-pass (butteraugli.cc:1476-1529): one "diffs" sample.
GB_HD float malta_diff(float v0, float v1, const MaltaParams& mp) {
  // one float -> double conversion per input; everything below reuses them
  const double r0 = v0, r1 = v1;
  const double fabs0 = ::fabs(r0), fabs1 = ::fabs(r1);
  const float absval = static_cast<float>(0.5 * fabs0 + 0.5 * fabs1);
  const float diff = v0 - v1;
  const float den = mp.norm1 + absval;
  float d = (mp.norm2_0gt1 / den) * diff;
  const double too_small = 0.55 * fabs0;
  const double too_big = 1.05 * fabs0;
  // which half-open objective applies, if any (the second division only happens then)
  double excess;
  bool has = false;
  if (v0 < 0) {
    if (r1 > -too_small) {
      excess = r1 + too_small;
      has = true;
    } else if (r1 < -too_big) {
      excess = -r1 - too_big;
      has = true;
    }
  } else {
    if (r1 < too_small) {
      excess = too_small - r1;
      has = true;
    } else if (r1 > too_big) {
      excess = r1 - too_big;
      has = true;
    }
  }
  if (has) {
    const float scaler2 = mp.norm2_0lt1 / den;
    const double impact = scaler2 * excess;
    if (diff < 0) {
      d = static_cast<float>(static_cast<double>(d) - impact);
    } else {
      d = static_cast<float>(static_cast<double>(d) + impact);
    }
  }
  return d;
}

// ---------------------------------------------------------------------------
// InterpolateClampNegative (butteraugli.cc:236) over a 512-entry LUT.
GB_HD double mask_lut_eval(const double* lut, double ix) {
  if (ix < 0) ix = 0;
  const int baseix = static_cast<int>(ix);
  if (baseix >= 511) return lut[511];
  const double mix = ix - baseix;
  return lut[baseix] + mix * (lut[baseix + 1] - lut[baseix]);
}

// Final per-pixel step of Mask (butteraugli.cc:1790-1815): blurred activity
// s0 (X), s1 (Y) -> three AC masks and three DC masks (stored as float).
GB_HD void mask_from_activity(const double* luts, float s0f, float s1f, float mask[3],
                              float mask_dc[3]) {
  const double s0 = s0f, s1 = s1f;
  const double p1 = 2.1364621982 * 2.1887170895 * s1;
  const double p0 = 16.6963293877 * 36.4671237619 * s0 + 0.0513061271723 * p1;
  const double my = mask_lut_eval(luts + 512, p1);
  const double mdy = mask_lut_eval(luts + 1536, p1);
  mask[0] = static_cast<float>(mask_lut_eval(luts, p0));
  mask[1] = static_cast<float>(my);
  mask[2] = static_cast<float>(0.086624184478 * my);
  mask_dc[0] = static_cast<float>(mask_lut_eval(luts + 1024, p0));
  mask_dc[1] = static_cast<float>(mdy);
  mask_dc[2] = static_cast<float>(21.6804277046 * mdy);
}

}  // namespace gb200
