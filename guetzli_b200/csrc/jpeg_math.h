// Integer JPEG arithmetic of the hot path, single source for device + CPU port.
// All of it is exact integer work: results must equal the reference bit for bit.
#pragma once
#include "hd.h"

namespace gb200 {

// ---------------------------------------------------------------------------
// Quantize (quantize.h:24): nearest multiple of q, C '%' semantics, ties to zero.
GB_HD int quantize_coeff(int raw, int q) {
  const int r = raw % q;
  const int delta = 2 * r > q ? q - r : (-2) * r > q ? -q - r : -r;
  return static_cast<int16_t>(raw + delta);
}

// ---------------------------------------------------------------------------
// 8-point inverse DCT as a dot product with the 13-bit basis (idct.cc:41):
// out[x] = sum_u basis[8x+u] * in[u].  int32 wrap-around like the reference.
GB_HD void idct_1d(const int* basis, const int in[8], int out[8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int x = 0; x < 8; ++x) {
    int acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int u = 0; u < 8; ++u) acc += basis[8 * x + u] * in[u];
    out[x] = acc;
  }
}

// ComputeBlockIDCT (idct.cc:139): column pass (+2^10 >> 11, stored as int16),
// row pass (+257*2^17 >> 18, includes the +128 level shift), clamp to u8.
GB_HD void idct_8x8(const int* basis, const int16_t* block, uint8_t* out) {
  int16_t col[64];
  for (int x = 0; x < 8; ++x) {
    int in[8], o[8];
    for (int u = 0; u < 8; ++u) in[u] = block[8 * u + x];
    idct_1d(basis, in, o);
    for (int y = 0; y < 8; ++y) col[8 * y + x] = static_cast<int16_t>((o[y] + (1 << 10)) >> 11);
  }
  for (int y = 0; y < 8; ++y) {
    int in[8], o[8];
    for (int u = 0; u < 8; ++u) in[u] = col[8 * y + u];
    idct_1d(basis, in, o);
    for (int x = 0; x < 8; ++x) {
      int v = (o[x] + (257 << 17)) >> 18;
      out[8 * y + x] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

// ---------------------------------------------------------------------------
// YCbCr -> RGB (color_transform.h:211) with the range-limit table folded into a clamp.
GB_HD int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

GB_HD void ycc_to_rgb(const int* cr_r, const int* cb_b, const int* cr_g, const int* cb_g, int y,
                      int cb, int cr, int* r, int* g, int* b) {
  *r = clamp_u8(y + cr_r[cr]);
  *g = clamp_u8(y + ((cr_g[cr] + cb_g[cb]) >> 16));
  *b = clamp_u8(y + cb_b[cb]);
}

// ---------------------------------------------------------------------------
// RGB -> YCbCr, 16.16 fixed point, output centred on zero (jpeg_data_encoder.cc:40).
GB_HD void rgb_to_ycc16(int r, int g, int b, int16_t* y, int16_t* cb, int16_t* cr) {
  const int kHalf = 1 << 15;
  *y = static_cast<int16_t>((19595 * r + 38469 * g + 7471 * b - (128 << 16) + kHalf) >> 16);
  *cb = static_cast<int16_t>((-11059 * r - 21709 * g + 32768 * b + kHalf - 1) >> 16);
  *cr = static_cast<int16_t>((32768 * r - 27439 * g - 5329 * b + kHalf - 1) >> 16);
}

// Forward DCT (fdct.cc:160-240), output scaled by 16.  The fixed-point data
// flow (which products are truncated by >>16, what is narrowed to int16 between
// the passes) is what makes the result; additions are exact in int32.
GB_HD int mulhi16(int a, int b) { return (a * b) >> 16; }

// Vertical pass over one column (stride 8), in place.
GB_HD void fdct_column(int16_t* v) {
  const int kTan1 = 13036, kTan2 = 27146, kTan3m1 = -21746, k2Sqrt2 = 23170;
  // first butterfly stage: d = differences, s = sums of mirrored rows
  int d07 = v[0 * 8] - v[7 * 8], s07 = v[0 * 8] + v[7 * 8];
  int d25 = v[2 * 8] - v[5 * 8], s25 = v[2 * 8] + v[5 * 8];
  int d34 = v[3 * 8] - v[4 * 8], s34 = v[3 * 8] + v[4 * 8];
  int d16 = v[1 * 8] - v[6 * 8], s16 = v[1 * 8] + v[6 * 8];
  // even part
  int e_d = s07 - s34, e_s = s07 + s34;   // (m7, m4) after BUTTERFLY(m7, m4)
  int f_d = s16 - s25, f_s = s16 + s25;   // (m6, m5) after BUTTERFLY(m6, m5)
  int es = e_s << 3, fs = f_s << 3;
  v[0 * 8] = static_cast<int16_t>(es + fs);
  v[4 * 8] = static_cast<int16_t>(es - fs);
  int ed = e_d << 3, fd = f_d << 3;
  v[2 * 8] = static_cast<int16_t>(mulhi16(kTan2, fd) + ed);
  v[6 * 8] = static_cast<int16_t>(mulhi16(kTan2, ed) - fd);
  // odd part
  int a3 = d34 << 3, a0 = d07 << 3;
  int b2 = d25 << 4, b1 = d16 << 4;
  int p = mulhi16(b1 + b2, k2Sqrt2);   // m2 after BUTTERFLY(m1,m2); MULT
  int q = mulhi16(b1 - b2, k2Sqrt2);   // m1
  int t3 = a3 - q, t1 = a3 + q;        // BUTTERFLY(m3, m1): m3 = a3 - q, m1 = a3 + q
  int t0 = a0 - p, t2 = a0 + p;        // BUTTERFLY(m0, m2)
  int m3 = mulhi16(t3, kTan3m1) + t3;  // t3 * tan3
  int m1 = mulhi16(t1, kTan1) + t2;
  m1 += 1;                             // CORRECT_LSB
  m3 += 1;
  int m4 = mulhi16(kTan3m1, t0) + t0;  // t0 * tan3
  int m5 = mulhi16(kTan1, t2);
  v[1 * 8] = static_cast<int16_t>(m1);
  v[3 * 8] = static_cast<int16_t>(t0 - m3);
  v[5 * 8] = static_cast<int16_t>(t3 + m4);
  v[7 * 8] = static_cast<int16_t>(m5 - t1);
}

// Horizontal pass over one row with its 7-entry cosine table (fdct.cc:173).
GB_HD void fdct_row(int16_t* in, const int16_t* table) {
  const int a0 = in[0] + in[7], b0 = in[0] - in[7];
  const int a1 = in[1] + in[6], b1 = in[1] - in[6];
  const int a2 = in[2] + in[5], b2 = in[2] - in[5];
  const int a3 = in[3] + in[4], b3 = in[3] - in[4];
  const int C1 = table[0], C2 = table[1], C3 = table[2], C4 = table[3], C5 = table[4],
            C6 = table[5], C7 = table[6];
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  in[0] = static_cast<int16_t>((C4 * (c0 + c2)) >> 16);
  in[4] = static_cast<int16_t>((C4 * (c0 - c2)) >> 16);
  in[2] = static_cast<int16_t>((C2 * c1 + C6 * c3) >> 16);
  in[6] = static_cast<int16_t>((C6 * c1 - C2 * c3) >> 16);
  in[1] = static_cast<int16_t>((C1 * b0 + C3 * b1 + C5 * b2 + C7 * b3) >> 16);
  in[3] = static_cast<int16_t>((C3 * b0 - C7 * b1 - C1 * b2 - C5 * b3) >> 16);
  in[5] = static_cast<int16_t>((C5 * b0 - C1 * b1 + C7 * b2 + C3 * b3) >> 16);
  in[7] = static_cast<int16_t>((C7 * b0 - C5 * b1 + C3 * b2 - C1 * b3) >> 16);
}

GB_HD void fdct_8x8(int16_t* c) {
  // cos(k*pi/16)/sqrt(2) in 15 bits; rows 1/7, 2/6, 3/5 pre-multiplied by
  // 2cos(pi/16), 2cos(2pi/16), 2cos(3pi/16) (fdct.cc:29-36)
  const int16_t t04[7] = {22725, 21407, 19266, 16384, 12873, 8867, 4520};
  const int16_t t17[7] = {31521, 29692, 26722, 22725, 17855, 12299, 6270};
  const int16_t t26[7] = {29692, 27969, 25172, 21407, 16819, 11585, 5906};
  const int16_t t35[7] = {26722, 25172, 22654, 19266, 15137, 10426, 5315};
  for (int x = 0; x < 8; ++x) fdct_column(c + x);
  fdct_row(c + 0 * 8, t04);
  fdct_row(c + 1 * 8, t17);
  fdct_row(c + 2 * 8, t26);
  fdct_row(c + 3 * 8, t35);
  fdct_row(c + 4 * 8, t04);
  fdct_row(c + 5 * 8, t35);
  fdct_row(c + 6 * 8, t26);
  fdct_row(c + 7 * 8, t17);
}

// Quantize by 1 after the x16 DCT scale (jpeg_data_encoder.cc:33 with iquant=65537).
GB_HD int16_t fdct_descale(int16_t v) {
  return static_cast<int16_t>((v * 65537 + (0x80 << 12)) >> 20);
}

}  // namespace gb200
