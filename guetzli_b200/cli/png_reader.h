// Minimal PNG decoder for the drop-in `guetzli` CLI (the reference uses libpng,
// guetzli/guetzli.cc:47-152; libpng headers are not available in this image).
// Produces what png_read_png(PACKING | EXPAND | STRIP_16) + the reference's
// channel handling produce: 8-bit RGB, alpha blended on black.
// Supports colour types 0/2/3/4/6, bit depths 1..16, tRNS, non-interlaced and Adam7.
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

namespace gb200_cli {

inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

inline uint8_t BlendOnBlack(const uint8_t val, const uint8_t alpha) {
  return (static_cast<int>(val) * static_cast<int>(alpha) + 128) / 255;
}

inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Un-filters one pass of `h` scanlines of `rowbytes` bytes each (bpp = bytes per complete pixel, >= 1).
inline bool unfilter(const uint8_t* in, size_t in_len, int h, size_t rowbytes, int bpp, std::vector<uint8_t>* out) {
  if (in_len < static_cast<size_t>(h) * (rowbytes + 1)) return false;
  out->assign(static_cast<size_t>(h) * rowbytes, 0);
  for (int y = 0; y < h; ++y) {
    const uint8_t ft = in[y * (rowbytes + 1)];
    const uint8_t* src = in + y * (rowbytes + 1) + 1;
    uint8_t* cur = out->data() + static_cast<size_t>(y) * rowbytes;
    const uint8_t* up = y ? cur - rowbytes : nullptr;
    for (size_t i = 0; i < rowbytes; ++i) {
      const int a = i >= static_cast<size_t>(bpp) ? cur[i - bpp] : 0;
      const int b = up ? up[i] : 0;
      const int c = (up && i >= static_cast<size_t>(bpp)) ? up[i - bpp] : 0;
      int v = src[i];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return false;
      }
      cur[i] = static_cast<uint8_t>(v);
    }
  }
  return true;
}

// Decodes to 8-bit RGBA (alpha 255 where the file has none); *has_alpha tells whether the
// file carries transparency (alpha channel or tRNS).  16-bit samples keep their high byte.
inline bool ReadPNGRGBA(const std::string& data, int* xsize, int* ysize, std::vector<uint8_t>* rgba,
                        bool* file_has_alpha) {
  std::vector<uint8_t>* rgb = rgba;
  static const uint8_t kMagic[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (data.size() < 8 || memcmp(data.data(), kMagic, 8) != 0) return false;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(data.data());
  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  bool have_trns = false, seen_iend = false;
  while (pos + 12 <= data.size()) {
    const uint32_t len = be32(p + pos);
    if (pos + 12 + static_cast<size_t>(len) > data.size()) return false;
    const uint8_t* type = p + pos + 4;
    const uint8_t* body = p + pos + 8;
    // libpng treats a CRC mismatch in a critical chunk (upper-case first letter) as an error
    if ((type[0] & 0x20) == 0) {
      const uint32_t want_crc = be32(body + len);
      const uint32_t got_crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), type, static_cast<uInt>(len + 4)));
      if (want_crc != got_crc) return false;
    }
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return false;
      w = be32(body);
      h = be32(body + 4);
      depth = body[8];
      ctype = body[9];
      if (body[10] != 0 || body[11] != 0) return false;
      interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
      have_trns = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      seen_iend = true;
      break;
    }
    pos += 12 + len;
  }
  // Guetzli itself refuses 65536 pixels and more per side (g/jpeg_data_encoder.cc:68): nothing
  // larger is ever decoded, whatever the header claims
  if (!seen_iend || w == 0 || h == 0 || w > 65535u || h > 65535u || interlace > 1) return false;
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return false;
  }
  if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) return false;
  if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) return false;
  if (ctype == 3 && depth > 8) return false;
  const int bits_pp = channels * depth;
  const int bpp = (bits_pp + 7) / 8;

  // size of the filtered image data the header promises (per Adam7 pass when interlaced)
  size_t expect_raw = 0;
  if (!interlace) {
    expect_raw = static_cast<size_t>(h) * ((static_cast<size_t>(w) * bits_pp + 7) / 8 + 1);
  } else {
    static const int xs0[7] = {0, 4, 0, 2, 0, 1, 0}, ys0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int dxs0[7] = {8, 8, 4, 4, 2, 2, 1}, dys0[7] = {8, 8, 8, 4, 4, 2, 2};
    for (int pass = 0; pass < 7; ++pass) {
      const long pw = (static_cast<long>(w) - xs0[pass] + dxs0[pass] - 1) / dxs0[pass];
      const long ph = (static_cast<long>(h) - ys0[pass] + dys0[pass] - 1) / dys0[pass];
      if (pw <= 0 || ph <= 0) continue;
      expect_raw += static_cast<size_t>(ph) * ((static_cast<size_t>(pw) * bits_pp + 7) / 8 + 1);
    }
  }
  // inflate, never past the promised size (a small file must not expand without bound)
  std::vector<uint8_t> raw;
  {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) return false;
    zs.next_in = idat.data();
    zs.avail_in = static_cast<uInt>(idat.size());
    uint8_t buf[1 << 16];
    int rc;
    do {
      zs.next_out = buf;
      zs.avail_out = sizeof(buf);
      rc = inflate(&zs, Z_NO_FLUSH);
      if (rc != Z_OK && rc != Z_STREAM_END) {
        inflateEnd(&zs);
        return false;
      }
      raw.insert(raw.end(), buf, buf + (sizeof(buf) - zs.avail_out));
      if (raw.size() > expect_raw) {
        inflateEnd(&zs);
        return false;
      }
    } while (rc != Z_STREAM_END);
    inflateEnd(&zs);
  }
  if (raw.size() != expect_raw) return false;  // before any image-sized allocation

  // samples[y][x][ch] as 16-bit values at native depth
  std::vector<uint16_t> samples(static_cast<size_t>(w) * h * channels);
  auto put_pass = [&](const std::vector<uint8_t>& px, int pw, int ph, int x0, int y0, int dx, int dy) {
    const size_t rowbytes = (static_cast<size_t>(pw) * bits_pp + 7) / 8;
    for (int y = 0; y < ph; ++y) {
      const uint8_t* row = px.data() + y * rowbytes;
      for (int x = 0; x < pw; ++x) {
        for (int c = 0; c < channels; ++c) {
          const size_t s = static_cast<size_t>(x) * channels + c;
          uint16_t v;
          if (depth == 16) {
            v = static_cast<uint16_t>((row[2 * s] << 8) | row[2 * s + 1]);
          } else if (depth == 8) {
            v = row[s];
          } else {
            const size_t bit = s * depth;
            v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
          }
          samples[((static_cast<size_t>(y0 + y * dy)) * w + (x0 + x * dx)) * channels + c] = v;
        }
      }
    }
  };
  if (!interlace) {
    std::vector<uint8_t> px;
    const size_t rowbytes = (static_cast<size_t>(w) * bits_pp + 7) / 8;
    if (!unfilter(raw.data(), raw.size(), h, rowbytes, bpp, &px)) return false;
    put_pass(px, w, h, 0, 0, 1, 1);
  } else {
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int dxs[7] = {8, 8, 4, 4, 2, 2, 1}, dys[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t off = 0;
    for (int pass = 0; pass < 7; ++pass) {
      const int pw = (static_cast<int>(w) - xs[pass] + dxs[pass] - 1) / dxs[pass];
      const int ph = (static_cast<int>(h) - ys[pass] + dys[pass] - 1) / dys[pass];
      if (pw <= 0 || ph <= 0) continue;
      const size_t rowbytes = (static_cast<size_t>(pw) * bits_pp + 7) / 8;
      std::vector<uint8_t> px;
      if (off > raw.size() || !unfilter(raw.data() + off, raw.size() - off, ph, rowbytes, bpp, &px)) return false;
      off += static_cast<size_t>(ph) * (rowbytes + 1);
      put_pass(px, pw, ph, xs[pass], ys[pass], dxs[pass], dys[pass]);
    }
  }

  // EXPAND / STRIP_16 semantics of libpng, then the reference's channel handling
  *xsize = static_cast<int>(w);
  *ysize = static_cast<int>(h);
  rgb->resize(static_cast<size_t>(4) * w * h);
  *file_has_alpha = ctype == 4 || ctype == 6 || have_trns;
  const int maxv = (1 << depth) - 1;
  auto to8 = [&](uint16_t v) -> uint8_t {
    if (depth == 16) return static_cast<uint8_t>(v >> 8);
    if (depth == 8) return static_cast<uint8_t>(v);
    return static_cast<uint8_t>(v * 255 / maxv);  // 1,2,4-bit gray replicated to 8 bits
  };
  for (size_t i = 0; i < static_cast<size_t>(w) * h; ++i) {
    const uint16_t* s = &samples[i * channels];
    uint8_t r, g, b, a = 255;
    switch (ctype) {
      case 0: {
        r = g = b = to8(s[0]);
        if (have_trns && trns.size() >= 2 && s[0] == ((trns[0] << 8) | trns[1])) a = 0;
        break;
      }
      case 2: {
        r = to8(s[0]);
        g = to8(s[1]);
        b = to8(s[2]);
        if (have_trns && trns.size() >= 6 && s[0] == ((trns[0] << 8) | trns[1]) &&
            s[1] == ((trns[2] << 8) | trns[3]) && s[2] == ((trns[4] << 8) | trns[5]))
          a = 0;
        break;
      }
      case 3: {
        const size_t idx = s[0];
        if (3 * idx + 2 < plte.size()) {
          r = plte[3 * idx];
          g = plte[3 * idx + 1];
          b = plte[3 * idx + 2];
        } else {
          r = g = b = 0;  // libpng: its 256-entry palette is zero-filled beyond PLTE (it only warns)
        }
        if (have_trns && idx < trns.size()) a = trns[idx];
        break;
      }
      case 4: {
        r = g = b = to8(s[0]);
        a = to8(s[1]);
        break;
      }
      default: {
        r = to8(s[0]);
        g = to8(s[1]);
        b = to8(s[2]);
        a = to8(s[3]);
        break;
      }
    }
    (*rgb)[4 * i + 0] = r;
    (*rgb)[4 * i + 1] = g;
    (*rgb)[4 * i + 2] = b;
    (*rgb)[4 * i + 3] = a;
  }
  return true;
}

// guetzli's reader (g/guetzli.cc:43-152): RGB, transparent pixels blended on black.
inline bool ReadPNG(const std::string& data, int* xsize, int* ysize, std::vector<uint8_t>* rgb) {
  std::vector<uint8_t> rgba;
  bool has_alpha = false;
  if (!ReadPNGRGBA(data, xsize, ysize, &rgba, &has_alpha)) return false;
  const size_t n = static_cast<size_t>(*xsize) * *ysize;
  rgb->resize(3 * n);
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c)
      (*rgb)[3 * i + c] = has_alpha ? BlendOnBlack(rgba[4 * i + c], rgba[4 * i + 3]) : rgba[4 * i + c];
  return true;
}

}  // namespace gb200_cli
