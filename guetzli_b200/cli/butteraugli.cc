// Stand-alone `butteraugli` comparison tool on the device kernels (scope row f4):
//   butteraugli image1.png image2.png [heatmap.ppm]
// Interface, messages, score format and heat map of
// third_party/butteraugli/butteraugli/butteraugli_main.cc:362-455; the distance map
// comes from gb200_butteraugli_diffmap (ButteraugliInterface, butteraugli.cc:1858).
// PNG input only: the reference decodes JPEG with libjpeg, whose IDCT this repo does
// not reproduce, so JPEG arguments are refused instead of scored differently.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "guetzli_b200.h"
#include "png_reader.h"

namespace {

struct Picture {
  int w = 0, h = 0;
  bool has_alpha = false;
  std::vector<uint8_t> rgba;
};

bool ReadAll(const char* name, std::string* out) {
  FILE* f = fopen(name, "rb");
  if (!f) {
    fprintf(stderr, "Cannot open %s\n", name);
    return false;
  }
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

Picture ReadImageOrDie(const char* name) {
  std::string data;
  if (!ReadAll(name, &data)) exit(1);
  if (data.size() < 2) {
    fprintf(stderr, "Cannot read from %s\n", name);
    exit(1);
  }
  Picture p;
  if (static_cast<uint8_t>(data[0]) == 0xff && static_cast<uint8_t>(data[1]) == 0xd8) {
    fprintf(stderr, "File %s is a JPEG; this build scores PNG input only (no libjpeg decoder).\n", name);
    exit(1);
  }
  if (!gb200_cli::ReadPNGRGBA(data, &p.w, &p.h, &p.rgba, &p.has_alpha)) {
    fprintf(stderr, "File %s is neither a valid JPEG nor a valid PNG.\n", name);
    exit(1);
  }
  return p;
}

// butteraugli_main.cc:137 / :239: sRGB byte -> linear 0..255; transparent pixels are laid
// over `background` in sRGB space first.
void ToLinear(const Picture& p, int background, std::vector<float>* planes) {
  static double table[256];
  static bool ready = false;
  if (!ready) {
    for (int i = 0; i < 256; ++i) {
      const double srgb = i / 255.0;
      table[i] = 255.0 * (srgb <= 0.04045 ? srgb / 12.92 : pow((srgb + 0.055) / 1.055, 2.4));
    }
    ready = true;
  }
  const size_t n = static_cast<size_t>(p.w) * p.h;
  planes->resize(3 * n);
  for (int c = 0; c < 3; ++c)
    for (size_t i = 0; i < n; ++i) {
      int v = p.rgba[4 * i + c];
      if (p.has_alpha) {
        const int a = p.rgba[4 * i + 3];
        if (a == 0) {
          v = background;
        } else if (a != 255) {
          v = (v * a + background * (255 - a) + 127) / 255;
        }
      }
      (*planes)[c * n + i] = static_cast<float>(table[v]);
    }
}

// butteraugli.cc:1902 / :1923
double FuzzyClass(double score) {
  const double width_up = 6.07887388532, width_down = 5.50793514384, m0 = 2.0, scaler = 0.840253347958;
  if (score < 1.0) {
    double v = m0 / (1.0 + exp((score - 1.0) * width_down));
    v -= 1.0;
    v *= 2.0 - scaler;
    return v + scaler;
  }
  return m0 / (1.0 + exp((score - 1.0) * width_up)) * scaler;
}
double FuzzyInverse(double seek) {
  double pos = 0;
  for (double range = 1.0; range >= 1e-10; range *= 0.5) pos += FuzzyClass(pos) < seek ? -range : range;
  return pos;
}

// butteraugli_main.cc:311
void ScoreToRgb(double score, double good, double bad, uint8_t rgb[3]) {
  static const double kMap[12][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0},       {1, 1, 0},       {1, 0, 0},
                                     {1, 0, 1}, {0.5, 0.5, 1.0}, {1.0, 0.5, 0.5}, {1.0, 1.0, 0.5}, {1, 1, 1}, {1, 1, 1}};
  if (score < good) {
    score = (score / good) * 0.3;
  } else if (score < bad) {
    score = 0.3 + (score - good) / (bad - good) * 0.15;
  } else {
    score = 0.45 + (score - bad) / (bad * 12) * 0.5;
  }
  score = std::min<double>(std::max<double>(score * 11, 0.0), 10);
  const int ix = static_cast<int>(score);
  const double mix = score - ix;
  for (int i = 0; i < 3; ++i) {
    const double v = mix * kMap[ix + 1][i] + (1 - mix) * kMap[ix][i];
    rgb[i] = static_cast<uint8_t>(255 * pow(v, 0.5) + 0.5);
  }
}

bool Compare(const Picture& a, const Picture& b, int background, int device, std::vector<float>* map, double* value) {
  std::vector<float> l0, l1;
  ToLinear(a, background, &l0);
  ToLinear(b, background, &l1);
  map->resize(static_cast<size_t>(a.w) * a.h);
  if (!gb200_butteraugli_diffmap(l0.data(), l1.data(), a.w, a.h, device, map->data(), value)) {
    fprintf(stderr, "Butteraugli comparison failed\n");
    if (*gb200_last_error()) fprintf(stderr, "%s\n", gb200_last_error());
    return false;
  }
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc != 3 && argc != 4) {
    fprintf(stderr, "Usage: %s {image1.(png|jpg|jpeg)} {image2.(png|jpg|jpeg)} [heatmap.ppm]\n", argv[0]);
    return 1;
  }
  const Picture a = ReadImageOrDie(argv[1]);
  const Picture b = ReadImageOrDie(argv[2]);
  if (a.has_alpha != b.has_alpha) {
    fprintf(stderr, "Different number of channels: %lu vs %lu\n", a.has_alpha ? 4ul : 3ul, b.has_alpha ? 4ul : 3ul);
    return 1;
  }
  if (a.w != b.w || a.h != b.h) {
    fprintf(stderr, "The images are not equal in size: (%lu,%lu) vs (%lu,%lu)\n", static_cast<unsigned long>(a.w),
            static_cast<unsigned long>(b.w), static_cast<unsigned long>(a.h), static_cast<unsigned long>(b.h));
    return 1;
  }
  int device = 0;
  if (const char* e = getenv("GUETZLI_B200_DEVICE")) device = atoi(e);
  std::vector<float> map, map_white;
  double value = 0;
  if (!Compare(a, b, 0, device, &map, &value)) return 1;
  const std::vector<float>* best = &map;
  if (a.has_alpha) {  // also over a white background; the worse of the two counts
    double value_white = 0;
    if (!Compare(a, b, 255, device, &map_white, &value_white)) return 1;
    if (value_white > value) {
      value = value_white;
      best = &map_white;
    }
  }
  printf("%lf\n", value);
  if (argc == 4) {
    const double good = FuzzyInverse(1.5), bad = FuzzyInverse(0.5);
    std::vector<uint8_t> rgb(3 * best->size());
    for (size_t i = 0; i < best->size(); ++i) ScoreToRgb((*best)[i], good, bad, &rgb[3 * i]);
    FILE* f = fopen(argv[3], "wb");
    if (f == NULL) {
      fprintf(stderr, "Cannot open %s\n", argv[3]);
      perror("fopen");
      return 1;
    }
    bool ok = fprintf(f, "P6\n%lu %lu\n255\n", static_cast<unsigned long>(a.w), static_cast<unsigned long>(a.h)) >= 0;
    ok = ok && fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
    if (fclose(f) != 0) ok = false;
    if (!ok) return 1;
  }
  return 0;
}
