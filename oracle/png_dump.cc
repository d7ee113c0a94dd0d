// Test helper: decodes a PNG with the CLI's reader and writes raw RGB (w h as two
// int32 then 3*w*h bytes) to stdout.  tests/test_cli.py compares it with PIL.
#include <stdio.h>

#include <string>
#include <vector>

#include "../guetzli_b200/cli/png_reader.h"

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::string data;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
  fclose(f);
  int w, h;
  std::vector<unsigned char> rgb;
  if (!gb200_cli::ReadPNG(data, &w, &h, &rgb)) return 1;
  fwrite(&w, 4, 1, stdout);
  fwrite(&h, 4, 1, stdout);
  fwrite(rgb.data(), 1, rgb.size(), stdout);
  return 0;
}
