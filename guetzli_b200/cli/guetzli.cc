// Drop-in `guetzli` command line (guetzli/guetzli.cc:232-326): same flags, same
// exit codes, same messages; the work goes through guetzli::Process(RGB) of
// include/guetzli_b200_compat.h (C ABI of libguetzli_b200.so).  JPEG input (4:4:4)
// goes through guetzli::Process(jpeg bytes) of the same header.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#include "guetzli_b200_compat.h"
#include "png_reader.h"

namespace {

constexpr int kDefaultJPEGQuality = 95;
constexpr int kBytesPerPixel = 350;
constexpr int kLowestMemusageMB = 100;
constexpr int kDefaultMemlimitMB = 6000;

std::string ReadFileOrDie(const char* filename) {
  const bool read_from_stdin = strncmp(filename, "-", 2) == 0;
  FILE* f = read_from_stdin ? stdin : fopen(filename, "rb");
  if (!f) {
    perror("Can't open input file");
    exit(1);
  }
  std::string result;
  char buf[1 << 16];
  for (;;) {
    const size_t n = fread(buf, 1, sizeof(buf), f);
    if (ferror(f)) {
      perror("fread");
      exit(1);
    }
    result.append(buf, n);
    if (n == 0 || feof(f)) break;
  }
  if (!read_from_stdin) fclose(f);
  return result;
}

void WriteFileOrDie(const char* filename, const std::string& contents) {
  const bool write_to_stdout = strncmp(filename, "-", 2) == 0;
  FILE* f = write_to_stdout ? stdout : fopen(filename, "wb");
  if (!f) {
    perror("Can't open output file for writing");
    exit(1);
  }
  if (fwrite(contents.data(), 1, contents.size(), f) != contents.size()) {
    perror("fwrite");
    exit(1);
  }
  if (fclose(f) < 0) {
    perror("fclose");
    exit(1);
  }
}

void TerminateHandler() {
  fprintf(stderr,
          "Unhandled exception. Most likely insufficient memory available.\n"
          "Make sure that there is 300MB/MPix of memory available.\n");
  exit(1);
}

void Usage() {
  fprintf(stderr,
          "Guetzli JPEG compressor. Usage: \n"
          "guetzli [flags] input_filename output_filename\n"
          "\n"
          "Flags:\n"
          "  --verbose    - Print a verbose trace of all attempts to standard output.\n"
          "  --quality Q  - Visual quality to aim for, expressed as a JPEG quality value.\n"
          "                 Default value is %d.\n"
          "  --memlimit M - Memory limit in MB. Guetzli will fail if unable to stay under\n"
          "                 the limit. Default limit is %d MB.\n"
          "  --nomemlimit - Do not limit memory usage.\n",
          kDefaultJPEGQuality, kDefaultMemlimitMB);
  exit(1);
}

}  // namespace

int main(int argc, char** argv) {
  std::set_terminate(TerminateHandler);
  int verbose = 0;
  int quality = kDefaultJPEGQuality;
  int memlimit_mb = kDefaultMemlimitMB;
  int opt_idx = 1;
  for (; opt_idx < argc; opt_idx++) {
    if (strnlen(argv[opt_idx], 2) < 2 || argv[opt_idx][0] != '-' || argv[opt_idx][1] != '-') break;
    if (!strcmp(argv[opt_idx], "--verbose")) {
      verbose = 1;
    } else if (!strcmp(argv[opt_idx], "--quality")) {
      opt_idx++;
      if (opt_idx >= argc) Usage();
      quality = atoi(argv[opt_idx]);
    } else if (!strcmp(argv[opt_idx], "--memlimit")) {
      opt_idx++;
      if (opt_idx >= argc) Usage();
      memlimit_mb = atoi(argv[opt_idx]);
    } else if (!strcmp(argv[opt_idx], "--nomemlimit")) {
      memlimit_mb = -1;
    } else if (!strcmp(argv[opt_idx], "--")) {
      opt_idx++;
      break;
    } else {
      fprintf(stderr, "Unknown commandline flag: %s\n", argv[opt_idx]);
      Usage();
    }
  }
  if (argc - opt_idx != 2) Usage();

  std::string in_data = ReadFileOrDie(argv[opt_idx]);
  std::string out_data;
  guetzli::Params params;
  params.butteraugli_target = static_cast<float>(guetzli::ButteraugliScoreForQuality(quality));
  guetzli::ProcessStats stats;
  if (verbose) stats.debug_output_file = stderr;

  static const unsigned char kPNGMagicBytes[] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (in_data.size() >= 8 && memcmp(in_data.data(), kPNGMagicBytes, sizeof(kPNGMagicBytes)) == 0) {
    int xsize, ysize;
    std::vector<uint8_t> rgb;
    if (!gb200_cli::ReadPNG(in_data, &xsize, &ysize, &rgb)) {
      fprintf(stderr, "Error reading PNG data from input file\n");
      return 1;
    }
    const double pixels = static_cast<double>(xsize) * ysize;
    if (memlimit_mb != -1 &&
        (pixels * kBytesPerPixel / (1 << 20) > memlimit_mb || memlimit_mb < kLowestMemusageMB)) {
      fprintf(stderr, "Memory limit would be exceeded. Failing.\n");
      return 1;
    }
    if (!guetzli::Process(params, &stats, rgb, xsize, ysize, &out_data)) {
      fprintf(stderr, "Guetzli processing failed\n");
      return 1;
    }
  } else {
    int xsize = 0, ysize = 0;
    if (!gb200_jpeg_dimensions(reinterpret_cast<const uint8_t*>(in_data.data()), in_data.size(), &xsize, &ysize)) {
      fprintf(stderr, "Error reading JPG data from input file\n");
      return 1;
    }
    const double pixels = static_cast<double>(xsize) * ysize;
    if (memlimit_mb != -1 &&
        (pixels * kBytesPerPixel / (1 << 20) > memlimit_mb || memlimit_mb < kLowestMemusageMB)) {
      fprintf(stderr, "Memory limit would be exceeded. Failing.\n");
      return 1;
    }
    if (!guetzli::Process(params, &stats, in_data, &out_data)) {
      fprintf(stderr, "Guetzli processing failed\n");
      return 1;
    }
  }
  WriteFileOrDie(argv[opt_idx + 1], out_data);
  return 0;
}
