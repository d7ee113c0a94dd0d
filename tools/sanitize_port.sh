#!/bin/bash
# AddressSanitizer + UBSan run of the host-side code (search, JPEG reader/writer, order replay,
# strip mode, the kernel bodies as host loops) through the CPU port.  Test infrastructure only.
#   tools/sanitize_port.sh            (from the repository root; ~1 min)
# shift-base is excluded: the integer DCT shifts negative values left exactly like the
# reference (guetzli/fdct.cc), which every targeted compiler defines as arithmetic.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${OUT:-/tmp/gb200_asan}
mkdir -p "$OUT"
make -s -C "$ROOT/oracle" -f Makefile.port OUT="$OUT" \
  CXXFLAGS="-O1 -g -std=c++17 -fPIC -ffp-contract=off -fopenmp -DGB200_HOSTSIM -I../guetzli_b200/csrc -I../include -fsanitize=address,undefined -fno-sanitize=shift-base -fno-omit-frame-pointer" \
  "$OUT/libguetzli_port.so"
ASAN=$(ls /usr/lib/x86_64-linux-gnu/libasan.so.? | head -1)
UBSAN=$(ls /usr/lib/x86_64-linux-gnu/libubsan.so.? | head -1)
cd "$ROOT"
GB200_SAN_LIB="$OUT/libguetzli_port.so" GB200_DEVICE_ORDER=check GB200_ORDER_HOST_RANGE=256 \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
LD_PRELOAD="$ASAN $UBSAN" PYTHONPATH="$ROOT" python tools/sanitize_port.py
