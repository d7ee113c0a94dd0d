#!/bin/sh
# Known answers for all 64 images of BASELINE configs[4] (6 reference processes in parallel).
cd "$(dirname "$0")"
seq 1001 1063 | xargs -P 6 -I{} python make_golden_large.py gradnoise1024_s{}_q84
