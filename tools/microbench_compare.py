#!/usr/bin/env python
"""Per-kernel CUDA-event times of N Compare passes (a9+a10) on one device-resident
image: a development aid for A/B-ing kernel variants (e.g. GB200_MALTA=0|1|2) in a
single GPU call.  Prints one JSON line: wall ms per Compare, distance, kernel table."""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

import guetzli_b200 as gb  # noqa: E402
from guetzli_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--tag", default="")
    ap.add_argument("--zeroing", action="store_true", help="also run the zeroing-order kernel (a14) once")
    args = ap.parse_args()
    lib = gb.load_library()
    rgb = synth.noise(args.h, args.w, 1234)
    img = gb.DeviceImage(rgb, device=0)
    q = np.full(192, 3, dtype=np.int32)
    img.apply_global_quant(q)
    d = img.compare()  # warm-up
    if args.zeroing:
        img.zeroing_orders(1.0)
    lib.gb200_profile_reset()
    lib.gb200_profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(args.n):
        d = img.compare()
    wall = (time.perf_counter() - t0) / args.n * 1e3
    lib.gb200_profile_enable(0)
    cap = 64
    names = ((C.c_char * 48) * cap)()
    kl, kms, kel = (C.c_long * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
    nk = lib.gb200_profile_get(names, kl, kms, kel, cap)
    ks = [(names[i].value.decode(), kl[i], kms[i]) for i in range(min(nk, cap))]
    ks.sort(key=lambda k: -k[2])
    total = sum(k[2] for k in ks)
    print(json.dumps({"tag": args.tag, "wall_ms_per_compare": round(wall, 3), "kernel_ms_per_compare": round(total / args.n, 3),
                      "distance": d, "sha_dist": gb.synth.sha256(img.distmap())[:16],
                      "kernels_us_per_compare": {k[0]: round(k[2] / args.n * 1e3, 1) for k in ks[:14]}}))
    img.close()


if __name__ == "__main__":
    main()
