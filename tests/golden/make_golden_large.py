#!/usr/bin/env python3
"""Known answers of the UNMODIFIED reference for the BASELINE.json full-size
configurations (minutes to hours of CPU each; run in the background where
/root/reference exists).  Appends to tests/golden/golden_large.json."""
import fcntl
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import reflib  # noqa: E402
from guetzli_b200 import synth  # noqa: E402

CASES = {
    "noise1080p_s1234_q95": (lambda: synth.noise(1080, 1920, 1234), 95),       # BASELINE configs[1]
    "gradnoise4k_s4321_q90": (lambda: synth.gradnoise(2160, 3840, 4321), 90),  # BASELINE configs[2]
    "gradnoise1024_s1000_q84": (lambda: synth.gradnoise(1024, 1024, 1000), 84),  # configs[4], image 0
    "gradnoise8k_s8192_q95": (lambda: synth.gradnoise(4320, 7680, 8192), 95),   # BASELINE configs[3] (tiled)
}
# BASELINE configs[4]: the other 63 images of the batch (seed 1000+i)
for _i in range(1, 64):
    CASES["gradnoise1024_s%d_q84" % (1000 + _i)] = ((lambda s: (lambda: synth.gradnoise(1024, 1024, s)))(1000 + _i), 84)

name = sys.argv[1]
gen, q = CASES[name]
rgb = gen()
ok, jpeg, trace, counters, secs = reflib.process_rgb(rgb, q)
path = os.path.join(HERE, "golden_large.json")
with open(path + ".lock", "w") as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    out = json.load(open(path)) if os.path.exists(path) else {}
    out[name] = {"quality": q, "shape": list(rgb.shape), "input_sha256": synth.sha256(rgb), "ok": ok,
                 "jpeg_sha256": hashlib.sha256(jpeg).hexdigest(), "jpeg_size": len(jpeg),
                 "trace_sha256": hashlib.sha256(trace.encode()).hexdigest(), "iterations": counters,
                 "ref_seconds_here": round(secs, 1)}
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print(name, out[name])
