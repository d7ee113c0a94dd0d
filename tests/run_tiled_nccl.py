#!/usr/bin/env python3
"""Multi-GPU check of the row-strip mode (one image tiled over the GPUs of a node,
NCCL exchange).  Launch with torchrun, one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tests/run_tiled_nccl.py [workload]

Every rank encodes the same image collectively; the bytes must equal the golden
answer of the reference (or, for 'noise1080p'/'gradnoise4k', golden_large.json).
Prints one line per rank-0 with timing vs the untiled single-GPU run."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import guetzli_b200 as gb  # noqa: E402
import parity  # noqa: E402
from guetzli_b200 import synth  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "bees"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo")  # only carries the NCCL id of the library's own communicator
    if what == "bees":
        g = parity.GOLDEN["bees_444x258_q95"]
        rgb = parity.golden_input("bees_444x258_q95")
    else:
        large = json.load(open(os.path.join(HERE, "golden", "golden_large.json")))
        key = {"noise1080p": "noise1080p_s1234_q95", "gradnoise4k": "gradnoise4k_s4321_q90",
               "gradnoise1024": "gradnoise1024_s1000_q84"}[what]
        g = large[key]
        rgb = {"noise1080p": lambda: synth.noise(1080, 1920, 1234),
               "gradnoise4k": lambda: synth.gradnoise(2160, 3840, 4321),
               "gradnoise1024": lambda: synth.gradnoise(1024, 1024, 1000)}[what]()
    h, w, _ = rgb.shape
    box = [gb.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    gb.dist_init(box[0], rank, world, local)
    params = gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(g["quality"]))
    dist.barrier()
    t0 = time.perf_counter()
    st = gb.ProcessStats()
    ok, jpeg = gb.process_tiled(params, st, rgb, w, h)
    t_tiled = time.perf_counter() - t0
    sha = hashlib.sha256(jpeg).hexdigest()
    assert ok and sha == g["jpeg_sha256"], f"rank {rank}: tiled output differs from the reference ({len(jpeg)} bytes)"
    dist.barrier()
    if rank == 0:
        t0 = time.perf_counter()
        st1 = gb.ProcessStats()
        ok1, jpeg1 = gb.process(params, st1, rgb, w, h, device=local)
        t_single = time.perf_counter() - t0
        assert jpeg1 == jpeg
        print(json.dumps({"tiled_check": what, "world": world, "bit_exact_vs_reference": True,
                          "seconds_tiled": round(t_tiled, 3), "seconds_single_gpu": round(t_single, 3),
                          "ms_compare_tiled": round(st.device["ms_compare"], 1),
                          "ms_compare_single": round(st1.device["ms_compare"], 1),
                          "iterations": st.counters["number of iterations"]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
