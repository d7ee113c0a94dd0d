#!/usr/bin/env python3
"""Headline benchmark: MPix/s of guetzli::Process(RGB) (bit-exact JPEG) on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...          # reference CPU arm (oracle/_ref)

A "step" = one full Process() of one synthetic image per GPU (BASELINE.json
configs[1]: 1920x1080 sRGB noise, --quality 95, unless --workload says otherwise).
Multi-GPU = independent images sharded over ranks (weak scaling, no data-path
collective; torch.distributed only for the barrier and the max-over-ranks time).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] -- the single-GPU configuration the metric is quoted on
    "noise1080p_q95": dict(gen="noise", h=1080, w=1920, seed=1234, quality=95,
                           cpu_sample=dict(gen="noise", h=160, w=160, seed=1234)),
    "gradnoise4k_q90": dict(gen="gradnoise", h=2160, w=3840, seed=4321, quality=90,
                            cpu_sample=dict(gen="gradnoise", h=320, w=320, seed=4321)),
    "gradnoise1024_q84": dict(gen="gradnoise", h=1024, w=1024, seed=1000, quality=84,
                              cpu_sample=dict(gen="gradnoise", h=384, w=384, seed=1000)),
    "gradnoise512_q90": dict(gen="gradnoise", h=512, w=512, seed=4321, quality=90,
                             cpu_sample=dict(gen="gradnoise", h=256, w=256, seed=4321)),
    "noise512_q95": dict(gen="noise", h=512, w=512, seed=1234, quality=95,
                         cpu_sample=dict(gen="noise", h=160, w=160, seed=1234)),
}

# Algorithmic bytes per launched element (pixel of one plane, or 8x8 block) of the
# staged v1 kernels: compulsory reads + writes of that stage ("kernels").
ALG_BYTES = {
    "malta_channel": 28, "blur_x": 8, "blur_y": 8, "sub_planes": 12, "opsin_px": 36, "split_mf_hf": 44, "split_hf_uhf": 68,
    "malta_pre": 12, "malta_acc_hf": 8, "malta_acc_lf": 12, "noise_pre": 12, "noise_asym_acc": 20,
    "mask_diff_pre": 40, "combine_sqrt": 44, "diffmap_mix": 12, "render_blocks": 1152,
    "block_max": 260, "jpeg_unit_bits": 128, "jpeg_emit": 140, "jpeg_hist_acc": 128, "jpeg_hist": 128, "linearize_rgb": 15, "quantize_coeffs": 4, "fdct_blocks": 576,
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu
# --set full captures (profiles/), bytes; None until captured.
NCU_TRAFFIC = {
    # profiles/r01_final2_ncu_full.csv (k_malta_pre3 + k_malta_sums) and r01_ncu_full_malta_blur.csv
    # (noise1080p, per launch: read + write)
    "malta_channel": 49.8e6 + 1.7e6 + 24.9e6,
    "blur_x": 8.33e6, "blur_y": 8.32e6,
}


def make_image(spec, rank=0):
    from guetzli_b200 import synth
    from guetzli_b200.distributed import image_seed
    seed = image_seed(spec["seed"], rank)
    if spec["gen"] == "noise":
        return synth.noise(spec["h"], spec["w"], seed)
    return synth.gradnoise(spec["h"], spec["w"], seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    from guetzli_b200 import distributed as gdist
    return gdist.setup("nccl")


def barrier_sync(dist):
    from guetzli_b200 import distributed as gdist
    gdist.barrier(dist, cuda=True)


def max_over_ranks(dist, seconds, local):
    from guetzli_b200 import distributed as gdist
    return gdist.max_over_ranks(dist, seconds, device=f"cuda:{local}")


def cpu_reference_seconds(spec, quality):
    """One reference Process() on one core -> (seconds, pixels, sha of output)."""
    import reflib
    rgb = make_image(spec)
    ok, jpeg, _, counters, secs = reflib.process_rgb(rgb, quality, trace=False)
    assert ok
    return secs, rgb.shape[0] * rgb.shape[1], hashlib.sha256(jpeg).hexdigest(), counters


def _ref_worker(args):
    spec, quality, rank = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    rgb = make_image(spec, rank)
    ok, jpeg, _, _, secs = reflib.process_rgb(rgb, quality, trace=False)
    return secs


def run_reference(args, wl, name):
    """Reference arm: the unmodified reference (oracle/_ref) on the host cores,
    one single-threaded Process() per core (the reference's own parallelism idiom,
    tests/golden_test.sh:25), each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import reflib
    if not reflib.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libguetzli_ref.so not built"}))
        return
    spec = wl["cpu_sample"]
    cores = max(1, min(os.cpu_count() or 1, 64))
    px = spec["h"] * spec["w"]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, [(spec, wl["quality"], r) for r in range(cores)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, [(spec, wl["quality"], r) for r in range(cores)])
        dt = time.perf_counter() - t0
    value = cores * args.steps * px / dt / 1e6
    sample = f"{spec['gen']}({spec['h']},{spec['w']},seed {spec['seed']}+core) q{wl['quality']}, one image per core per step"
    line = {
        "impl": "reference", "metric": "MPix/s (bit-exact JPEG, guetzli::Process)", "value": value,
        "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/f64+int16 (CPU)", "data": "synthetic",
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="noise1080p_q95", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=16,
                    help="images per GPU per step, encoded concurrently (one host thread + CUDA stream each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, wl, args.workload)
        return

    from concurrent.futures import ThreadPoolExecutor
    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    import guetzli_b200 as gb
    lib = gb.load_library()
    if lib.gb200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    M = args.batch
    # weak scaling: rank r encodes images r*M .. r*M+M-1 of the generator, every step
    images = [make_image(wl, rank * M + j) for j in range(M)]
    h, w, _ = images[0].shape
    params = gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(wl["quality"]))
    px = h * w
    pool = ThreadPoolExecutor(M)
    shas = [set() for _ in range(M)]

    def encode_host(j):  # reference-facing call: host buffer in, JPEG bytes out
        st = gb.ProcessStats()
        ok, jpeg = gb.process(params, st, images[j], w, h, device=local)
        assert ok
        shas[j].add(hashlib.sha256(jpeg).hexdigest())
        return st

    def encode_resident(args_):  # image already uploaded
        j, img = args_
        st = gb.ProcessStats()
        ok, jpeg = img.process(params, st)
        assert ok
        shas[j].add(hashlib.sha256(jpeg).hexdigest())
        return st

    for _ in range(args.warmup):
        list(pool.map(encode_host, range(M)))

    # ---- value: images already resident in HBM when the timed region starts ----
    resident = [[gb.DeviceImage(images[j], device=local, prepare=False) for j in range(M)]
                for _ in range(args.steps)]
    sampler = ClockSampler(local)
    barrier_sync(dist)
    sampler.start()
    n0, _, _ = gb.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    stats_list = []
    for step in range(args.steps):
        stats_list = list(pool.map(encode_resident, list(enumerate(resident[step]))))
    ev1.record()
    barrier_sync(dist)
    clocks = sampler.stop()
    n1, _, _ = gb.counters()
    dt = max_over_ranks(dist, ev0.elapsed_time(ev1) / 1e3, local)
    launches = n1 - n0
    for row in resident:
        for img in row:
            img.close()
    value = world * args.steps * M * px / dt / 1e6

    # ---- e2e: same job through the reference-facing call with HOST buffers ----
    barrier_sync(dist)
    _, a0, b0 = gb.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        list(pool.map(encode_host, range(M)))
    e1.record()
    barrier_sync(dist)
    _, a1, b1 = gb.counters()
    dt_e2e = max_over_ranks(dist, e0.elapsed_time(e1) / 1e3, local)
    e2e_value = world * args.steps * M * px / dt_e2e / 1e6
    h2d, d2h = (a1 - a0) / args.steps, (b1 - b0) / args.steps
    assert all(len(x) == 1 for x in shas), "non-deterministic output"

    # ---- per-kernel CUDA-event times: one more image, alone on the GPU ----------
    kernels, gpu_ms, prof_st = [], 0.0, None
    if rank == 0:
        import ctypes as C
        lib.gb200_profile_reset()
        lib.gb200_profile_enable(1)
        img = gb.DeviceImage(images[0], device=local, prepare=False)
        prof_st = encode_resident((0, img))
        img.close()
        lib.gb200_profile_enable(0)
        cap = 64
        names = ((C.c_char * 48) * cap)()
        kl = (C.c_long * cap)()
        kms = (C.c_double * cap)()
        kel = (C.c_double * cap)()
        nk = lib.gb200_profile_get(names, kl, kms, kel, cap)
        for i in range(min(nk, cap)):
            kernels.append({"name": names[i].value.decode(), "launches": kl[i], "ms": kms[i], "elements": kel[i]})
        kernels.sort(key=lambda k: -k["ms"])
        gpu_ms = sum(k["ms"] for k in kernels)
    barrier_sync(dist)
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_kind = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_kind = 6650.0, "fallback (B200_PROFILING.md)"
    roofline = None
    ranked = [k for k in kernels if k["name"] in ALG_BYTES and k["launches"] > 1]
    if ranked:
        top = ranked[0]
        bpe = ALG_BYTES[top["name"]]
        achieved = bpe * top["elements"] / (top["ms"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(top["name"]),
                    "peak_source": peak_kind, "avg_launch_us": top["ms"] / max(1, top["launches"]) * 1e3,
                    "share_of_gpu_time": top["ms"] / gpu_ms if gpu_ms else None,
                    "alg_bytes_per_element": bpe,
                    "measured_on": "one extra image encoded alone (single stream) after the timed region"}
    st = stats_list[-1]
    line = {
        "metric": "MPix/s (bit-exact JPEG, guetzli::Process)", "value": value, "unit": "MPix/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (+f64 sub-expressions), int16/int32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_gpu": M,
                   "image": f"{wl['gen']}({h},{w},seed {wl['seed']}+rank*{M}+j), j<{M}",
                   "quality": wl["quality"], "input_sha256_rank0_img0": hashlib.sha256(images[0].tobytes()).hexdigest(),
                   "output_sha256_rank0_img0": sorted(shas[0])[0],
                   "iterations_img_last": st.counters["number of iterations"],
                   "sharding": f"{world} GPU(s) x {M} independent images per step, one host thread + stream each",
                   "l2_policy": "working set of one step (>= 16 images x 45 float planes) exceeds L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "MPix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": dt_e2e / args.steps * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "single_image_ms": {k: round(prof_st.device[k], 1) for k in
                            ("ms_total", "ms_compare", "ms_zeroing", "ms_jpeg", "ms_sort", "ms_walk",
                             "order_partial", "order_exact")},
        "single_image_gpu_kernel_ms": round(gpu_ms, 2),
        "top_kernels": [{"name": k["name"], "ms": round(k["ms"], 2), "launches": k["launches"]} for k in kernels[:30]],
    }
    if world == 1 and not args.no_cpu_baseline:
        import reflib
        if reflib.available():
            spec = wl["cpu_sample"]
            secs, spx, sha, counters = cpu_reference_seconds(spec, wl["quality"])
            line["cpu_baseline"] = {
                "value": spx / secs / 1e6, "unit": "MPix/s", "cores": 1, "kind": "reference",
                "sample": f"{spec['gen']}({spec['h']},{spec['w']},seed {spec['seed']}) q{wl['quality']}, "
                          f"{secs:.1f} s, {counters[0]} iterations (crop-sized sample of the same generator)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
