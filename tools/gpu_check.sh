#!/bin/bash
# One GPU call: parity tests, a short bench, the one-thread-per-unit JPEG kernels for comparison.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 400 gpurun_out/bench_a.err
GB200_JPEG=thread timeout 120 python bench.py --batch 1 --steps 1 --warmup 1 --no-tiled-leg --no-cpu-baseline > gpurun_out/bench_b1_thread.json 2> gpurun_out/bench_b1_thread.err
timeout 120 python bench.py --batch 1 --steps 1 --warmup 1 --no-tiled-leg --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
python - <<'PY'
import json
for f in ["bench_a", "bench_b1_thread", "bench_b1"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["single_image"]["ms"], d["single_image"]["breakdown_ms"], d["single_image_gpu_kernel_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
