#!/bin/bash
# Final single-GPU evidence of the round: parity tests, bench, launch list, batch-size probe.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 420 python bench.py --gpus 1 --steps 8 --warmup 4 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 1 --batch 1 --no-tiled-leg --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 200 python bench.py --batch 24 --steps 2 --warmup 2 --no-tiled-leg --no-cpu-baseline > gpurun_out/bench_b24.json 2> gpurun_out/bench_b24.err
python - <<'PY'
import json
for f in ["bench_final", "bench_b24"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["single_image"]["ms"], d["single_image"]["breakdown_ms"], d["single_image_gpu_kernel_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
