#!/bin/bash
# Two GPUs: smoke, the NCCL strip-mode test, the bench under torchrun (batch per rank + the 8K image over both ranks).
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
(time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tiled_nccl or strip_mode") > gpurun_out/pytest_n2.log 2>&1; tail -4 gpurun_out/pytest_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n2.json"))
    print("N2", d["value"], d["e2e"]["value"], d["tiled"])
except Exception as e:
    print("N2 failed", e)
PY
