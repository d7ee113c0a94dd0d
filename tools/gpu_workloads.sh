#!/bin/bash
# The other BASELINE configurations with the final code: 4K alone and 8 in flight, 1024^2 q84 with 16 in flight.
mkdir -p gpurun_out
F="--no-tiled-leg --no-cpu-baseline"
timeout 150 python bench.py --workload gradnoise4k_q90 --batch 1 --steps 2 --warmup 1 $F > gpurun_out/w_4k_b1.json 2> gpurun_out/w_4k_b1.err
timeout 200 python bench.py --workload gradnoise4k_q90 --batch 8 --steps 2 --warmup 2 $F > gpurun_out/w_4k_b8.json 2> gpurun_out/w_4k_b8.err
timeout 150 python bench.py --workload gradnoise1024_q84 --batch 16 --steps 4 --warmup 3 $F > gpurun_out/w_1024_b16.json 2> gpurun_out/w_1024_b16.err
python - <<'PY'
import json
for f in ["w_4k_b1", "w_4k_b8", "w_1024_b16"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["single_image"]["ms"], d["config"].get("output_matches_reference_golden"))
    except Exception as e:
        print(f, "failed", e)
PY
