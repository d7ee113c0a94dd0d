#!/bin/bash
# A/B of the overlapped iteration tail (GB200_OVERLAP) on one image, plus ncu --set full of the size pass.
mkdir -p gpurun_out
B="python bench.py --batch 1 --steps 2 --warmup 1 --no-tiled-leg --no-cpu-baseline"
timeout 120 $B > gpurun_out/ab_overlap1.json 2> gpurun_out/ab_overlap1.err
GB200_OVERLAP=0 timeout 120 $B > gpurun_out/ab_overlap0.json 2> gpurun_out/ab_overlap0.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_emit\|k_jpeg_unit_bits -s 40 -c 4 \
  -o gpurun_out/r02b_jpeg python bench.py --batch 1 --steps 1 --warmup 0 --no-tiled-leg --no-cpu-baseline > gpurun_out/ncu_jpeg.log 2>&1
ncu -i gpurun_out/r02b_jpeg.ncu-rep --page raw --csv > gpurun_out/r02b_jpeg_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ["ab_overlap1", "ab_overlap0"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        tk = {k["name"]: (k["ms"], k["launches"]) for k in d["top_kernels"]}
        print(f, d["value"], d["single_image"]["ms"], d["single_image"]["breakdown_ms"], d["single_image_gpu_kernel_ms"],
              "emit", tk.get("jpeg_emit"), "bits", tk.get("jpeg_unit_bits"), "ff", tk.get("jpeg_count_ff"))
    except Exception as e:
        print(f, "failed", e)
PY
ls -la gpurun_out/*.ncu-rep 2>/dev/null
