#!/bin/bash
# GPU-box script: scan tests, long-horizon bench at a few segment lengths, optional ncu capture.
# usage: tools/gpu_scan.sh <outdir> "<L list>" [ncu]
out=gpurun_out/$1; mkdir -p $out
python -m pytest tests/test_gpu_scan.py -m gpu -q -x 2>&1 | tail -3
for L in $2; do
  SIPOC_SCAN_SEGMENT=$L python bench.py --workload long_horizon_quadrotor --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/scan_L$L.json 2> $out/scan_L$L.err
  python - <<PY
import json
d=json.load(open("$out/scan_L$L.json"))
print($L, round(d["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d.get("check"))
PY
done
if [ "$3" = "ncu" ]; then
  SIPOC_SCAN_SEGMENT=${4:-32} ncu --set full --import-source on --clock-control none -k regex:scan_ -c 2 -o $out/scan_seg -f python bench.py --workload long_horizon_quadrotor --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $out/ncu_scan.log 2>&1
  tail -3 $out/ncu_scan.log
fi
