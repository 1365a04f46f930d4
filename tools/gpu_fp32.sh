#!/bin/bash
# GPU-box script: FP32-mode tests and the cartpole bench lines (FP64 default, FP32 mode).
out=gpurun_out/${1:-fp32}; mkdir -p $out
python -m pytest tests/test_gpu_fp32.py -m gpu -q -x -s 2>&1 | tail -22
for f in "" "--fp32"; do
  python bench.py --workload cartpole $f --no-cpu-baseline > $out/cartpole$f.json 2> $out/cartpole$f.err
  tail -2 $out/cartpole$f.err
  python - <<PY
import json
d=json.load(open("$out/cartpole$f.json"))
print(d["dtype"], round(d["value"]), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d["check"])
PY
done
