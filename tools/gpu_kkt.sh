#!/bin/bash
# GPU-box script: Newton-KKT bench lines (uniform quadrotor-shaped and config 4), eager and graph.
# usage: tools/gpu_kkt.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
for wl in newton_kkt_uniform newton_kkt; do
  for g in "" "--graph"; do
    python bench.py --workload $wl $g --steps 20 --warmup 3 --no-cpu-baseline > $out/$wl$g.json 2> $out/$wl$g.err
    python - <<PY
import json
d=json.load(open("$out/$wl$g.json"))
print("$wl$g", round(d["value"]), round(d["ms_per_step"],3), d["config"].get("kernel_variant"), d["config"].get("r2_max"), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d.get("check"))
PY
  done
done
