#!/bin/bash
# ncu evidence for the config-4 (Newton-KKT, per-stage dims) and cartpole steps: launch lists and one
# full capture of config 4's dominant kernel.  Outputs in gpurun_out/kkt/.
set -u
O=gpurun_out/kkt
mkdir -p $O
K="python bench.py --workload newton_kkt --steps 2 --warmup 1"
$K > $O/kkt_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_newton_kkt.csv $K > $O/ncu_launch_kkt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:strict_factor_thread -c 1 -f -o $O/strict_factor $K > $O/ncu_strict.log 2>&1
C="python bench.py --workload cartpole --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$C > $O/cartpole_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cartpole.csv $C > $O/ncu_launch_cartpole.log 2>&1
ls -la $O
