#!/bin/bash
# Round-end measurement pass on one B200: parity tests, the bench lines of every workload,
# the reference arm, and the ncu evidence (launch list + one full capture per CTA kernel).
# Outputs land in gpurun_out/ (scratch); the summaries are copied into profiles/ by hand.
set -u
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 2> $O/bench_reference.err | tail -1 > $O/bench_reference.json
python bench.py 2> $O/bench_quadrotor_n1.err | tail -1 > $O/bench_quadrotor_n1.json
for w in cartpole humanoid newton_kkt newton_kkt_uniform; do
  python bench.py --workload $w 2> $O/bench_$w.err | tail -1 > $O/bench_$w.json
done
python bench.py --workload humanoid --input-layout problem_major --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_humanoid_problem_major.json
python bench.py --workload newton_kkt --force-generic 2> /dev/null | tail -1 > $O/bench_newton_kkt_generic.json
python bench.py --workload newton_kkt --pad-variable-dims 2> /dev/null | tail -1 > $O/bench_newton_kkt_padded.json
for w in long_horizon_quadrotor long_horizon_humanoid; do
  python bench.py --workload $w --steps 5 --no-e2e --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_$w.json
done
cut -c1-260 $O/bench_*.json
# ncu: launch list of the humanoid step, then full captures of its two kernels (2 waves).
H="python bench.py --workload humanoid --batch 592 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$H > $O/hum_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_humanoid_b592.csv $H > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_backward_cta -c 1 -f -o $O/hum_backward $H > $O/ncu_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_forward_cta -c 1 -f -o $O/hum_forward $H > $O/ncu_fwd.log 2>&1
ls -la $O | tail -20
# ncu: the uniform Newton-KKT step (launch list + the operator and reduction kernels).
K="python bench.py --workload newton_kkt_uniform --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$K > $O/kkt_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_newton_kkt_uniform.csv $K > $O/ncu_kkt_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kkt_apply_chain -c 1 -f -o $O/kkt_apply_chain $K > $O/ncu_kkt_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kkt_reduce_chain -c 1 -f -o $O/kkt_reduce_chain $K > $O/ncu_kkt_r.log 2>&1
ls -la $O | tail -8
