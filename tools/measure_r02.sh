#!/bin/bash
# Round-2 measurement pass on one B200: parity tests, smoke, every bench line, ncu evidence.
# Outputs land in gpurun_out/r02final (scratch); summaries are copied into profiles/r02/.
set -u
O=gpurun_out/r02final
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 2> $O/bench_reference.err | tail -1 > $O/bench_reference.json
python bench.py 2> $O/bench_quadrotor_n1.err | tail -1 > $O/bench_quadrotor_n1.json
for w in cartpole humanoid newton_kkt newton_kkt_uniform; do
  python bench.py --workload $w 2> $O/bench_$w.err | tail -1 > $O/bench_$w.json
done
python bench.py --workload newton_kkt --graph --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_newton_kkt_graph.json
python bench.py --workload newton_kkt_uniform --graph --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_newton_kkt_uniform_graph.json
python bench.py --workload humanoid --input-layout problem_major --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_humanoid_problem_major.json
python bench.py --workload long_horizon_quadrotor --steps 10 --no-e2e --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_long_horizon_quadrotor_scan.json
python bench.py --workload long_horizon_quadrotor --serial-in-time --steps 5 --no-e2e --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_long_horizon_quadrotor_serial.json
python bench.py --workload long_horizon_humanoid --steps 3 --no-e2e --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_long_horizon_humanoid.json
python bench.py --batch 8192 --no-e2e --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_quadrotor_shard8192.json
python bench.py --workload cartpole --fp32 --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_cartpole_fp32.json
python bench.py --e2e-packed --e2e-steps 3 --no-cpu-baseline 2> /dev/null | tail -1 > $O/bench_quadrotor_n1_e2e_packed.json
cut -c1-230 $O/bench_*.json
# ncu: quadrotor step at batch 16 384 (two kernels) and at 8 192 (fused): launch lists + full captures
Q="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_quadrotor_b65536.csv $Q > $O/ncu_q_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_backward_subwarp -c 1 -f -o $O/quad_backward65536 $Q > $O/ncu_q_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_forward -c 1 -f -o $O/quad_forward65536 $Q > $O/ncu_q_f.log 2>&1
F="python bench.py --batch 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_quadrotor_b8192_fused.csv $F > $O/ncu_f_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:riccati_backward_subwarp -c 1 -f -o $O/quad_fused $F > $O/ncu_f.log 2>&1
S="python bench.py --workload long_horizon_quadrotor --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_long_horizon_scan.csv $S > $O/ncu_s_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_ -c 4 -f -o $O/scan_kernels $S > $O/ncu_s.log 2>&1
ls -la $O | tail -40
