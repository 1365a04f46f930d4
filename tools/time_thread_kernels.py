import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from sip_optimal_control_b200 import LQR, Dimensions, Topology
n, m, T, batch = 4, 1, 100, 16384
lqr = LQR(Dimensions.uniform(T, n, m), Topology.chain(T), batch, device=0)
inp = lqr.generate_benchmark(seed=1, problem_offset=0)
out = lqr.alloc_output()
st = lqr.engine.empty_int()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("default fp64 factor_solve ms", timeit(lambda: lqr.factor_solve(inp, out, status=st)))
print("thread kernels on double  ms", timeit(lambda: lqr.factor_solve_thread_f64(inp, out, status=st)))
inp32, out32 = lqr.narrow_f32(inp), lqr.alloc_output_f32()
print("fp32 mode                 ms", timeit(lambda: lqr.factor_solve_f32(inp32, out32, status=st)))
