"""The multi-GPU side of the C ABI: per-device engines over shards of one batch and the
all-reduce of the four statistics through the engine's own NCCL communicator
(include/sipoc.h, "several devices").  Runs with one rank on a single GPU; the two-device
cases need a second GPU (gpurun --gpus 2)."""
import ctypes

import numpy as np
import pytest

import problem_gen as pg
from gpu_helpers import to_structs
from oracle import pyoracle
from sip_optimal_control_b200 import LQR, _capi
from sip_optimal_control_b200._capi import lib
from sip_optimal_control_b200.sharding import Communicator, shard_range

pytestmark = pytest.mark.gpu


def _failing_batch(total):
    s, host = pg.lqr_benchmark_batch(4, 1, 6, total, seed=5)
    host["delta"][1, 0] = -1.0          # INVALID_DELTA in the first shard
    host["delta"][total - 2, 3] = 0.0   # ... and in the last
    return s, host


def test_single_rank_communicator_and_attached_engine():
    import torch

    s, host = _failing_batch(48)
    dims, topo = to_structs(s)
    lqr = LQR(dims, topo, 48)
    comm = Communicator(device=0)
    assert (comm.rank, comm.world) == (0, 1)
    inp, out = lqr.pack_input(host), lqr.alloc_output()
    status = lqr.factor_solve(inp, out)
    _, plain = lqr.residual(inp, out, status)
    comm.attach(lqr.engine)
    _, reduced = lqr.residual(inp, out, status)   # one rank: the all-reduce is the identity
    torch.cuda.synchronize()
    assert np.array_equal(plain.cpu().numpy(), reduced.cpu().numpy())
    assert reduced[2].item() == 2 and reduced[3].item() == 48
    comm.close()


def test_two_devices_one_process_shards_and_allreduce():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    total = 101
    s, host = _failing_batch(total)
    dims, topo = to_structs(s)
    ref = pyoracle.lqr_factor_solve(s, host)
    ok = ref["status"] == 0
    handles = (ctypes.c_void_p * 2)()
    devs = (ctypes.c_int * 2)(0, 1)
    assert lib.sipoc_comm_create_all(devs, 2, handles) == 0
    stats, keep = [], []
    for rank in range(2):
        b, e = shard_range(total, rank, 2)
        lqr = LQR(dims, topo, e - b, device=rank)
        mine = {k: v[b:e] for k, v in host.items()}
        with torch.cuda.device(rank):
            inp, out = lqr.pack_input(mine), lqr.alloc_output()
            status = lqr.factor_solve(inp, out)
            _, st = lqr.residual(inp, out, status)
            torch.cuda.synchronize()
        stats.append(st)
        keep.append((lqr, inp, out))
    # one thread drives both communicators: the gathers go inside one NCCL group
    assert lib.sipoc_comm_group_begin() == 0
    for rank in range(2):
        with torch.cuda.device(rank):
            sp = int(torch.cuda.current_stream(rank).cuda_stream)
            assert lib.sipoc_comm_allgather_stats(handles[rank], stats[rank].data_ptr(), sp) == 0
    assert lib.sipoc_comm_group_end() == 0
    for rank in range(2):
        with torch.cuda.device(rank):
            sp = int(torch.cuda.current_stream(rank).cuda_stream)
            assert lib.sipoc_comm_fold_stats(handles[rank], stats[rank].data_ptr(), sp) == 0
            torch.cuda.synchronize()
    got = [st.cpu().numpy() for st in stats]
    assert np.array_equal(got[0], got[1])
    assert got[0][2] == 2 and got[0][3] == total
    assert np.isclose(got[0][0], (ref["residual"][ok] ** 2).sum(), rtol=1e-6)
    assert np.isclose(got[0][1], ref["residual"][ok].max(), rtol=1e-6)
    for rank in range(2):
        lib.sipoc_comm_destroy(handles[rank])
