"""N > 1 host logic on CPU: contiguous batch shards and the per-iteration
statistics all-reduce, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle
from sip_optimal_control_b200.sharding import allreduce_stats, shard_range
import problem_gen as pg


def test_shards_partition_the_batch():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # Every rank builds the same global batch, solves only its shard with the
        # oracle, and contributes {sum sq, max, failed, count} of its slice.
        s, host = pg.lqr_benchmark_batch(3, 2, 5, total, seed=11)
        host["delta"][total - 1, 0] = -1.0  # one failing problem, owned by the last rank
        b, e = shard_range(total, rank, world)
        mine = {k: v[b:e] for k, v in host.items()}
        res = pyoracle.lqr_factor_solve(s, mine)
        ok = res["status"] == 0
        norms = res["residual"][ok]
        stats = torch.tensor([float((norms ** 2).sum()), float(norms.max(initial=0.0)),
                              float((~ok).sum()), float(e - b)], dtype=torch.float64)
        allreduce_stats(stats)
        np.save(os.path.join(out_dir, f"stats{rank}.npy"), stats.numpy())
    finally:
        dist.destroy_process_group()


def test_stats_allreduce_world2_gloo(tmp_path):
    total, world = 13, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"stats{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])
    s, host = pg.lqr_benchmark_batch(3, 2, 5, total, seed=11)
    host["delta"][total - 1, 0] = -1.0
    ref = pyoracle.lqr_factor_solve(s, host)
    ok = ref["status"] == 0
    assert got[0][2] == 1 and got[0][3] == total
    assert np.isclose(got[0][0], (ref["residual"][ok] ** 2).sum(), rtol=1e-12)
    assert np.isclose(got[0][1], ref["residual"][ok].max(), rtol=1e-12)
