"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/sipoc.h declares, validates structures like the reference does, and
refuses to run without a GPU (no CPU fallback)."""
import ctypes

import numpy as np
import pytest

import sip_optimal_control_b200 as pkg
from sip_optimal_control_b200 import _capi
from sip_optimal_control_b200.lqr import Dimensions, Engine, Topology


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = pkg.declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert _capi.lib.sipoc_version() == 200


def test_header_cites_the_reference_interfaces():
    text = open(_capi.HEADER_PATH).read()
    for needle in ("lqr.hpp:192-193", "lqr.cpp:735-871", "helpers.cpp:190-407",
                   "helpers.cpp:953-977", "types.cpp:24-64"):
        assert needle in text


def _create_status(topology, dims, batch=4):
    e = Engine(dims, topology, batch)
    st = e.create_status
    e.close()
    return st


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_rejects_invalid_structures_before_touching_cuda():
    # lqr_test.cpp:452 (two edges into one child), :955 (disconnected), :969 (cycle)
    dims = Dimensions(0, [2, 2, 2], [1, 1])
    assert _create_status(Topology(2, 0, [0, 0], [1, 1]), dims) == _capi.SIPOC_INVALID_TOPOLOGY
    d5 = Dimensions(0, [3, 1, 2, 4, 2], [2, 1, 3, 1])
    assert _create_status(Topology(4, 0, [0, 0, 1, 4], [1, 2, 3, 3]), d5) == \
        _capi.SIPOC_INVALID_TOPOLOGY
    assert _create_status(Topology(4, 0, [4, 0, 1, 1], [1, 2, 3, 4]), d5) == \
        _capi.SIPOC_INVALID_TOPOLOGY
    # variable_dimensions_test.cpp:208-223: DAG rejected, negative dim rejected
    dk = Dimensions(0, [2, 1, 3], [1, 2], [0, 1, 0], [1, 0, 2], [2, 1], [1, 3])
    assert _create_status(Topology(2, 0, [0, 1], [2, 2]), dk) == _capi.SIPOC_INVALID_TOPOLOGY
    dneg = Dimensions(0, [2, 1, 3], [1, 2], [0, 1, 0], [1, 0, 2], [-1, 1], [1, 3])
    assert _create_status(Topology(2, 0, [0, 0], [1, 2]), dneg) == \
        _capi.SIPOC_INVALID_DIMENSIONS
    # theta (Schur) variables beyond the engine's limit are refused, not ignored
    dth = Dimensions(33, [2, 1, 3], [1, 2])
    assert _create_status(Topology(2, 0, [0, 0], [1, 2]), dth) == 6  # SIPOC_UNSUPPORTED


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_gpu():
    st = _create_status(Topology.chain(3), Dimensions.uniform(3, 2, 1))
    assert st == 2  # SIPOC_CUDA_ERROR: the engine refuses to run on the CPU
    with pytest.raises(pkg.SipocError):
        pkg.LQR(Dimensions.uniform(3, 2, 1), Topology.chain(3), 4)


def test_dimension_queries_match_reference_formulas():
    # lqr.cpp:146-180 on variable_dimensions_test.cpp:266-271
    d = Dimensions(0, [2, 1, 3], [1, 2], [1, 0, 2], [0, 2, 1], [1, 2], [2, 1])
    assert d.get_stagewise_x_dim(2) == 9
    assert d.get_y_dim(2) == 12
    assert d.get_z_dim(2) == 6
    assert d.get_stagewise_kkt_dim(2) == 27
    u = Dimensions.uniform(5, 4, 2, edge_c_dim=3)
    assert u.get_x_dim(5) == 5 * 6 + 4 and u.get_y_dim(5) == 6 * 4 + 15
    t = Topology(3)
    t.set_chain()
    assert t.edge_parents.tolist() == [0, 1, 2] and t.edge_children.tolist() == [1, 2, 3]
    assert t.num_nodes() == 4


def test_graph_entry_points_reject_null_handles():
    # argument checks (SIPOC_INVALID_ARGUMENT = 1) come before any CUDA call
    lib = _capi.lib
    out = ctypes.c_void_p()
    assert lib.sipoc_graph_begin(None, None) == 1
    assert lib.sipoc_graph_end(None, None, ctypes.byref(out)) == 1
    assert lib.sipoc_graph_launch(None, None, None) == 1
    assert lib.sipoc_graph_kernel_count(None) == 0
    lib.sipoc_graph_destroy(None)


def _structure(topology, dims, batch=1):
    keep = [np.ascontiguousarray(a, dtype=np.int32) for a in
            (topology.edge_parents, topology.edge_children, dims.state_dims, dims.control_dims)]
    ip = lambda a: a.ctypes.data_as(_capi.c_int_p)
    s = _capi.Structure()
    s.num_edges, s.root = topology.num_edges, topology.root
    s.edge_parents, s.edge_children, s.state_dims, s.control_dims = (ip(a) for a in keep)
    s.theta_dim, s.batch, s.device, s.flags = dims.theta_dim, batch, -1, 0
    return s, keep


def test_validate_runs_without_a_device():
    # sipoc_validate == validate_input (types.cpp:68-134): what the C++ shim calls
    s, keep = _structure(Topology(2, 0, [0, 0], [1, 2]), Dimensions(2, [2, 1, 3], [1, 2]))
    assert _capi.lib.sipoc_validate(ctypes.byref(s)) == _capi.SIPOC_OK
    s, keep = _structure(Topology(2, 0, [0, 1], [2, 2]), Dimensions(0, [2, 1, 3], [1, 2]))
    assert _capi.lib.sipoc_validate(ctypes.byref(s)) == _capi.SIPOC_INVALID_TOPOLOGY
    s, keep = _structure(Topology(2, 0, [0, 0], [1, 2]), Dimensions(0, [2, -1, 3], [1, 2]))
    assert _capi.lib.sipoc_validate(ctypes.byref(s)) == _capi.SIPOC_INVALID_DIMENSIONS


def test_shard_range_of_the_c_abi_matches_the_python_one():
    from sip_optimal_control_b200.sharding import shard_range

    b, e = ctypes.c_int64(), ctypes.c_int64()
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                assert _capi.lib.sipoc_shard_range(total, rank, world, ctypes.byref(b),
                                                   ctypes.byref(e)) == 0
                assert (b.value, e.value) == shard_range(total, rank, world)
    assert _capi.lib.sipoc_shard_range(8, 2, 2, ctypes.byref(b), ctypes.byref(e)) == 1
    # communicator entry points check their arguments before touching NCCL / CUDA
    assert _capi.lib.sipoc_comm_allreduce_stats(None, None, None) == 1
    assert _capi.lib.sipoc_attach_comm(None, None) == 1
    assert _capi.lib.sipoc_comm_size(None) == 0
