"""The C++ host side (sip_optimal_control_b200/host): the reference's class names over
the C ABI.  CPU: the library is built and exports the mirrored interface.  GPU: the
restated reference unit tests (host_tests.cpp) pass against the engine."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "sip_optimal_control_b200", "lib", "libsipoc_host.so")
HOST_TESTS = os.path.join(ROOT, "build", "host_tests")


def _ensure_built():
    if not (os.path.exists(HOST_LIB) and os.path.exists(HOST_TESTS)):
        subprocess.check_call(["make", "-C", ROOT, "-j", "4", "all"])


def test_host_library_exports_the_reference_interface():
    _ensure_built()
    out = subprocess.run(["nm", "-D", "-C", "--defined-only", HOST_LIB], check=True,
                         capture_output=True, text=True).stdout
    for symbol in (
        "sip::optimal_control::Topology::set_chain()",
        "sip::optimal_control::Topology::set_tree(int, int const*, int const*)",
        "sip::optimal_control::Dimensions::set_uniform(int, int, int, int, int, int, int, int)",
        "sip::optimal_control::Dimensions::get_stagewise_kkt_dim(int) const",
        "sip::optimal_control::LQR::LQR(",
        "sip::optimal_control::LQR::compile_topology()",
        "sip::optimal_control::LQR::factor_with_status()",
        "sip::optimal_control::LQR::factor()",
        "sip::optimal_control::LQR::solve(sip::optimal_control::LQR::Output&)",
        "sip::optimal_control::CallbackProvider::factor(double const*, double const*, "
        "double const*, double const*)",
        "sip::optimal_control::CallbackProvider::solve(double const*, double*)",
        "sip::optimal_control::CallbackProvider::add_Kx_to_y(",
        "sip::optimal_control::validate_input(",
        "sip::optimal_control::BatchedLQR::factor_solve(",
    ):
        assert symbol in out, symbol


@pytest.mark.gpu
def test_restated_reference_unit_tests_pass_in_cpp():
    _ensure_built()
    run = subprocess.run([HOST_TESTS], capture_output=True, text=True, timeout=300)
    print(run.stdout[-3000:], run.stderr[-2000:])
    assert run.returncode == 0, run.stdout[-3000:]
    assert "0 check(s) failed" in run.stdout
