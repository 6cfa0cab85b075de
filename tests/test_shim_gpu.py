"""A solver written against SPADE's own C++ API (integration/tgv_shim_demo.cc, compiled in the dev container against the
unmodified reference headers + include/spade_b200_shim.hpp) runs the reference's generic CUDA path and the drop-in
back to back on the GPU; the final states must agree to the north-star tolerance."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "tgv_shim_demo")


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [0, 1])
def test_spade_api_solver_through_the_shim(scheme):
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/tgv_shim_demo not built (needs /root/reference at build time)")
    out = subprocess.run([BIN, "2", "16", "2", str(scheme)], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(BIN))
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["rel_l2"] < 1e-12                  # callbacks as lambdas: flux_div(b200) + spb_rk_update + exchange
    assert line["rel_l2_fused"] < 1e-12            # callbacks as b200::flux_div_rhs / b200::exchange_bc: one kernel per stage
    assert line["umax"] > 300.0


BIN_CURV = os.path.join(ROOT, "integration", "_build", "channel_curv_demo")


@pytest.mark.gpu
def test_spade_api_solver_on_a_stretched_grid_through_the_shim():
    """coords::diagonal_coords(scaled, integrated_tanh_1D, identity): the reference's own CUDA flux_div(basic) for totani_lr
    (compiled with the header repair of oracle/ref_driver_curv.cc) against the drop-in reading the same grid object; and the
    config-3 functor (hybrid + ducros + visc_lr) on that grid through both drop-in paths."""
    if not os.path.exists(BIN_CURV):
        pytest.skip("integration/_build/channel_curv_demo not built (needs /root/reference at build time)")
    out = subprocess.run([BIN_CURV, "2", "16", "2"], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(BIN_CURV))
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["rel_l2_convective"] < 1e-12
    assert line["rel_l2_hybrid_fused_vs_unfused"] < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [0, 1])
def test_spade_api_solver_on_two_gpus_in_one_process(scheme):
    """The reference's own multi-GPU model: one host thread per GPU inside one process (compute_env_t::exec,
    compute_pool.h:497-514). The drop-in's exchange packs each message straight into the peer GPU's receive buffer."""
    import torch
    if not os.path.exists(BIN):
        pytest.skip("integration/_build/tgv_shim_demo not built (needs /root/reference at build time)")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([BIN, "2", "16", "2", str(scheme), "2"], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(BIN))
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["gpus"] == 2
    assert line["rel_l2"] < 1e-12
    assert line["rel_l2_fused"] < 1e-12


BIN_BENCH = os.path.join(ROOT, "integration", "_build", "bench_shim")
BIN_REFGPU = os.path.join(ROOT, "integration", "_build", "ref_gpu_bench")


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [1, 2])
def test_bench_shim_runs_the_overlapped_schedule(gpus):
    """integration/bench_shim.cc: the weak-scaling benchmark through the C++ shim (named callbacks -> block parts on two streams,
    peer buffers with stream-ordered flags); here only that it runs and stays finite on a small lattice"""
    import torch
    if not os.path.exists(BIN_BENCH):
        pytest.skip("integration/_build/bench_shim not built (needs /root/reference at build time)")
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    out = subprocess.run([BIN_BENCH, str(gpus), "3", "2", "2", "16"], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(BIN_BENCH))
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == gpus and line["finite"] and line["value"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_drop_in_against_every_valid_reference_gpu_variant(scheme):
    """integration/ref_gpu_bench.cc: the reference's own CUDA path with every algorithm tag that applies (basic, longf, fldbc,
    fused) next to the drop-in on the same solver; the drop-in must agree with tag `basic` to 1e-12 and a best valid tag exists"""
    if not os.path.exists(BIN_REFGPU):
        pytest.skip("integration/_build/ref_gpu_bench not built (needs /root/reference at build time)")
    out = subprocess.run([BIN_REFGPU, "2", "16", "1", str(scheme)], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(BIN_REFGPU))
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["variants"]["basic"]["valid"] and line["tag"] != "none"
    assert 0.0 <= line["b200_rel_l2_vs_basic"] < 1e-12
