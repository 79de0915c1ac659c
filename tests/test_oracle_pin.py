"""CPU tests (no GPU): the oracle (oracle/mpm_oracle.c) is pinned against the UNMODIFIED reference.

* against committed golden vectors produced by the reference itself (tests/golden/make_golden.py);
* directly against the reference's host build (oracle/_ref/libmaniskill_mpm_cpu.so) on fresh seeds, kernel by kernel.
"""
import numpy as np
import pytest

from abi1_driver import Abi1Sim, loss_seed
from conftest import rel_err
from dexdeform_b200.scenes import make_scene
from parity_util import check_against_golden, golden_cases


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(oracle_lib, name):
    check_against_golden(oracle_lib, name)


def test_reference_host_build_reproduces_golden(ref_cpu):
    # the fixtures themselves are reproducible (up to float-atomic ordering) from the reference build
    for name in golden_cases():
        check_against_golden(ref_cpu, name)


@pytest.mark.parametrize("seed", [21, 22])
def test_oracle_vs_reference_kernel_by_kernel(oracle_lib, ref_cpu, seed):
    sc = make_scene(700, 32, box_width=(0.1, 0.1, 0.1), steps=1, perturb=0.03, vel_scale=0.5, on_floor=True, seed=seed, nb=7)
    ref, orc = Abi1Sim(ref_cpu, sc, 1), Abi1Sim(oracle_lib, sc, 1)
    for s in (ref, orc):
        s.substep(0)
    names = ("U", "V", "sig", "F", "grid_m", "grid_v_in", "grid_v_out", "grid_body_v_in")
    a, b = ref.get_temp(*names), orc.get_temp(*names)
    for k in ("U", "V", "sig", "F"):
        assert np.array_equal(a[k], b[k]), k  # same algorithm in double -> identical floats
    for k in ("grid_m", "grid_v_in", "grid_v_out", "grid_body_v_in"):
        assert rel_err(b[k], a[k]) < 1e-5, k
    seedg = loss_seed(sc["n"], seed)
    for s in (ref, orc):
        for k, v in seedg.items():
            s.states[1][k].upload(v)
        s.substep_grad(0)
    gn = ("grid_v_out_grad", "grid_v_in_grad", "grid_m_grad", "U_grad", "V_grad", "sig_grad", "F_grad")
    a, b = ref.get_temp(*gn), orc.get_temp(*gn)
    for k in gn:
        assert rel_err(b[k], a[k]) < (2e-2 if k == "F_grad" else 1e-4), (k, rel_err(b[k], a[k]))


def test_oracle_svd_properties(oracle_lib):
    import ctypes
    rng = np.random.default_rng(0)
    n = 500
    A = (np.eye(3)[None] + 0.3 * rng.normal(size=(n, 3, 3))).astype(np.float32)
    A[:5] = np.eye(3)  # repeated singular values
    U, s, V = np.zeros_like(A), np.zeros((n, 3), np.float32), np.zeros_like(A)
    fn = oracle_lib.raw.orc_svd3
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    fn(A.ctypes.data, U.ctypes.data, s.ctypes.data, V.ctypes.data, n)
    rec = np.einsum("nij,nj,nkj->nik", U, s, V)
    assert np.abs(rec - A).max() < 5e-6
    assert np.abs(np.einsum("nij,nkj->nik", U, U) - np.eye(3)).max() < 1e-5
    assert np.abs(np.einsum("nij,nkj->nik", V, V) - np.eye(3)).max() < 1e-5
    assert np.allclose(np.abs(s), np.linalg.svd(A.astype(np.float64), compute_uv=False), atol=5e-6)
    assert (np.abs(s[:, 0]) >= np.abs(s[:, 1]) - 1e-6).all() and (np.abs(s[:, 1]) >= np.abs(s[:, 2]) - 1e-6).all()
    assert np.linalg.det(U.astype(np.float64)).min() > 0.99 and np.linalg.det(V.astype(np.float64)).min() > 0.99


def test_oracle_empty_and_single_particle(oracle_lib):
    sc = make_scene(1, 32, steps=1, nb=2, seed=1, on_floor=True)
    sim = Abi1Sim(oracle_lib, sc, 1)
    sim.substep(0)
    out = sim.get(1)
    assert np.isfinite(out["x"]).all() and np.isfinite(out["C"]).all()
    m = sim.get_temp("grid_m")["grid_m"]
    assert np.isclose(m.sum(), sc["mass"][0], rtol=1e-5) and (m > 0).sum() == 27
    sim.n = 0  # dim = 0: every kernel is a no-op
    sim.substep(0)


def test_host_and_gpu_reference_fixtures_agree():
    """The two fixture sets (reference built with g++ vs nvcc) describe the same computation: they agree to within the
    FMA-contraction sensitivity of the path (worst at rest, where stress is pure rounding noise)."""
    import json
    from parity_util import load_golden
    for name in golden_cases():
        _, a = load_golden(name)
        _, b = load_golden(name, "_gpu")
        sp = json.loads(str(b["ref_spread"]))
        assert np.array_equal(a["grid_idx"], b["grid_idx"])
        for k in ("s4_x", "s4_v", "s4_F", "s4_C", "grid_m", "grid_v_out", "g0_x_grad", "pos_grad"):
            if k in a.files:
                assert rel_err(b[k], a[k]) < max(2e-3, 4 * sp.get(k, 0)), (name, k)
