"""CPU: host-side logic (shll_sve_cfd_b200.programs) against the oracle, and the C-ABI library's surface.
No compute call is made on the library here (there is no GPU in the build container)."""
import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ORACLE_IC, ROOT, bits, has_gpu, load_golden, oracle_cfg_for, problem_from_manifest
from shll_sve_cfd_b200 import capi, programs

PBS = [
    programs.BASE_SHLL, programs.BASE_SHLL.resized(1000), programs.BASE_SHLL_2D.resized(64, 80),
    programs.SECOND_ORDER_2D.resized(64, 48), programs.SECOND_ORDER_1D.resized(512), programs.BASE_OMP_2D.resized(32, 36),
]


@pytest.mark.parametrize("pb", PBS, ids=[f"{p.name}_{p.nx}x{p.ny}" for p in PBS])
def test_ic_and_cons_from_prim_match_oracle(pb, oracle):
    O = oracle
    cfg = oracle_cfg_for(O, pb)
    p_o = O.init_prim(cfg, ORACLE_IC[pb.ic])
    p_h = programs.initial_primitives(pb)
    assert np.array_equal(bits(p_h), bits(p_o))
    u_o = O.cons_from_prim(cfg, p_o)
    u_h = programs.cons_from_prim(pb, p_h)
    assert np.array_equal(bits(u_h), bits(u_o))
    # advance a little so the state is not piecewise constant, then compare Compute_P_from_U
    u = O.run(cfg, u_o, 7)
    p_o2, a_o = O.prim_from_cons(cfg, u, want_a=True)
    p_h2, a_h = programs.prim_from_cons(pb, u)
    assert np.array_equal(bits(p_h2), bits(p_o2))
    assert np.array_equal(bits(a_h), bits(a_o))


def test_slab_initial_condition_is_a_slice_of_the_global_one():
    pb = programs.SECOND_ORDER_2D.resized(64, 48)
    full = programs.initial_primitives(pb).reshape(4, 64, 48)
    for i0, nl in ((0, 16), (16, 16), (48, 16), (5, 23)):
        part = programs.initial_primitives(pb, i0=i0, nx_local=nl, nx_global=64).reshape(4, nl, 48)
        assert np.array_equal(bits(part), bits(full[:, i0:i0 + nl]))


def test_time_constants_and_step_counts(oracle):
    for n, steps in ((256, 410), (1024, 1639), (65536, 104858)):
        pb = programs.BASE_SHLL.resized(n)
        dx, dy, dt, dtdx, dtdy = programs.time_constants(pb)
        assert dtdx == np.float32(0.125) and dt == np.float32(0.125) / np.float32(n)
        assert programs.count_steps(pb) == steps == oracle.count_steps(dt, pb.total_time)
    assert programs.count_steps(programs.BASE_SHLL_2D) == 205
    assert programs.count_steps(programs.SECOND_ORDER_2D) == 1639
    pb = programs.BASE_SHLL_2D.resized(96, 160)
    _, _, _, dtdx, dtdy = programs.time_constants(pb)
    assert dtdx == np.float32(0.125) and dtdy != np.float32(0.125)
    with pytest.raises(ValueError):
        programs.count_steps(programs.BASE_SHLL.resized(2 ** 26))


@pytest.mark.parametrize("case", ["1d_o1_256", "2d_o1_64", "2d_o2_64"])
def test_save_results_text_matches_reference_md5(case, manifest, tmp_path):
    pb = problem_from_manifest(manifest[case])
    _, gp, _ = load_golden(case)
    out = os.path.join(tmp_path, "results.dat")
    programs.save_results(pb, gp, out)
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == manifest[case]["results_dat_md5"]


# ------------------------------------------------------------------------------------------------ C-ABI surface

def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "shll_b200.h")).read()
    return sorted(set(re.findall(r"SHLL_API\s+[\w\s\*]+?\b(shll_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(capi.API_SYMBOLS)


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "libshll_b200.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("shll_"))
    assert exported == _declared_symbols()
    assert capi.lib().shll_abi_version() == 1
    assert ctypes.sizeof(capi.Config) == 92 and ctypes.sizeof(capi.PeerDesc) == 184


def test_count_steps_through_the_c_abi():
    assert capi.count_steps(np.float32(0.125) / np.float32(1024), 0.2) == 1639
    assert capi.count_steps(np.float32(0.125) / np.float32(100000), 0.2) == 159820
    with pytest.raises(capi.ShllError) as ei:
        capi.count_steps(np.float32(0.125) / np.float32(2 ** 26), 0.2)
    assert ei.value.code == capi.E_INVAL and "stalls" in str(ei.value)


def test_bad_configs_are_rejected_before_touching_cuda():
    for kw in (dict(dims=3, nx=8), dict(dims=1, nx=1), dict(dims=2, nx=8, ny=1), dict(dims=1, nx=8, order=3),
               dict(dims=1, nx=8, bc=7), dict(dims=1, nx=8, dt_on_dx=0.0), dict(dims=1, nx=8, rank=2, nranks=2)):
        with pytest.raises(capi.ShllError) as ei:
            capi.Solver(**kw)
        assert ei.value.code == capi.E_INVAL


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    """The product path must fail loudly when there is no CUDA device -- never route through the oracle."""
    with pytest.raises(capi.ShllError) as ei:
        capi.Solver(1, 256)
    assert ei.value.code == capi.E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shll_sve_cfd_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "shll_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn
    for fn in os.listdir(os.path.join(ROOT, "host")):
        if fn.endswith((".c", ".h")) or fn == "Makefile":
            assert "oracle" not in open(os.path.join(ROOT, "host", fn)).read(), fn
