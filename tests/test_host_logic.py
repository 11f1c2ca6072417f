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


def test_fetch_local_reads_both_output_formats(tmp_path):
    """tools/fetch_local.py (the counterpart of the reference's plot script) takes the grid shape from the file."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("fetch_local", os.path.join(ROOT, "tools", "fetch_local.py"))
    fl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fl)
    nx, ny = 6, 4
    rng = np.random.default_rng(0)
    p = rng.random((4, nx, ny)).astype(np.float32)
    with open(tmp_path / "results.dat", "w") as f:          # base_shll_2d.c:322-340
        for i in range(nx):
            for j in range(ny):
                f.write("%e\t%e\t%e\t%e\t%e\t%e\n" % ((i + 0.5) / nx, (j + 0.5) / ny, p[0, i, j], p[1, i, j], p[2, i, j], p[3, i, j]))
    r = fl.read_any(str(tmp_path / "results.dat"))
    assert (r["dims"], r["nx"], r["ny"]) == (2, nx, ny)
    assert np.allclose(r["fields"]["T"], p[3], rtol=1e-6) and np.allclose(r["fields"]["rho"], p[0], rtol=1e-6)
    hdr = np.zeros(16, dtype="<i4")                           # host/shll_main.c Save_Binary
    hdr[2:7] = [2, nx, ny, 4, 77]
    raw = bytearray(hdr.tobytes()); raw[:8] = b"SHLLBIN1"
    (tmp_path / "snapshot.bin").write_bytes(bytes(raw) + p.tobytes())
    r = fl.read_any(str(tmp_path / "snapshot.bin"))
    assert (r["dims"], r["nx"], r["ny"], r["steps"]) == (2, nx, ny, 77) and np.array_equal(r["fields"]["uy"], p[2])
    with open(tmp_path / "tube.dat", "w") as f:               # base_shll.c:180-192
        for i in range(9):
            f.write("%e\t%e\t%e\t%e\n" % ((i + 0.5) / 9, 1.0 + i, 0.5 * i, 2.0))
    r = fl.read_any(str(tmp_path / "tube.dat"))
    assert (r["dims"], r["nx"]) == (1, 9) and r["fields"]["u"][4] == 2.0
    assert "rho" in fl.summary(r) and fl.main([str(tmp_path / "tube.dat"), "--no-plot"]) == 0


# ------------------------------------------------------------------------------------------------------------------
# STRICT-mode exactness shortcuts (csrc/shll_math.cuh) in exact rational arithmetic.  The device twin -- the same
# sequences executed by the GPU against __fdiv_rn / __ddiv_rn over 2^28 operand pairs -- is
# tests/test_gpu_parity.py::test_exactness_shortcuts_device_selftest (csrc/selftest.cu).

def _rn(F, p, emin):
    """Round the rational F to the nearest binary floating-point number with p significand bits (ties to even), gradual
    underflow below 2^emin; returned as an exact Fraction."""
    from fractions import Fraction
    if F == 0:
        return Fraction(0)
    s = -1 if F < 0 else 1
    F = abs(F)
    e = F.numerator.bit_length() - F.denominator.bit_length()
    if Fraction(2) ** e > F:
        e -= 1
    e = max(e, emin)
    scale = Fraction(2) ** (e - (p - 1))
    m = F / scale
    fl = m.numerator // m.denominator
    r = m - fl
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return s * fl * scale


def _markstein_chunk(args):
    """div_by_cv (shll_math.cuh:100-114) on `count` doubles: q = RN(n*rc); rem = RN(n - CV*q); q' = RN(q + rem*rc) must equal
    RN(n / CV).  float(Fraction) is correctly rounded in CPython, so every RN53 here is exact."""
    import random
    import struct
    from fractions import Fraction
    seed, count = args
    CV = float(np.float32(1.0 / (float(np.float32(1.4)) - 1.0)))
    rc = float.fromhex("0x1.9999970a3d74cp-2")
    FCV, Frc = Fraction(CV), Fraction(rc)
    rng = random.Random(seed)
    bad, done = [], 0
    while done < count:
        kind = rng.random()
        if kind < 0.5:      # uniform bit patterns inside the guarded exponent range
            n = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        elif kind < 0.9:    # the operands the kernels produce: (double)e - 0.5*(double)k of float32 e, k of a gas state
            e, k = np.float32(rng.uniform(0.1, 50.0)), np.float32(rng.uniform(0.0, 20.0))
            n = float(e) - 0.5 * float(k)
        else:               # the guard edges |n| ~ 2^-969 and 2^976, and values whose quotient sits next to a rounding boundary
            ex = rng.choice([-968, -967, -900, -1, 0, 1, 900, 974, 975])
            n = (1.0 + rng.getrandbits(52) * 2.0 ** -52) * 2.0 ** ex * rng.choice([-1.0, 1.0])
        if n != n or abs(n) == float("inf") or abs(n) < 2.0 ** -968 or abs(n) > 2.0 ** 976:
            continue        # outside the guard: the kernel executes __ddiv_rn there
        q = n * rc
        rem = float(Fraction(n) - FCV * Fraction(q))
        q2 = float(Fraction(rem) * Frc + Fraction(q))
        if q2 != float(Fraction(n) / FCV):
            bad.append(n)
        done += 1
    return bad


def test_division_by_cv_markstein_sequence_is_correctly_rounded_in_exact_arithmetic():
    """>= 10^6 random + boundary doubles: the multiply + Markstein correction equals RN53(n / CV) in exact rational arithmetic."""
    import multiprocessing as mp
    from fractions import Fraction
    CV = float(np.float32(1.0 / (float(np.float32(1.4)) - 1.0)))
    assert CV == 2.5000002384185791015625 and float(1 / Fraction(CV)) == float.fromhex("0x1.9999970a3d74cp-2")   # rc = RN53(1/CV)
    nproc = max(1, min(8, (os.cpu_count() or 2) // 2))
    chunks = [(1000 + i, 1_000_000 // 40) for i in range(40)]
    with mp.get_context("fork").Pool(nproc) as pool:
        bad = [n for part in pool.map(_markstein_chunk, chunks) for n in part]
    assert not bad, f"{len(bad)} of 1e6 quotients are not correctly rounded, e.g. n = {bad[0].hex()}"


def test_shared_reciprocal_float_division_is_correctly_rounded_in_exact_arithmetic():
    """div_rn_shared / div_rn_spec (shll_math.cuh:76-98,124-140): with ANY starting reciprocal within 1 ulp of 1/b (MUFU.RCP's
    bound), one Newton step + quotient + remainder + correction gives RN24(a / b) whenever the guard holds
    (|b| in [2^-60, 2^60], |q0| in [2^-40, 2^40])."""
    import random
    from fractions import Fraction
    rng = random.Random(7)
    R24 = lambda F: _rn(F, 24, -126)

    def f32(ex):
        return Fraction(rng.choice([-1, 1]) * (2 ** 23 + rng.getrandbits(23))) * Fraction(2) ** (ex - 23)

    bad = 0
    for _ in range(60000):
        kind = rng.random()
        b = f32(rng.randint(-60, 59)) if kind < 0.3 else f32(rng.randint(-10, 10))
        a = f32(rng.randint(-80, 80)) if kind < 0.3 else f32(rng.randint(-20, 20))
        r0 = R24(1 / b)
        ulp = Fraction(2) ** ((abs(r0).numerator.bit_length() - abs(r0).denominator.bit_length()) - 23)
        r0 = r0 + rng.choice([-1, 0, 0, 1]) * ulp          # MUFU.RCP: within 1 ulp, not necessarily correctly rounded
        e = R24(1 - b * r0)
        r = R24(r0 + r0 * e)
        q0 = R24(a * r)
        if not (Fraction(2) ** -40 <= abs(q0) <= Fraction(2) ** 40):
            continue                                       # the kernel executes __fdiv_rn there
        rem = R24(a - b * q0)
        q = R24(q0 + r * rem)
        bad += (q != R24(a / b))
    assert bad == 0


def test_plan_halo_steps_is_one_value_for_the_whole_chain_of_slabs(monkeypatch):
    """shll_plan_halo_steps (include/shll_b200.h): steps per halo exchange round from the WHOLE domain, so that every slab of a
    balanced partition gets the same value -- 1D: K*order <= 32 cells and <= the smallest slab; 2D: 2 (two-step launches, two-row
    exchange) only where the 1st-order FAST face-flux kernel is selected and every slab has >= 16 rows.  Pure host logic: no GPU."""
    monkeypatch.delenv("SHLL_HALO_K", raising=False)
    plan = capi.plan_halo_steps
    F, S = capi.MODE_FAST, capi.MODE_STRICT
    assert plan(1, 1 << 26, 1, 2, F, 8) == 16 and plan(1, 1 << 26, 1, 1, F, 8) == 16      # order 2: 32 cells; order 1: 16 cells
    assert plan(1, 1 << 26, 1, 2, F, 1) == 1                                               # one slab: nothing to exchange
    assert plan(1, 40, 1, 2, F, 8) == 2 and plan(1, 15, 1, 2, S, 3) == 2 and plan(1, 9, 1, 1, S, 3) == 3   # limited by the smallest slab
    assert plan(1, 6, 1, 2, S, 3) == 1
    monkeypatch.setenv("SHLL_HALO_K", "5")
    assert plan(1, 1 << 20, 1, 2, F, 4) == 5
    monkeypatch.delenv("SHLL_HALO_K")
    assert plan(2, 8 * 4096, 4096, 1, F, 8) == 2                   # configs[2] weak-scaled: two-step launches
    assert plan(2, 8 * 4096, 4096, 1, S, 8) == 1                   # STRICT: one step per launch
    assert plan(2, 16384, 16384, 2, F, 8) == 1                     # 2nd order: one step per launch
    assert plan(2, 8 * 4096, 4100, 1, F, 8) == 1                   # ny % 8 != 0: not the face-flux kernel
    assert plan(2, 8 * 15, 256, 1, F, 8) == 1 and plan(2, 8 * 16, 256, 1, F, 8) == 2      # slabs thinner than 16 rows keep one step
    monkeypatch.setenv("SHLL_FUSE2", "0")
    assert plan(2, 8 * 4096, 4096, 1, F, 8) == 1
