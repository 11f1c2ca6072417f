"""CPU: the oracle restatement (oracle/shll_oracle.c) reproduces the reference's own output bit for bit.

The golden fixtures were produced by the reference's unmodified arithmetic compiled from /root/reference
(tests/golden/make_golden.py); the md5s of results.dat for the shipped configurations are the ones recorded in
SURVEY.md App. B.  This is what "parity pinned" rests on.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import FAST_TOL, FAST_TOL_DEFAULT, ORACLE_IC, bits, load_golden, oracle_cfg_for, problem_from_manifest

CASES = ["1d_o1_256", "1d_o1_1024", "2d_o1_64", "2d_o1_96x160", "2d_o1_256", "2d_o2_64", "2d_o2_96x160",
         "1d_o2_slice_1024", "omp_o2_64"]

# md5(results.dat) of the unmodified reference, as recorded in SURVEY.md App. B
SURVEY_MD5 = {
    "1d_o1_256": "746eebed4f5bbd855b4f19f3a2d4aad3",
    "1d_o1_1024": "d193b8261925301b45682f9eccb520f3",
    "2d_o1_256": "97e02f6194cd77a5ad1b2925e1df45fa",
    "2d_o2_64": "caaec7d8c05b006f1353b4d37586f775",
    "1d_o2_slice_1024": "fab74c0de3b2c6bcd0e1a5e3f4d5fdb2",
}


def test_manifest_matches_survey_md5(manifest):
    for case, md5 in SURVEY_MD5.items():
        assert manifest[case]["results_dat_md5"] == md5


@pytest.mark.parametrize("case", CASES)
def test_oracle_bit_exact_vs_reference(case, manifest, oracle, tmp_path):
    O = oracle
    entry = manifest[case]
    pb = problem_from_manifest(entry)
    gu, gp, gsteps = load_golden(case)
    if case.startswith("1d_o2_slice"):
        pb = pb.resized(entry["nx"])
    cfg = oracle_cfg_for(O, pb, nthreads=2)
    dx = np.float32(1.0) / np.float32(pb.nx)
    steps = O.count_steps(np.float32(0.125) * dx, pb.total_time)
    assert steps == gsteps == entry["steps"]
    p0 = O.init_prim(cfg, ORACLE_IC[pb.ic])
    u = O.run(cfg, O.cons_from_prim(cfg, p0), steps)
    assert np.array_equal(bits(u), bits(gu)), f"max |du| = {np.abs(u - gu).max()}"
    p = O.prim_from_cons(cfg, u)
    assert np.array_equal(bits(p), bits(gp))
    # results.dat text: only comparable when the fixture is the program's own dump (not the 1D column of the slice run)
    if not case.startswith("1d_o2_slice"):
        out = os.path.join(tmp_path, "results.dat")
        O.save_results(cfg, p, out)
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == entry["results_dat_md5"]


def test_step_counts_match_readme(oracle):
    """README.md:77-85 (1D), :163-166 (2D 1st), :229-230 (2D 2nd) -- the only machine-checkable numbers the reference publishes."""
    O = oracle
    readme_1d = {256: 410, 512: 820, 1024: 1639, 2048: 3277, 4096: 6554, 8192: 13108, 16384: 26215, 32768: 52429, 65536: 104858}
    for n, s in readme_1d.items():
        assert O.count_steps(np.float32(0.125) / np.float32(n), 0.2) == s
    for n, s in {256: 205, 512: 410, 1024: 820, 2048: 1639}.items():
        assert O.count_steps(np.float32(0.125) / np.float32(n), 0.1) == s
    for n, s in {256: 1639, 512: 3277, 1024: 6554}.items():  # README prints 1649 for 256^2: a typo
        assert O.count_steps(np.float32(0.125) / np.float32(n), 0.8) == s
    # float clock stalls for N >= 2^24
    assert O.count_steps(np.float32(0.125) / np.float32(2 ** 26), 0.2) == -1
    assert O.count_steps(np.float32(0.125) / np.float32(100000), 0.2) == 159820


def test_minmod_and_mc_cases(oracle):
    L = oracle.lib()
    assert L.shll_oracle_minmod(1.0, 2.0) == 1.0
    assert L.shll_oracle_minmod(-3.0, -2.0) == -2.0
    assert L.shll_oracle_minmod(1.0, -2.0) == 0.0
    assert L.shll_oracle_minmod(2.0, 2.0) == 2.0          # tie -> right
    assert L.shll_oracle_minmod(0.0, 5.0) == 0.0          # product 0 is not < 0 -> smaller magnitude
    # underflowed product (+-0) takes the else branch: denormals matter
    tiny = float(np.float32(1e-30))
    assert L.shll_oracle_minmod(tiny, -tiny) == np.float32(-tiny)
    assert L.shll_oracle_mc(0.0, 1.0, 2.0, 1.25) == 1.0    # central 1.0 vs alpha*1 = 1.25
    assert L.shll_oracle_mc(0.0, 1.0, 4.0, 1.25) == 1.25   # central 2.0 vs alpha*min(1,3)=1.25


def test_threads_do_not_change_bits(oracle):
    O = oracle
    cfg1 = O.make_cfg(2, 48, 40, order=2, bc=O.BC_OUTFLOW, nthreads=1)
    cfg4 = O.make_cfg(2, 48, 40, order=2, bc=O.BC_OUTFLOW, nthreads=4)
    u0 = O.cons_from_prim(cfg1, O.init_prim(cfg1, O.IC_FOUR_SHOCK))
    assert np.array_equal(bits(O.run(cfg1, u0, 30)), bits(O.run(cfg4, u0, 30)))


@pytest.mark.parametrize("case", ["2d_o1_64", "2d_o1_96x160", "2d_o2_64", "2d_o2_96x160", "omp_o2_64", "1d_o1_1024"])
def test_fast_tolerance_is_anchored_to_the_reference_sensitivity(case, manifest, oracle):
    """The FAST-mode tolerance of the CUDA path (conftest.FAST_TOL) is stated against how much the reference's own
    arithmetic moves when the compiler is allowed to contract a*b+c into FMAs (the default on the Graviton builds the
    README reports): never tighter than 2x that, never looser than max(3e-5, 5x that)."""
    O = oracle
    pb = problem_from_manifest(manifest[case])
    cfg = oracle_cfg_for(O, pb, nthreads=2)
    gu, gp, gsteps = load_golden(case)
    u0 = O.cons_from_prim(cfg, O.init_prim(cfg, ORACLE_IC[pb.ic]))
    uf = O.run_fma_contracted(cfg, u0, gsteps)
    if uf is None:
        pytest.skip("host CPU has no FMA")
    pf = O.prim_from_cons(cfg, uf).astype(np.float64)
    sens = float((np.abs(pf - gp) / (1.0 + np.abs(gp.astype(np.float64)))).max())
    tol = FAST_TOL.get(case, FAST_TOL_DEFAULT)
    assert tol >= 2.0 * sens, f"{case}: tolerance {tol:g} is tighter than 2x the reference's own FMA sensitivity {sens:.2e}"
    assert tol <= max(FAST_TOL_DEFAULT, 5.0 * sens), f"{case}: tolerance {tol:g} is looser than 5x the sensitivity {sens:.2e}"


def test_long_run_fast_tolerances_are_anchored():
    """conftest.FAST_TOL_LONG: each entry at least 2x the reference's own FMA-contraction sensitivity on that run (recorded in the
    fixture by tests/golden/make_golden_long.py) and at most max(3e-5, 5x) of it."""
    import json
    from conftest import FAST_TOL_LONG, GOLDEN_DIR
    meta = json.load(open(os.path.join(GOLDEN_DIR, "MANIFEST_LONG.json")))
    assert set(meta) == set(FAST_TOL_LONG) - {"long_1d_o2_65536"}
    for case, tol in FAST_TOL_LONG.items():
        if case not in meta:
            continue
        sens = float(np.load(os.path.join(GOLDEN_DIR, case + ".npz"))["ref_fma"])
        assert abs(sens - meta[case]["ref_fma_sensitivity"]) < 1e-12
        assert tol >= 2.0 * sens or tol >= FAST_TOL_DEFAULT, (case, tol, sens)
        assert tol <= max(FAST_TOL_DEFAULT, 5.0 * sens), (case, tol, sens)
