"""GPU (-m gpu): the CUDA path, called through the C ABI, against the golden fixtures and the oracle.

STRICT mode must be bit-exact (0 ULP) on every conserved component.  FAST mode must satisfy the tolerance stated in
BASELINE.md / SURVEY.md App. A on the parity configs: |x - ref| <= 3e-5 + 3e-5*|ref| on every primitive field.
Nothing here reads /root/reference: the fixtures are committed, the oracle is built from oracle/shll_oracle.c.
"""
from dataclasses import replace

import numpy as np
import pytest

from conftest import FAST_TOL, FAST_TOL_DEFAULT, ORACLE_IC, bits, load_golden, oracle_cfg_for, problem_from_manifest
from shll_sve_cfd_b200 import capi, programs

pytestmark = pytest.mark.gpu

GOLDEN = ["1d_o1_256", "1d_o1_1024", "2d_o1_64", "2d_o1_96x160", "2d_o1_256", "2d_o2_64", "2d_o2_96x160",
          "1d_o2_slice_1024", "omp_o2_64"]
# FAST-mode tolerance: conftest.FAST_TOL_DEFAULT / FAST_TOL (stated there, anchored to the reference's own sensitivity to
# FMA contraction by tests/test_oracle_golden.py).


def _pb(case, manifest):
    pb = problem_from_manifest(manifest[case])
    if case.startswith("1d_o2_slice"):
        pb = pb.resized(manifest[case]["nx"])
    return pb


@pytest.mark.parametrize("case", GOLDEN)
def test_strict_mode_is_bit_exact_vs_reference_golden(case, manifest):
    pb = _pb(case, manifest)
    gu, gp, gsteps = load_golden(case)
    r = programs.run_program(pb, capi.MODE_STRICT)
    assert r["steps"] == gsteps
    assert r["launches"] >= 1
    bad = int((bits(r["u"]) != bits(gu)).sum())
    assert bad == 0, f"{case}: {bad} words differ, max |du| = {np.abs(r['u'] - gu).max()} ({r['variant']})"
    assert np.array_equal(bits(r["p"]), bits(gp)), "device-side Compute_P_from_U differs"


@pytest.mark.parametrize("case", GOLDEN)
def test_fast_mode_within_stated_tolerance(case, manifest):
    pb = _pb(case, manifest)
    gu, gp, gsteps = load_golden(case)
    r = programs.run_program(pb, capi.MODE_FAST)
    err = np.abs(r["p"].astype(np.float64) - gp.astype(np.float64))
    t = FAST_TOL.get(case, FAST_TOL_DEFAULT)
    tol = t + t * np.abs(gp.astype(np.float64))
    assert np.isfinite(r["p"]).all()
    assert (err <= tol).all(), f"{case}: max err {err.max():.3e}, worst ratio {(err / tol).max():.2f} ({r['variant']})"


def _random_state(pb, seed):
    """Smooth random positive state: not one of the reference's ICs, exercises every limiter branch."""
    rng = np.random.default_rng(seed)
    n = pb.ncells
    if pb.dims == 1:
        x = np.linspace(0, 1, n, dtype=np.float64)
        rho = 1.0 + 0.5 * np.sin(2 * np.pi * (3 * x + rng.random())) + 0.2 * rng.random(n)
        ux = 0.6 * np.sin(2 * np.pi * (2 * x + rng.random())) + 0.1 * rng.standard_normal(n)
        T = 1.0 + 0.3 * np.cos(2 * np.pi * (x + rng.random())) + 0.1 * rng.random(n)
        p = np.stack([rho, ux, T]).astype(np.float32)
    else:
        x = np.linspace(0, 1, pb.nx)[:, None]
        y = np.linspace(0, 1, pb.ny)[None, :]
        rho = 1.0 + 0.5 * np.sin(2 * np.pi * (2 * x + y + rng.random())) + 0.2 * rng.random((pb.nx, pb.ny))
        ux = 0.7 * np.sin(2 * np.pi * (x - 2 * y + rng.random())) + 0.1 * rng.standard_normal((pb.nx, pb.ny))
        uy = 0.7 * np.cos(2 * np.pi * (3 * x + y + rng.random())) + 0.1 * rng.standard_normal((pb.nx, pb.ny))
        T = 1.0 + 0.3 * np.cos(2 * np.pi * (x + y + rng.random())) + 0.1 * rng.random((pb.nx, pb.ny))
        p = np.stack([rho, ux, uy, T]).astype(np.float32).reshape(4, n)
    return programs.cons_from_prim(pb, p)


def _gpu_vs_oracle(pb, O, u0, steps, mode=capi.MODE_STRICT, variant=0):
    cfg = oracle_cfg_for(O, pb, nthreads=4)
    ref = O.run(cfg, u0, steps)
    with programs.make_solver(pb, mode, variant=variant) as s:
        s.upload_u(u0)
        s.run(steps)
        got = s.download_u()
        name = s.variant
    return got, ref, name


SCHEMES = {
    "1d_o1_reflect": programs.BASE_SHLL,
    "1d_o1_outflow": programs.Problem("x", 1, 256, order=1, bc=capi.BC_OUTFLOW),
    "1d_o2_outflow": programs.SECOND_ORDER_1D,
    "1d_o2_reflect_mc": programs.Problem("x", 1, 256, order=2, bc=capi.BC_REFLECT, limiter=capi.LIM_MC, tform=capi.TFORM_2D),
    "1d_o2_tform1d": programs.Problem("x", 1, 256, order=2, bc=capi.BC_OUTFLOW, tform=capi.TFORM_1D),
    "2d_o1_reflect": programs.BASE_SHLL_2D,
    "2d_o1_outflow": programs.Problem("x", 2, 64, 64, order=1, bc=capi.BC_OUTFLOW, ic="implosion"),
    "2d_o2_outflow": programs.SECOND_ORDER_2D,
    "2d_o2_reflect": programs.Problem("x", 2, 64, 64, order=2, bc=capi.BC_REFLECT, ic="four_shock"),
    "2d_o2_outflow_mc": programs.BASE_OMP_2D,
    "2d_o2_reflect_mc": programs.Problem("x", 2, 64, 64, order=2, bc=capi.BC_REFLECT, limiter=capi.LIM_MC, ic="config6"),
}


@pytest.mark.parametrize("scheme", sorted(SCHEMES))
def test_random_state_strict_bit_exact_all_scheme_combinations(scheme, oracle):
    base = SCHEMES[scheme]
    pb = base.resized(777) if base.dims == 1 else base.resized(70, 90)
    u0 = _random_state(pb, seed=hash(scheme) % 1000)
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 25)
    bad = int((bits(got) != bits(ref)).sum())
    assert bad == 0, f"{scheme} ({name}): {bad} words differ, max |du| {np.abs(got - ref).max()}"


# Ragged / edge shapes: tile remainders, N not a multiple of 4 or of the warp tile, tiny grids, single-chunk rows.
SHAPES_1D = [2, 3, 5, 119, 120, 121, 127, 240, 241, 1000, 4099]
SHAPES_2D = [(2, 2), (3, 5), (2, 64), (64, 2), (33, 31), (65, 30), (30, 61), (129, 124), (40, 257)]


@pytest.mark.parametrize("n", SHAPES_1D)
@pytest.mark.parametrize("order", [1, 2])
def test_1d_ragged_sizes(n, order, oracle):
    pb = (programs.BASE_SHLL if order == 1 else programs.SECOND_ORDER_1D).resized(n)
    u0 = _random_state(pb, seed=n)
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 9)
    assert np.array_equal(bits(got), bits(ref)), f"N={n} order={order} {name}"


@pytest.mark.parametrize("shape", SHAPES_2D, ids=[f"{a}x{b}" for a, b in SHAPES_2D])
@pytest.mark.parametrize("order", [1, 2])
def test_2d_ragged_sizes(shape, order, oracle):
    pb = (programs.BASE_SHLL_2D if order == 1 else programs.SECOND_ORDER_2D).resized(*shape)
    pb = replace(pb, lx=shape[0] / shape[1])   # DX == DY, so DT_ON_DY == 0.125 and the scheme stays stable
    u0 = _random_state(pb, seed=shape[0] * 1000 + shape[1])
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 7)
    assert np.array_equal(bits(got), bits(ref)), f"{shape} order={order} {name}"


@pytest.mark.parametrize("tma", [0, 1])
@pytest.mark.parametrize("vec", [1, 2, 4])
@pytest.mark.parametrize("order", [1, 2])
def test_2d_every_kernel_variant_gives_the_same_bits(vec, order, tma, oracle, monkeypatch):
    """LDG kernel (1/2/4 cells per thread) and TMA-fed kernel: different data paths, identical bits."""
    if (tma == 0 and order == 2 and vec == 4) or (tma == 1 and vec > 2):
        pytest.skip("variant not instantiated")
    monkeypatch.setenv("SHLL_TMA", str(tma))
    pb = replace((programs.BASE_SHLL_2D if order == 1 else programs.SECOND_ORDER_2D).resized(72, 256), lx=72 / 256)
    u0 = _random_state(pb, seed=vec)
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 11, variant=vec)
    assert f"vec{vec}" in name and (("_tma_" in name) == bool(tma)), name
    bad = int((bits(got) != bits(ref)).sum())
    assert bad == 0, f"{name}: {bad} words differ"


@pytest.mark.parametrize("stages", [2, 3, 7])
@pytest.mark.parametrize("rows_per_chunk", [2, 5, 9, 1000])
def test_2d_tma_ring_geometry_does_not_change_bits(stages, rows_per_chunk, oracle, monkeypatch):
    """Any ring depth / chunk height (including chunks shorter than a box and a single chunk) gives the same result."""
    monkeypatch.setenv("SHLL_TMA_STAGES", str(stages))
    monkeypatch.setenv("SHLL_ROWS_PER_CHUNK", str(rows_per_chunk))
    for order in (1, 2):
        pb = replace((programs.BASE_SHLL_2D if order == 1 else programs.SECOND_ORDER_2D).resized(41, 128), lx=41 / 128)
        u0 = _random_state(pb, seed=stages * 10 + order)
        got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 6)
        assert "_tma_" in name
        assert np.array_equal(bits(got), bits(ref)), name


def test_general_dt_ratio_uses_the_exact_double_update(oracle):
    """NX != NY with L == H: DT_ON_DY is not a power of two, the 2nd-order update needs the literal double expression."""
    pb = programs.SECOND_ORDER_2D.resized(48, 80)
    u0 = _random_state(pb, seed=5)
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, 40)
    assert "gendt" in name
    assert np.array_equal(bits(got), bits(ref))


def test_denormal_slopes_are_not_flushed(oracle):
    """minmod's sign test is on the rounded float product; an underflowed product must take the else branch."""
    pb = programs.SECOND_ORDER_1D.resized(240)
    p = np.stack([np.ones(240), np.zeros(240), np.ones(240)]).astype(np.float32)
    p[0, 100:140] += (np.arange(40) * 1e-22).astype(np.float32)   # tiny ramps -> denormal-product slopes
    p[2, 60:90] -= (np.arange(30) * 3e-23).astype(np.float32)
    u0 = programs.cons_from_prim(pb, p)
    got, ref, _ = _gpu_vs_oracle(pb, oracle, u0, 12)
    assert np.array_equal(bits(got), bits(ref))


def test_fast_mode_close_to_strict_on_random_state(oracle):
    pb = programs.SECOND_ORDER_2D.resized(96, 128)
    u0 = _random_state(pb, seed=9)
    got, ref, _ = _gpu_vs_oracle(pb, oracle, u0, 50, mode=capi.MODE_FAST)
    assert np.isfinite(got).all()
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert (err <= 1e-4 + 1e-4 * np.abs(ref)).all(), err.max()


@pytest.mark.parametrize("n", [64, 100, 1000, 4097, 65536, 146000])
@pytest.mark.parametrize("scheme", ["1d_o1_reflect", "1d_o2_outflow", "1d_o2_reflect_mc"])
def test_persistent_1d_march_is_bit_exact(n, scheme, oracle):
    """The register-resident persistent kernel (one cooperative launch for all steps) against the oracle."""
    pb = SCHEMES[scheme].resized(n)
    u0 = _random_state(pb, seed=n % 97)
    steps = 77 if n <= 4097 else 41
    got, ref, name = _gpu_vs_oracle(pb, oracle, u0, steps)
    assert np.array_equal(bits(got), bits(ref)), f"N={n} {scheme}"


@pytest.mark.parametrize("K", [1, 3, 8, 16])
def test_persistent_1d_any_round_length_and_both_modes(K, oracle, monkeypatch):
    monkeypatch.setenv("SHLL_PERSIST_K", str(K))
    pb = programs.SECOND_ORDER_1D.resized(30000)
    u0 = _random_state(pb, seed=K)
    got, ref, _ = _gpu_vs_oracle(pb, oracle, u0, 100)
    assert np.array_equal(bits(got), bits(ref))
    # FAST mode: the persistent kernel runs, operation for operation, the arithmetic of the streaming kernel (face-flux form,
    # step1d_acc.cuh) -> identical bits, and both stay within the FAST tolerance of the oracle
    with programs.make_solver(pb, capi.MODE_FAST) as s:
        s.upload_u(u0); s.run(100); a = s.download_u(); launches_persist = s.launches
    monkeypatch.setenv("SHLL_PERSIST", "0")
    monkeypatch.setenv("SHLL_GRAPH", "0")
    with programs.make_solver(pb, capi.MODE_FAST) as s:
        s.upload_u(u0); s.run(100); b = s.download_u(); launches_stream = s.launches; name = s.variant
    assert "_acc_" in name and launches_persist == 1 and launches_stream == 50     # (streaming: two steps per launch)
    assert np.array_equal(bits(a), bits(b))
    assert (np.abs(b.astype(np.float64) - ref) <= 3e-5 + 3e-5 * np.abs(ref)).all(), np.abs(b - ref).max()


def test_persistent_and_streaming_fast_kernels_agree_bitwise_with_walls_and_mc(monkeypatch):
    """Reflective walls + MC limiter through both 1D FAST 2nd-order kernels (persistent: scalar; streaming: packed pairs)."""
    pb = programs.Problem("x", 1, 20000, order=2, bc=capi.BC_REFLECT, limiter=capi.LIM_MC, tform=capi.TFORM_2D)
    u0 = _random_state(pb, seed=11)
    with programs.make_solver(pb, capi.MODE_FAST) as s:
        s.upload_u(u0); s.run(64); a = s.download_u(); assert s.launches == 1
    monkeypatch.setenv("SHLL_PERSIST", "0")
    monkeypatch.setenv("SHLL_GRAPH", "0")
    with programs.make_solver(pb, capi.MODE_FAST) as s:
        s.upload_u(u0); s.run(64); b = s.download_u(); assert s.launches == 32 and "_acc_" in s.variant   # two steps per launch
    assert np.array_equal(bits(a), bits(b))


def test_cuda_graph_replay_small_grid_matches_single_launches(oracle, monkeypatch):
    monkeypatch.setenv("SHLL_PERSIST", "0")
    """Launch-bound regime: shll_run replays a captured graph of 128 steps; odd / split step counts must not matter."""
    pb = programs.SECOND_ORDER_1D.resized(4096)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    ref = oracle.run(oracle_cfg_for(oracle, pb, nthreads=4), u0, 3 + 700 + 1 + 300)
    with programs.make_solver(pb) as s:
        s.upload_u(u0)
        s.run(3)        # single launches
        s.run(700)      # graph (5 x 128) + 60 single
        s.run(1)
        s.run(300)      # realign + graph
        got = s.download_u()
        assert s.launches >= 1004
    assert np.array_equal(bits(got), bits(ref))
    monkeypatch.setenv("SHLL_GRAPH", "0")
    with programs.make_solver(pb) as s:
        s.upload_u(u0)
        s.run(1004)
        assert s.launches == 1004
        assert np.array_equal(bits(s.download_u()), bits(ref))


@pytest.mark.parametrize("pb", [programs.BASE_SHLL_2D.resized(256, 256), programs.SECOND_ORDER_2D.resized(96, 160),
                                programs.SECOND_ORDER_1D.resized(200000), programs.BASE_SHLL.resized(200000)],
                         ids=["2d_o1", "2d_o2", "1d_o2", "1d_o1"])
def test_programmatic_dependent_launch_does_not_change_bits(pb, monkeypatch):
    """Every step kernel is launched so that its blocks are placed during the predecessor's tail and wait in
    griddepcontrol.wait (halo_sync.cuh).  Same bits with the attribute off, in both modes; small grids use thinner chunks."""
    u0 = _random_state(pb, seed=3)
    monkeypatch.setenv("SHLL_GRAPH", "0")
    monkeypatch.setenv("SHLL_PERSIST", "0")
    for mode in (capi.MODE_STRICT, capi.MODE_FAST):
        outs = []
        for pdl in ("1", "0"):
            monkeypatch.setenv("SHLL_PDL", pdl)
            with programs.make_solver(pb, mode) as s:
                s.upload_u(u0)
                s.run(37)
                outs.append(s.download_u())
                assert s.launches == (19 if s.variant.endswith("_x2") else 37)   # (two steps per launch: 18 + the odd one)
        assert np.array_equal(bits(outs[0]), bits(outs[1]))


def test_small_2d_grids_get_thinner_chunks():
    """plan_2d shrinks the chunk height until the launch has enough blocks for the kernel at hand."""
    def chunks(pb, mode):
        with programs.make_solver(pb, mode) as s:
            return int(s.variant.split("_chunks")[1].split("_")[0])
    assert chunks(programs.SECOND_ORDER_2D.resized(256, 256), capi.MODE_FAST) == 64      # 4-row chunks instead of 64-row ones
    assert chunks(programs.BASE_SHLL_2D.resized(256, 256), capi.MODE_FAST) == 64         # 4-row chunks
    assert chunks(programs.BASE_SHLL_2D.resized(4096, 4096), capi.MODE_FAST) == 94       # the tuned 44 rows (two-step launches) at configs[2]
    assert chunks(programs.SECOND_ORDER_2D.resized(2048, 16384), capi.MODE_FAST) == 32   # the tuned 64 rows at configs[4]
    assert chunks(programs.BASE_SHLL_2D.resized(1024, 1024), capi.MODE_FAST) == 64       # 16-row chunks: 0.6 of a resident wave of the two-step kernel
    assert chunks(programs.BASE_SHLL_2D.resized(1024, 1024), capi.MODE_STRICT) == 171    # 6-row chunks: the bit-exact rows are ~3x longer
    assert chunks(programs.BASE_SHLL_2D.resized(4096, 4096), capi.MODE_STRICT) == 171    # the tuned 24 rows at full size
    assert chunks(programs.SECOND_ORDER_2D.resized(1024, 1024), capi.MODE_STRICT) == 128  # 8-row chunks


def test_cfl_diagnostic_and_api_state_errors(oracle):
    pb = programs.BASE_SHLL_2D.resized(64, 64)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    with programs.make_solver(pb) as s:
        with pytest.raises(capi.ShllError) as ei:
            s.run(1)
        assert ei.value.code == capi.E_STATE
        s.upload_u(u0)
        p, a = programs.prim_from_cons(pb, u0)
        want = float(np.max((np.maximum(np.abs(p[1]), np.abs(p[2])) + a) * np.float32(0.125)))
        assert abs(s.max_cfl() - want) <= 1e-6
        sums = s.conserved_sums()                          # mass, x / y momentum, energy: FP64, fixed order
        want_sums = u0.astype(np.float64).sum(axis=1)
        assert np.allclose(sums, want_sums, rtol=1e-12, atol=1e-9), (sums, want_sums)
        assert np.array_equal(sums, s.conserved_sums()), "fixed summation order: reproducible"
        before = s.download_u()
        assert np.array_equal(bits(before), bits(u0))     # the monitors never change the state or DT
        s.run(0)
        assert np.array_equal(bits(s.download_u()), bits(u0))


# ------------------------------------------------------------------------- full-size, size-independent properties

def test_4096x4096_conserves_mass_and_keeps_symmetry_properties():
    """configs[2] at full size (base_shll_2d.c, 4096^2): reflective walls conserve total mass and energy up to
    float rounding; the implosion IC is symmetric under (i,j) -> (N-1-i, N-1-j) only approximately because the
    reference applies X then Y updates, so we check conservation and positivity, plus bitwise repeatability."""
    pb = programs.BASE_SHLL_2D.resized(4096, 4096)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    with programs.make_solver(pb) as s:
        s.upload_u(u0)
        s.run(40)
        u1 = s.download_u()
        s.upload_u(u0)
        s.run(40)
        u2 = s.download_u()
        dev_sums = s.conserved_sums()
    assert np.array_equal(bits(u1), bits(u2)), "same input, same bits"
    m0, m1 = u0[0].astype(np.float64).sum(), u1[0].astype(np.float64).sum()
    e0, e1 = u0[3].astype(np.float64).sum(), u1[3].astype(np.float64).sum()
    assert abs(m1 - m0) / m0 < 1e-6 and abs(e1 - e0) / e0 < 1e-6
    assert np.allclose(dev_sums, u2.astype(np.float64).sum(axis=1), rtol=1e-12, atol=1e-6)   # device-side monitor, full size
    assert (u1[0] > 0).all() and np.isfinite(u1).all()


def test_4096x4096_prefix_matches_oracle_on_a_window(oracle):
    """A step-capped prefix at full size: the oracle advances only a window that the true solution at the window
    centre cannot distinguish from the full domain (finite speed of the stencil: 1 cell per step)."""
    pb = programs.BASE_SHLL_2D.resized(4096, 4096)
    steps, W = 10, 96
    i0 = j0 = int(0.2 * 4096) - W // 2                     # straddles the corner of the low-density box
    p0 = programs.initial_primitives(pb).reshape(4, 4096, 4096)
    u0 = programs.cons_from_prim(pb, p0.reshape(4, -1)).reshape(4, 4096, 4096)
    with programs.make_solver(pb) as s:
        s.upload_u(u0.reshape(4, -1))
        s.run(steps)
        got = s.download_u().reshape(4, 4096, 4096)
    sub = programs.Problem("w", 2, W, W, order=1, bc=capi.BC_OUTFLOW, ic="implosion")
    cfg = oracle_cfg_for(oracle, sub)
    ref = oracle.run(cfg, np.ascontiguousarray(u0[:, i0:i0 + W, j0:j0 + W]).reshape(4, -1), steps).reshape(4, W, W)
    m = steps + 1
    a = got[:, i0 + m:i0 + W - m, j0 + m:j0 + W - m]
    b = ref[:, m:W - m, m:W - m]
    assert np.array_equal(bits(a), bits(b))


def test_1d_64m_cells_fixed_step_count_window_vs_oracle(oracle):
    """configs[3] size (2^26 cells, fixed step count because the float clock stalls): window check around the diaphragm."""
    n, steps, W = 1 << 26, 6, 4096
    pb = programs.SECOND_ORDER_1D.resized(n)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    with programs.make_solver(pb) as s:
        s.upload_u(u0)
        s.run(steps)
        got = s.download_u()
    i0 = n // 2 - W // 2
    sub = programs.SECOND_ORDER_1D.resized(W)
    cfg = oracle_cfg_for(oracle, sub)
    ref = oracle.run(cfg, np.ascontiguousarray(u0[:, i0:i0 + W]), steps)
    m = 2 * steps + 2
    assert np.array_equal(bits(got[:, i0 + m:i0 + W - m]), bits(ref[:, m:W - m]))
    # far from the diaphragm nothing may have changed
    assert np.array_equal(bits(got[:, :1000]), bits(u0[:, :1000]))


def test_exactness_shortcuts_device_selftest():
    """STRICT mode never executes nvcc's division expansion on the hot path (csrc/shll_math.cuh: div_rn_shared / div_rn_spec /
    div_by_cv / div_by_cv_spec).  csrc/selftest.cu runs those sequences on the device against __fdiv_rn / __ddiv_rn over 2^28
    operand pairs (uniform bit patterns, physical magnitudes, the guard edges 2^+-60 / 2^+-40, denormals, signed zeros,
    infinities, NaNs): not one bit may differ.  The exact-rational-arithmetic twin is in tests/test_host_logic.py."""
    r = capi.selftest_exact_division(1 << 28, seed=2026)
    assert r["float_pairs"] >= (1 << 28) and r["doubles"] >= (1 << 28), r
    assert r["div_rn_shared_mismatch"] == 0 and r["div_rn_spec_mismatch"] == 0, r
    assert r["div_by_cv_mismatch"] == 0 and r["div_by_cv_spec_mismatch"] == 0, r
    # the guards are exercised on both sides: at least 2^26 draws stay on each fast path, at least 2^20 leave it
    for flagged, total in (("float_flagged", "float_pairs"), ("float_slow_path", "float_pairs"), ("double_flagged", "doubles")):
        assert (1 << 20) < r[flagged] < r[total] - (1 << 26), (flagged, r)
    r2 = capi.selftest_exact_division(1 << 22, seed=7)
    assert sum(r2[k] for k in r2 if k.endswith("mismatch")) == 0, r2


def test_config1_65536_cells_run_to_completion_is_bit_exact():
    """BASELINE.json configs[1] exactly: the derived 1D 2nd-order program at 65 536 cells, all 104 858 steps to t = 0.2 in ONE
    persistent cooperative launch, STRICT mode, bit for bit against the full-length fixture (tests/golden/make_golden_config1.py);
    FAST mode in the mean / outlier-share sense stated in conftest.py (FAST_LONG_1D_*)."""
    import os
    from conftest import FAST_LONG_1D_MEAN_TOL, FAST_LONG_1D_SHARE_ABOVE, FAST_TOL_LONG, GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "long_1d_o2_65536.npz"))
    pb = programs.SECOND_ORDER_1D.resized(65536)
    r = programs.run_program(pb, capi.MODE_STRICT)
    assert r["steps"] == 104858 == int(z["steps"])
    assert r["launches"] <= 2, r["launches"]          # the march itself is one launch (+ the device-side Compute_P_from_U)
    bad = int((bits(r["u"]) != bits(z["u"])).sum())
    assert bad == 0, f"{bad} words differ after 104858 steps, max |du| = {np.abs(r['u'] - z['u']).max()}"
    assert np.array_equal(bits(r["p"]), bits(z["p"]))
    # FAST mode after 104 858 steps: the smooth parts agree to rounding drift, but a discontinuity that lands one cell to the side
    # is an O(jump) pointwise difference, so the max norm says nothing here ("long runs should be judged in strict
    # mode or on prefix checkpoints").  Stated instead: the mean error and the share of cells outside the long-run tolerance.
    f = programs.run_program(pb, capi.MODE_FAST)
    ref = z["p"].astype(np.float64)
    err = np.abs(f["p"].astype(np.float64) - ref) / (1.0 + np.abs(ref))
    l1, worst, share = float(err.mean()), float(err.max()), float((err.max(axis=0) > FAST_TOL_LONG["long_1d_o2_65536"]).mean())
    e = err.max(axis=0)
    print(f"[fast long run] long_1d_o2_65536: 104858 steps, {f['variant']}: mean |dp|/(1+|p|) = {l1:.3e}, max = {worst:.3e}, "
          f"cells outside {FAST_TOL_LONG['long_1d_o2_65536']:g}: {share * 100:.4f} %; share of cells above 1e-4 / 3e-4 / 1e-3 / 1e-2: "
          f"{(e > 1e-4).mean():.4f} / {(e > 3e-4).mean():.4f} / {(e > 1e-3).mean():.4f} / {(e > 1e-2).mean():.5f}")
    assert np.isfinite(f["p"]).all()
    assert l1 <= FAST_LONG_1D_MEAN_TOL, l1
    for level, most in FAST_LONG_1D_SHARE_ABOVE.items():
        assert float((e > level).mean()) <= most, (level, float((e > level).mean()))


@pytest.mark.parametrize("early_blocks", [1, 8, 40, 100000])
def test_early_start_of_the_first_wave_any_size_of_the_wave_same_bits(early_blocks, monkeypatch):
    """The blocks of a launch's first wave start on the flags of the blocks of the previous launch they depend on instead of
    griddepcontrol.wait (csrc/step2d_tma.cuh, csrc/step1d.cuh).  Forced on small grids with every size of the early set --
    smaller than a row of tiles, a few chunk rows, the whole grid: same bits as the plain wait, and no flag is ever waited for
    that nobody publishes (a missing publisher would show as the 2 s poll timeout -> SHLL_E_TIMEOUT)."""
    import time
    monkeypatch.setenv("SHLL_GRAPH", "0")
    monkeypatch.setenv("SHLL_PERSIST", "0")
    cases = [replace(programs.BASE_SHLL_2D.resized(96, 512), lx=96 / 512), replace(programs.SECOND_ORDER_2D.resized(96, 512), lx=96 / 512),
             programs.SECOND_ORDER_1D.resized(200000)]
    for pb in cases:
        u0 = _random_state(pb, seed=early_blocks % 97)
        for mode in (capi.MODE_STRICT, capi.MODE_FAST):
            outs = []
            for early in ("0", "2"):
                monkeypatch.setenv("SHLL_EARLY", early)
                monkeypatch.setenv("SHLL_EARLY_BLOCKS", str(early_blocks))
                t0 = time.perf_counter()
                with programs.make_solver(pb, mode) as s:
                    s.upload_u(u0)
                    s.run(21)
                    outs.append(s.download_u())
                assert time.perf_counter() - t0 < 1.5, f"{pb.name} early={early}: a block waited for a flag nobody publishes"
            assert np.array_equal(bits(outs[0]), bits(outs[1])), f"{pb.name} mode {mode} early_blocks {early_blocks}"
