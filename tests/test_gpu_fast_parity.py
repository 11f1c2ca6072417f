"""GPU (-m gpu): the FAST kernels -- the ones bench.py times -- pinned against the oracle IN THE BENCHED GEOMETRY.

STRICT mode is bit-exact and covered by test_gpu_parity.py.  FAST mode (FP32 + explicit FMA, face-flux form) is
tolerance-matched; this file makes sure the tolerance claim is made about the kernels, tile shapes and chunk heights the
headline numbers come from, not about small-grid stand-ins:

  * step2d_acc at 4096^2 / 18-row chunks (configs[2]), at 2048 x 16384 / 64-row chunks (configs[4] slab) and step1d_acc at
    2^26 cells (configs[3]): a window of the full-size run against the oracle run on that window;
  * ring depth / chunk height, ragged shapes and all 11 scheme combinations in FAST mode: bitwise independent of the
    geometry, within the FAST tolerance of the oracle;
  * full-length FAST runs (820 and 1639 steps) against the oracle, with the tolerance that actually holds
    (conftest.FAST_TOL_LONG -- one table, quoted by README.md and bench.py).

Tolerances are on the primitive fields (rho, u, v, T), |x - ref| <= tol * (1 + |ref|), as north_star states it.
"""
import zlib
from dataclasses import replace

import numpy as np
import pytest

from conftest import FAST_TOL_DEFAULT, FAST_TOL_LONG, bits, oracle_cfg_for
from shll_sve_cfd_b200 import capi, programs
from test_gpu_parity import SCHEMES, SHAPES_1D, SHAPES_2D, _random_state

pytestmark = pytest.mark.gpu


def _prim_err(pb, got_u, ref_u):
    """max over fields and cells of |p - p_ref| / (1 + |p_ref|) on the primitive fields, and the absolute maximum"""
    pg, _ = programs.prim_from_cons(pb, np.ascontiguousarray(got_u, dtype=np.float32))
    pr, _ = programs.prim_from_cons(pb, np.ascontiguousarray(ref_u, dtype=np.float32))
    assert np.isfinite(pg).all()
    err = np.abs(pg.astype(np.float64) - pr.astype(np.float64))
    return float((err / (1.0 + np.abs(pr.astype(np.float64)))).max()), float(err.max())


def _fast_run(pb, u0, steps, variant=0):
    with programs.make_solver(pb, capi.MODE_FAST, variant=variant) as s:
        s.upload_u(u0)
        s.run(steps)
        return s.download_u(), s.variant, s.launches


# ------------------------------------------------------------------------------------ the benched kernels, full size

@pytest.mark.parametrize("steps", [10, 40])
def test_fast_4096x4096_benched_geometry_window_vs_oracle(steps, oracle):
    """configs[2] exactly as bench.py runs it: step2d_acc2 (two steps per launch), 60-column tiles, 44-row chunks (94 chunks), FAST."""
    n, W = 4096, 192
    pb = programs.BASE_SHLL_2D.resized(n, n)
    i0 = j0 = int(0.2 * n) - W // 2                      # straddles the corner of the low-density box
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb)).reshape(4, n, n)
    got, name, launches = _fast_run(pb, u0.reshape(4, -1), steps)
    assert "_tma_acc" in name and "chunks94" in name and "fast" in name and name.endswith("_x2"), name
    assert launches == steps // 2
    got = got.reshape(4, n, n)
    sub = programs.Problem("w", 2, W, W, order=1, bc=capi.BC_OUTFLOW, ic="implosion")
    ref = oracle.run(oracle_cfg_for(oracle, sub, nthreads=8), np.ascontiguousarray(u0[:, i0:i0 + W, j0:j0 + W]).reshape(4, -1), steps).reshape(4, W, W)
    m = steps + 1                                        # the window's own boundary pollutes one cell per step
    a = np.ascontiguousarray(got[:, i0 + m:i0 + W - m, j0 + m:j0 + W - m]).reshape(4, -1)
    b = np.ascontiguousarray(ref[:, m:W - m, m:W - m]).reshape(4, -1)
    rel, absmax = _prim_err(sub, a, b)
    assert rel <= FAST_TOL_DEFAULT, f"{name}: {rel:.3e} (abs {absmax:.3e}) after {steps} steps"
    # the window spans chunk seams (every ~43.6 rows: rows 741, 784, 828, 871 lie inside) and tile seams (columns 60k) of the benched geometry
    assert (i0 + m) * 94 // n != (i0 + W - m) * 94 // n and (j0 + m) // 60 != (j0 + W - m) // 60
    # far from the density jump the gas is at rest and must stay bit-for-bit at rest
    assert np.array_equal(bits(got[:, 2000:2100, 2000:2100]), bits(u0[:, 2000:2100, 2000:2100]))


def test_fast_2d_o2_2048x16384_benched_geometry_window_vs_oracle(oracle):
    """configs[4] slab exactly as bench.py runs it: 2nd order, minmod, outflow, 64-row chunks (32 chunks), FAST."""
    nx, ny, W, steps = 2048, 16384, 160, 12
    pb = replace(programs.SECOND_ORDER_2D.resized(nx, ny), lx=nx / ny, ly=1.0)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb)).reshape(4, nx, ny)
    got, name, _ = _fast_run(pb, u0.reshape(4, -1), steps)
    assert "_tma_acc" in name and "chunks32" in name, name
    got = got.reshape(4, nx, ny)
    i0, j0 = int(0.75 * nx) - W // 2, int(0.75 * ny) - W // 2   # the four-shock corner; rows 1456..1616 span the chunk seam at 1472, 1536, 1600
    sub = programs.Problem("w", 2, W, W, order=2, bc=capi.BC_OUTFLOW, ic="four_shock")
    ref = oracle.run(oracle_cfg_for(oracle, sub, nthreads=8), np.ascontiguousarray(u0[:, i0:i0 + W, j0:j0 + W]).reshape(4, -1), steps).reshape(4, W, W)
    m = 2 * steps + 2
    a = np.ascontiguousarray(got[:, i0 + m:i0 + W - m, j0 + m:j0 + W - m]).reshape(4, -1)
    b = np.ascontiguousarray(ref[:, m:W - m, m:W - m]).reshape(4, -1)
    rel, absmax = _prim_err(sub, a, b)
    assert rel <= FAST_TOL_DEFAULT, f"{name}: {rel:.3e} (abs {absmax:.3e})"


def test_fast_1d_64m_cells_benched_kernel_window_vs_oracle(oracle):
    """configs[3] at N=1 exactly as bench.py runs it: step1d_acc (cp.async ring, 8 tiles per warp), 2^26 cells, FAST."""
    n, steps, W = 1 << 26, 24, 8192
    pb = programs.SECOND_ORDER_1D.resized(n)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    got, name, launches = _fast_run(pb, u0, steps)
    assert "_acc_" in name and name.endswith("_x2") and launches == steps // 2, (name, launches)
    i0 = n // 2 - W // 2
    sub = programs.SECOND_ORDER_1D.resized(W)
    ref = oracle.run(oracle_cfg_for(oracle, sub), np.ascontiguousarray(u0[:, i0:i0 + W]), steps)
    m = 2 * steps + 2
    rel, absmax = _prim_err(sub, got[:, i0 + m:i0 + W - m], ref[:, m:W - m])
    assert rel <= FAST_TOL_DEFAULT, f"{name}: {rel:.3e} (abs {absmax:.3e})"
    assert np.array_equal(bits(got[:, :1000]), bits(u0[:, :1000]))


# ------------------------------------------------------------------------------------ geometry independence, FAST

@pytest.mark.parametrize("stages", [2, 3, 7])
@pytest.mark.parametrize("rows_per_chunk", [2, 5, 9, 18, 1000])
def test_fast_2d_ring_geometry_does_not_change_bits(stages, rows_per_chunk, oracle, monkeypatch):
    """FAST face-flux kernel: any ring depth / chunk height gives the bits of the default plan, and those are within the
    FAST tolerance of the oracle (the STRICT twin is test_2d_tma_ring_geometry_does_not_change_bits)."""
    for order in (1, 2):
        pb = replace((programs.BASE_SHLL_2D if order == 1 else programs.SECOND_ORDER_2D).resized(82, 128), lx=82 / 128)
        u0 = _random_state(pb, seed=stages * 10 + order)
        base, base_name, _ = _fast_run(pb, u0, 6)
        with monkeypatch.context() as mp:
            mp.setenv("SHLL_TMA_STAGES", str(stages))
            mp.setenv("SHLL_ROWS_PER_CHUNK", str(rows_per_chunk))
            got, name, _ = _fast_run(pb, u0, 6)
        assert "_tma_acc" in name and "_tma_acc" in base_name, (name, base_name)
        assert np.array_equal(bits(got), bits(base)), f"{name} vs {base_name}"
        ref = oracle.run(oracle_cfg_for(oracle, pb, nthreads=4), u0, 6)
        rel, absmax = _prim_err(pb, got, ref)
        assert rel <= FAST_TOL_DEFAULT, f"{name}: {rel:.3e}"


@pytest.mark.parametrize("n", SHAPES_1D)
@pytest.mark.parametrize("order", [1, 2])
def test_fast_1d_ragged_sizes(n, order, oracle):
    pb = (programs.BASE_SHLL if order == 1 else programs.SECOND_ORDER_1D).resized(n)
    u0 = _random_state(pb, seed=n)
    got, name, _ = _fast_run(pb, u0, 9)
    ref = oracle.run(oracle_cfg_for(oracle, pb, nthreads=2), u0, 9)
    rel, _ = _prim_err(pb, got, ref)
    assert rel <= FAST_TOL_DEFAULT, f"N={n} order={order} {name}: {rel:.3e}"


@pytest.mark.parametrize("shape", SHAPES_2D + [(37, 64), (130, 188), (19, 1024)], ids=lambda s: f"{s[0]}x{s[1]}")
@pytest.mark.parametrize("order", [1, 2])
def test_fast_2d_ragged_sizes(shape, order, oracle):
    """Tile remainders, ny % 4 != 0 (LDG kernel), ny % 8 != 0 (one cell per lane), tiny grids, ragged last tile of the
    face-flux kernel (188 = 3 x 60 + 8), single-chunk columns."""
    pb = (programs.BASE_SHLL_2D if order == 1 else programs.SECOND_ORDER_2D).resized(*shape)
    pb = replace(pb, lx=shape[0] / shape[1])
    u0 = _random_state(pb, seed=shape[0] * 1000 + shape[1])
    got, name, _ = _fast_run(pb, u0, 7)
    ref = oracle.run(oracle_cfg_for(oracle, pb, nthreads=4), u0, 7)
    rel, _ = _prim_err(pb, got, ref)
    assert rel <= FAST_TOL_DEFAULT, f"{shape} order={order} {name}: {rel:.3e}"


@pytest.mark.parametrize("scheme", sorted(SCHEMES))
def test_fast_random_state_all_scheme_combinations(scheme, oracle):
    """All 11 (dims, order, BC, limiter, T-form) combinations on a random smooth state, FAST; 2D shapes that select the
    face-flux kernel (ny % 8 == 0) so that its wall / MC-limiter / reflect instantiations are the ones checked."""
    base = SCHEMES[scheme]
    pb = base.resized(777) if base.dims == 1 else replace(base.resized(70, 96), lx=70 / 96)
    u0 = _random_state(pb, seed=zlib.crc32(scheme.encode()) % 1000)   # (str hash() is salted per process)
    got, name, _ = _fast_run(pb, u0, 25)
    if base.dims == 2:
        assert "_tma_acc" in name, name
    ref = oracle.run(oracle_cfg_for(oracle, pb, nthreads=4), u0, 25)
    rel, absmax = _prim_err(pb, got, ref)
    assert rel <= FAST_TOL_DEFAULT, f"{scheme} ({name}): {rel:.3e} (abs {absmax:.3e})"


@pytest.mark.parametrize("shape", [(64, 128), (41, 128), (130, 192), (257, 64), (96, 248)], ids=lambda s: f"{s[0]}x{s[1]}")
@pytest.mark.parametrize("bc", ["reflect", "outflow"])
def test_fused_two_step_launches_give_the_bits_of_single_steps(shape, bc, monkeypatch):
    """The 1st-order FAST kernel advances two steps per launch (csrc/step2d_acc.cuh: step2d_acc2_kernel): stage A's rows of
    U^(n+1) go to stage B in registers.  Same bits as one launch per step -- walls, ragged last tile, thin / uneven chunks, odd
    step counts (the last step is a one-step launch), split runs -- and through slabs sharing a device (two-row halo exchange)."""
    pb = replace(programs.Problem("x", 2, shape[0], shape[1], order=1, bc=capi.BC_REFLECT if bc == "reflect" else capi.BC_OUTFLOW,
                                  ic="implosion"), lx=shape[0] / shape[1])
    u0 = _random_state(pb, seed=shape[0] + shape[1])
    monkeypatch.setenv("SHLL_GRAPH", "0")
    outs = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("SHLL_FUSE2", fuse)
        for rpc in ("-", "5", "8", "1000"):
            if rpc == "-":
                monkeypatch.delenv("SHLL_ROWS_PER_CHUNK", raising=False)
            else:
                monkeypatch.setenv("SHLL_ROWS_PER_CHUNK", rpc)
            with programs.make_solver(pb, capi.MODE_FAST) as s:
                s.upload_u(u0)
                s.run(7); s.run(2); s.run(1); s.run(12)
                outs[(fuse, rpc)] = (s.download_u(), s.variant, s.launches)
    base = outs[("0", "-")]
    assert not base[1].endswith("_x2") and base[2] == 22
    for key, (u, name, launches) in outs.items():
        assert np.array_equal(bits(u), bits(base[0])), f"{key} {name} differs from one launch per step"
        if key[0] == "1":
            assert name.endswith("_x2") and launches == 4 + 1 + 1 + 6, (name, launches)
    # slabs on one device: halo_steps = 2 chosen by the group front end when every slab is at least 16 rows
    if shape[0] >= 48:
        monkeypatch.setenv("SHLL_FUSE2", "1")
        monkeypatch.delenv("SHLL_ROWS_PER_CHUNK", raising=False)
        _, _, _, dtdx, dtdy = programs.time_constants(pb)
        with capi.Group(2, pb.nx, pb.ny, ngpus=3, devices=[0, 0, 0], order=1, bc=pb.bc, mode=capi.MODE_FAST, dt_on_dx=float(dtdx),
                        dt_on_dy=float(dtdy)) as g:
            g.upload_u(u0)
            g.run(7); g.run(2); g.run(1); g.run(12)
            assert g.variant(0).endswith("_x2"), g.variant(0)
            assert np.array_equal(bits(g.download_u()), bits(base[0])), "3 slabs with two-step launches differ from one slab"


@pytest.mark.parametrize("n", [121, 1000, 4099, 30000, 200000])
@pytest.mark.parametrize("scheme", ["1d_o2_outflow", "1d_o2_reflect_mc"])
def test_fused_two_step_1d_launches_give_the_bits_of_single_steps(n, scheme, monkeypatch):
    """1D 2nd-order FAST: two steps per launch (csrc/step1d_acc.cuh, NSUB = 2) -- the second step runs on the first one's result
    in registers.  Same bits as one launch per step: walls (reflect) and outflow ends, MC limiter, ragged sizes, odd and split
    step counts; and through slabs sharing a device with exchange rounds of 2, 4 and 16 steps (a two-step launch never straddles
    a round; the send launch runs on a range extended by the first step's halo cells)."""
    monkeypatch.setenv("SHLL_PERSIST", "0")
    monkeypatch.setenv("SHLL_GRAPH", "0")
    pb = SCHEMES[scheme].resized(n)
    u0 = _random_state(pb, seed=n % 89)
    outs = {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("SHLL_FUSE1D", fuse)
        with programs.make_solver(pb, capi.MODE_FAST) as s:
            s.upload_u(u0)
            s.run(7); s.run(2); s.run(1); s.run(12)
            outs[fuse] = (s.download_u(), s.variant, s.launches)
    assert outs["0"][2] == 22 and outs["1"][2] == 4 + 1 + 1 + 6 and outs["1"][1].endswith("_x2")
    assert np.array_equal(bits(outs["0"][0]), bits(outs["1"][0])), f"N={n} {scheme}"
    if n >= 1000:
        _, _, _, dtdx, _ = programs.time_constants(pb)
        for K in ("2", "4", "16"):
            monkeypatch.setenv("SHLL_FUSE1D", "1")
            monkeypatch.setenv("SHLL_HALO_K", K)
            with capi.Group(1, pb.nx, 1, ngpus=3, devices=[0, 0, 0], order=2, bc=pb.bc, limiter=pb.limiter, tform=pb.tform, mode=capi.MODE_FAST,
                            alpha=pb.alpha, dt_on_dx=float(dtdx)) as g:
                g.upload_u(u0)
                g.run(7); g.run(2); g.run(1); g.run(12)
                assert g.variant(0).endswith("_x2")
                assert np.array_equal(bits(g.download_u()), bits(outs["0"][0])), f"N={n} {scheme}: 3 slabs, rounds of {K} steps"


# ------------------------------------------------------------------------------------ full-length FAST runs

@pytest.mark.parametrize("case", sorted(c for c in FAST_TOL_LONG if c != "long_1d_o2_65536"))   # (that one: test_gpu_parity.py)
def test_fast_full_length_runs_within_the_stated_long_run_tolerance(case):
    """The reference programs run to their own end time (820 / 1639 / 308 steps) in FAST mode, final primitive fields as
    shll_download_p returns them against the compiled reference's dump (committed fixture, tests/golden/make_golden_long.py).
    FAST_TOL_LONG is the table README.md and bench.py quote; it must stay within 5x of max(measured here, the reference's own
    sensitivity to FMA contraction on the same run) -- checked on the CPU side by test_oracle_golden.py."""
    import json
    import os
    from conftest import GOLDEN_DIR
    meta = json.load(open(os.path.join(GOLDEN_DIR, "MANIFEST_LONG.json")))[case]
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    pb = programs.PROGRAMS[meta["program"]].resized(meta["nx"], meta["ny"])
    r = programs.run_program(pb, capi.MODE_FAST)
    assert r["steps"] == meta["steps"] == int(z["steps"])
    st = int(z["stride"])
    got = r["p"].reshape(4, pb.nx, pb.ny)[:, ::st, ::st].astype(np.float64)
    ref = z["p"].astype(np.float64)
    assert np.isfinite(got).all()
    err = np.abs(got - ref)
    rel = float((err / (1.0 + np.abs(ref))).max())
    print(f"[fast long run] {case}: {r['steps']} steps, {r['variant']}: max |dp|/(1+|p|) = {rel:.3e}, max |dp| = {err.max():.3e}, "
          f"reference FMA sensitivity {float(z['ref_fma']):.3e}")
    assert rel <= FAST_TOL_LONG[case], f"{case} ({r['variant']}): {rel:.3e} > {FAST_TOL_LONG[case]:.1e}"


# ------------------------------------------------------------------------------------ a lost neighbour costs ONE timeout

def test_lost_neighbour_is_bounded_by_one_timeout(monkeypatch):
    """two connected slabs, only one of them ever runs.  Every step of the running slab used to wait the full
    timeout again (steps x timeout); the error word is sticky now, so the whole run costs one timeout."""
    import time
    monkeypatch.setenv("SHLL_HALO_TIMEOUT_MS", "300")
    pb = replace(programs.BASE_SHLL_2D.resized(128, 128), lx=1.0)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb)).reshape(4, 128, 128)
    a = programs.make_solver(pb, capi.MODE_FAST, rank=0, nranks=2, nx_local=64)
    b = programs.make_solver(pb, capi.MODE_FAST, rank=1, nranks=2, nx_local=64)
    try:
        da, db = a.peer_export(), b.peer_export()
        a.peer_connect(+1, db)
        b.peer_connect(-1, da)
        a.upload_u(np.ascontiguousarray(u0[:, :64]).reshape(4, -1))
        b.upload_u(np.ascontiguousarray(u0[:, 64:]).reshape(4, -1))
        t0 = time.perf_counter()
        with pytest.raises(capi.ShllError) as ei:
            a.run(60)          # slab b never steps: its flag for step 2 never arrives
            a.sync()
        dt = time.perf_counter() - t0
        assert ei.value.code == capi.E_TIMEOUT
        assert dt < 3.0, f"60 steps with a dead neighbour took {dt:.1f} s (one 0.3 s timeout expected)"
    finally:
        a.close()
        b.close()
