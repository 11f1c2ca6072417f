"""Single-process multi-GPU front end (shll_group_*, include/shll_b200.h): one handle, one slab per GPU.

CPU: the group entry points validate their arguments and fail loudly without a GPU.
GPU: the path is exercised on ONE GPU by listing device 0 for every slab (two or three slabs then share the device and still
exchange their halo rows through peer stores and flags), and on real device lists when the box has several GPUs.  The
result must be bitwise equal to the single-slab result in both arithmetic modes, and to the reference fixtures in STRICT
mode; the C host programs must write the same results.dat whatever SHLL_NGPUS is."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits, has_gpu, load_golden, problem_from_manifest

HOST = os.path.join(ROOT, "host")


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _group(pb, mode, ngpus, devices):
    from shll_sve_cfd_b200 import capi, programs
    _, _, _, dtdx, dtdy = programs.time_constants(pb)
    return capi.Group(pb.dims, pb.nx, pb.ny, ngpus=ngpus, devices=devices, order=pb.order, bc=pb.bc, limiter=pb.limiter,
                      tform=pb.tform, mode=mode, alpha=pb.alpha, dt_on_dx=float(dtdx), dt_on_dy=float(dtdy))


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_group_create_validates_arguments():
    from shll_sve_cfd_b200 import capi
    L = capi.lib()
    cfg = capi.Config()
    cfg.struct_size = C.sizeof(capi.Config)
    cfg.dims, cfg.nx, cfg.ny, cfg.order, cfg.dt_on_dx, cfg.dt_on_dy = 2, 64, 64, 1, 0.125, 0.125
    h = C.c_void_p(None)
    assert L.shll_group_create(C.byref(h), C.byref(cfg), 0, None) == capi.E_INVAL
    assert b"ngpus" in L.shll_group_last_error(None)
    assert L.shll_group_create(C.byref(h), C.byref(cfg), 100, None) == capi.E_INVAL and not h.value   # more slabs than rows
    assert L.shll_group_create(None, C.byref(cfg), 1, None) == capi.E_INVAL
    cfg.struct_size = 8
    assert L.shll_group_create(C.byref(h), C.byref(cfg), 1, None) == capi.E_INVAL
    assert L.shll_group_size(None) == 0 and L.shll_group_ctx(None, 0) is None and L.shll_group_destroy(None) == capi.OK
    assert L.shll_group_run(None, 1) == capi.E_INVAL


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_group_has_no_cpu_fallback():
    from shll_sve_cfd_b200 import capi
    with pytest.raises(capi.ShllError) as ei:
        capi.Group(2, 64, 64, ngpus=1)
    assert ei.value.code == capi.E_CUDA and "no CPU fallback" in str(ei.value)
    pr = subprocess.run([os.path.join(HOST, "base_shll"), "64"], capture_output=True, text=True, env=dict(os.environ, SHLL_NGPUS="2"))
    assert pr.returncode != 0 and "no CPU fallback" in pr.stderr and "Completed" not in pr.stdout


# ---------------------------------------------------------------------------------------------------------------- GPU
CASES = ["1d_o1_1024", "1d_o2_slice_1024", "2d_o1_64", "2d_o1_96x160", "2d_o2_64", "2d_o2_96x160", "omp_o2_64"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0]], ids=["2slabs", "3slabs"])
def test_group_on_one_device_is_bitwise_the_single_slab_run(case, devices, manifest):
    from shll_sve_cfd_b200 import capi, programs
    pb = problem_from_manifest(manifest[case])
    gu, gp, steps = load_golden(case)
    steps = min(steps, 300)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    for mode in (capi.MODE_STRICT, capi.MODE_FAST):
        with _group(pb, mode, 1, None) as g1:
            g1.upload_u(u0)
            g1.run(steps)
            want_u, (want_p, want_a) = g1.download_u(), g1.download_p(want_a=True)
            want_sums, want_cfl = g1.conserved_sums(), g1.max_cfl()
        with _group(pb, mode, len(devices), devices) as g:
            assert g.size == len(devices)
            g.upload_u(u0)
            g.run(steps // 2)
            g.run(steps - steps // 2)       # a second call continues from the slabs' own halos
            got_u, (got_p, got_a) = g.download_u(), g.download_p(want_a=True)
            assert np.array_equal(bits(got_u), bits(want_u)), f"{case} mode {mode}: {len(devices)} slabs differ from one"
            assert np.array_equal(bits(got_p), bits(want_p)) and np.array_equal(bits(got_a), bits(want_a))
            assert g.max_cfl() == want_cfl
            np.testing.assert_allclose(g.conserved_sums(), want_sums, rtol=1e-12, atol=1e-9)
            assert g.launches >= steps * len(devices) // 2 or pb.dims == 1    # (two steps per launch in 2D 1st-order FAST)
        if mode == capi.MODE_STRICT and steps == load_golden(case)[2]:
            assert np.array_equal(bits(got_u), bits(gu)), f"{case}: differs from the reference fixture"


@pytest.mark.gpu
def test_group_reupload_and_timed_run(manifest):
    from shll_sve_cfd_b200 import capi, programs
    pb = problem_from_manifest(manifest["2d_o1_64"])
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    with _group(pb, capi.MODE_STRICT, 2, [0, 0]) as g:
        g.upload_u(u0)
        g.run(50)
        first = g.download_u()
        g.upload_u(u0)                      # a new state resets every slab's halos
        assert g.run_timed(50) > 0.0
        assert np.array_equal(bits(g.download_u()), bits(first))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_group_on_real_devices(world, manifest):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from shll_sve_cfd_b200 import capi, programs
    for case in ("2d_o1_96x160", "2d_o2_96x160", "1d_o2_slice_1024"):
        pb = problem_from_manifest(manifest[case])
        gu, _, steps = load_golden(case)
        u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
        for mode in (capi.MODE_STRICT, capi.MODE_FAST):
            with _group(pb, mode, 1, None) as g1:
                g1.upload_u(u0); g1.run(steps); want = g1.download_u()
            with _group(pb, mode, world, None) as g:
                g.upload_u(u0); g.run(steps); got = g.download_u()
            assert np.array_equal(bits(got), bits(want)), f"{case} mode {mode}: {world} GPUs differ from one"
            if mode == capi.MODE_STRICT:
                assert np.array_equal(bits(got), bits(gu))


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.gpu
def test_host_programs_write_the_same_results_with_slabs(manifest, tmp_path):
    def run(exe, args, env):
        pr = subprocess.run([os.path.join(HOST, exe), *map(str, args)], cwd=tmp_path, env=dict(os.environ, **env), capture_output=True,
                            text=True, timeout=600)
        assert pr.returncode == 0, pr.stderr
        return pr.stdout
    two = {"SHLL_DEVICES": "0,0"} if _ngpus() < 2 else {"SHLL_NGPUS": "2"}
    assert run("base_shll", [1024], two) == "Completed in 1639 steps\n"          # BASELINE.json configs[0] on two slabs
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["1d_o1_1024"]["results_dat_md5"]
    out = run("base_shll_2d", [256], dict(two, SHLL_SAVE="1"))
    assert out == "Completed in 205 steps\nSaving to file\nCompleted saving data\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o1_256"]["results_dat_md5"]
    run("2nd_order_base_shll", [96, 160], {"SHLL_DEVICES": "0,0,0", "SHLL_SAVE": "1"})
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o2_96x160"]["results_dat_md5"]
    pr = subprocess.run([os.path.join(HOST, "base_shll"), "64"], cwd=tmp_path, env=dict(os.environ, SHLL_NGPUS="3", SHLL_DEVICES="0,0"),
                        capture_output=True, text=True)
    assert pr.returncode != 0 and "SHLL_DEVICES" in pr.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 2, 5, 16])
@pytest.mark.parametrize("order", [1, 2])
def test_1d_temporal_halo_blocking_any_round_length_is_bitwise_the_single_slab_run(K, order, monkeypatch):
    """1D slabs exchange K*order halo cells once every K steps (csrc/step1d.cuh, include/shll_b200.h: halo_steps).  Any K, any
    split of the step count over shll_run calls (rounds straddle calls), a re-upload in the middle of a round, ragged slab
    sizes, 2 / 3 / 5 slabs: the bits of one slab.  Random state, both arithmetic modes, streaming kernels (N above the tile)."""
    import zlib
    from shll_sve_cfd_b200 import capi, programs
    from test_gpu_parity import _random_state
    monkeypatch.setenv("SHLL_HALO_K", str(K))
    monkeypatch.setenv("SHLL_PERSIST", "0")          # the single-slab comparison run uses the streaming kernels too
    monkeypatch.setenv("SHLL_GRAPH", "0")
    base = programs.BASE_SHLL if order == 1 else programs.SECOND_ORDER_1D
    for n, nslabs in ((1003, 2), (4099, 3), (777, 5)):
        pb = base.resized(n)
        u0 = _random_state(pb, seed=zlib.crc32(f"{n}-{K}-{order}".encode()) % 1000)
        for mode in (capi.MODE_STRICT, capi.MODE_FAST):
            with _group(pb, mode, 1, None) as g1:
                g1.upload_u(u0); g1.run(7); g1.run(30); mid = g1.download_u()
                g1.upload_u(mid); g1.run(21); want = g1.download_u()
            with _group(pb, mode, nslabs, [0] * nslabs) as g:
                g.upload_u(u0)
                g.run(7); g.run(30)                 # 37 steps: not a multiple of any K > 1
                got_mid = g.download_u()
                assert np.array_equal(bits(got_mid), bits(mid)), f"N={n} slabs={nslabs} K={K} order={order} mode={mode}: mid state"
                g.upload_u(got_mid)                 # new state in the middle of a round: rounds restart here
                g.run(1); g.run(20)
                got = g.download_u()
                stats = capi.lib().shll_halo_wait_stats
            assert np.array_equal(bits(got), bits(want)), f"N={n} slabs={nslabs} K={K} order={order} mode={mode}"


@pytest.mark.gpu
def test_halo_steps_must_fit_the_slab_and_the_mailbox():
    from shll_sve_cfd_b200 import capi
    for kw in (dict(nx=40, order=2, halo_steps=17), dict(nx=20, order=2, halo_steps=16), dict(nx=10, order=1, halo_steps=11)):
        with pytest.raises(capi.ShllError) as ei:
            capi.Solver(1, kw["nx"], order=kw["order"], rank=0, nranks=2, halo_steps=kw["halo_steps"])
        assert ei.value.code == capi.E_INVAL and "halo_steps" in str(ei.value)
    capi.Solver(1, 64, order=2, rank=0, nranks=2, halo_steps=16).close()
    capi.Solver(1, 64, order=2, rank=0, nranks=1, halo_steps=99).close()    # a single slab ignores it
