"""GPU (-m gpu): the C host programs in host/ are drop-ins for the reference programs: same stdout, same results.dat
(md5 of the text file the unmodified reference writes, SURVEY.md App. B / tests/golden/MANIFEST.json)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bits, load_golden

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "host")


def _run(exe, args, tmp_path, env=None):
    path = os.path.join(HOST, exe)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", HOST])
    e = dict(os.environ)
    e.update(env or {})
    pr = subprocess.run([path, *map(str, args)], cwd=tmp_path, env=e, capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, pr.stderr
    return pr.stdout


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def test_base_shll_as_shipped_and_config0(manifest, tmp_path):
    out = _run("base_shll", [], tmp_path)
    assert out == "Completed in 410 steps\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["1d_o1_256"]["results_dat_md5"]
    out = _run("base_shll", [1024], tmp_path)          # BASELINE.json configs[0]
    assert out == "Completed in 1639 steps\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["1d_o1_1024"]["results_dat_md5"]


def test_base_shll_2d(manifest, tmp_path):
    out = _run("base_shll_2d", [], tmp_path)
    assert out == "Completed in 205 steps\n"            # Save_Results is commented out in the reference
    assert not os.path.exists(os.path.join(tmp_path, "results.dat"))
    out = _run("base_shll_2d", [256], tmp_path, env={"SHLL_SAVE": "1"})
    assert out == "Completed in 205 steps\nSaving to file\nCompleted saving data\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o1_256"]["results_dat_md5"]
    _run("base_shll_2d", [96, 160], tmp_path, env={"SHLL_SAVE": "1"})
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o1_96x160"]["results_dat_md5"]


def test_second_order_2d(manifest, tmp_path):
    out = _run("2nd_order_base_shll", [64], tmp_path, env={"SHLL_SAVE": "1"})
    assert out.startswith("Completed in 410 steps\n")
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o2_64"]["results_dat_md5"]
    _run("2nd_order_base_shll", [96, 160], tmp_path, env={"SHLL_SAVE": "1"})
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o2_96x160"]["results_dat_md5"]


def test_derived_second_order_1d(manifest, tmp_path):
    out = _run("2nd_order_base_shll_1d", [1024], tmp_path)
    assert out == "Completed in 1639 steps\n"
    _, gp, _ = load_golden("1d_o2_slice_1024")
    got = np.loadtxt(os.path.join(tmp_path, "results.dat"))
    # results.dat holds %e text (7 significant digits): compare against the golden primitives formatted the same way
    want = np.array([[float("%e" % v) for v in row] for row in gp]).T
    assert np.array_equal(got[:, 1:], want)


def test_fixed_step_count_for_grids_whose_float_clock_stalls(tmp_path):
    pr = subprocess.run([os.path.join(HOST, "2nd_order_base_shll_1d"), str(1 << 26)], cwd=tmp_path, capture_output=True, text=True,
                        env=dict(os.environ, SHLL_SAVE="0"), timeout=600)
    assert pr.returncode != 0 and "stalls" in pr.stderr     # the reference would loop forever here
    out = _run("2nd_order_base_shll_1d", [1 << 26], tmp_path, env={"SHLL_STEPS": "20", "SHLL_SAVE": "0"})
    assert out == "Completed in 20 steps\n"


def _read_bin(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"SHLLBIN1"
    dims, nx, ny, ncomp, steps = np.frombuffer(raw[8:28], dtype="<i4")
    planes = np.frombuffer(raw[64:], dtype="<f4").reshape(ncomp, nx * ny)
    return dict(dims=int(dims), nx=int(nx), ny=int(ny), steps=int(steps)), planes


def test_binary_dump_snapshots_and_monitor(manifest, tmp_path):
    """binary dump, periodic snapshots, CFL / conservation monitor -- none of them may change
    stdout, results.dat or the result itself."""
    env = {"SHLL_SAVE": "1", "SHLL_SAVE_BIN": "1", "SHLL_SNAPSHOT_EVERY": "100", "SHLL_MONITOR": "1"}
    path = os.path.join(HOST, "base_shll_2d")
    pr = subprocess.run([path, "256"], cwd=tmp_path, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert pr.returncode == 0, pr.stderr
    assert pr.stdout == "Completed in 205 steps\nSaving to file\nCompleted saving data\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["2d_o1_256"]["results_dat_md5"]
    _, gp, _ = load_golden("2d_o1_256")
    hdr, planes = _read_bin(os.path.join(tmp_path, "results.bin"))
    assert hdr == dict(dims=2, nx=256, ny=256, steps=205)
    assert np.array_equal(bits(planes), bits(gp))                      # raw float32: the exact primitives, not 7 digits
    snaps = sorted(f for f in os.listdir(tmp_path) if f.startswith("snapshot_"))
    assert snaps == ["snapshot_00000100.bin", "snapshot_00000200.bin"]
    assert _read_bin(os.path.join(tmp_path, snaps[1]))[0]["steps"] == 200
    mon = [l for l in pr.stderr.splitlines() if l.startswith("monitor step")]
    assert [int(l.split()[2].rstrip(":")) for l in mon] == [0, 100, 200, 205]
    mass = [float(l.split("mass")[1].split()[0]) for l in mon]
    assert max(mass) - min(mass) < 1e-6 * mass[0]                       # reflective walls conserve mass
    cfl = [float(l.split("max CFL")[1].split()[0]) for l in mon]
    assert all(0.1 < c < 0.5 for c in cfl)


def test_base_omp_second_order_mc_limiter(manifest, tmp_path):
    """-DPROGRAM=5: drop-in for base-omp/2nd_order_base_shll.c (MC limiter alpha = 1.25, configuration-6 IC, t = 0.3;
    `Completed in %d steps` once per OpenMP thread, 16 threads hard-coded in the reference)."""
    out = _run("base_omp_2nd_order", [64], tmp_path, env={"SHLL_SAVE": "1"})
    assert out == "Completed in 154 steps\n" * 16 + "Saving to file\nCompleted saving data\n"
    assert _md5(os.path.join(tmp_path, "results.dat")) == manifest["omp_o2_64"]["results_dat_md5"]
    out = _run("base_omp_2nd_order", [128], tmp_path, env={"SHLL_OMP_LINES": "1"})
    assert out == "Completed in 308 steps\n"          # Save_Results is commented out in the reference's main()
