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
