import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build_lib()
    return O


# FAST-mode tolerance of the CUDA path: |x - ref| <= tol + tol*|ref| on every primitive field.  3e-5 on the survey's parity
# configs (BASELINE.md section 2 / SURVEY.md App. A).  The scale behind it is the reference's OWN sensitivity to FMA
# contraction (the same C built -ffp-contract=fast, as on Graviton, vs the x86 build): <= 5e-6 on those configs, but
# 1.5e-4 on the non-square 2nd-order fixture this repo added (615 steps, density down to 0.12), so that one gets 6e-4.
# tests/test_oracle_golden.py::test_fast_tolerance_is_anchored_to_the_reference_sensitivity keeps the table honest:
# every entry must be >= 2x and <= max(3e-5, 5x) the measured sensitivity.
FAST_TOL_DEFAULT = 3e-5
FAST_TOL = {"2d_o2_96x160": 6e-4}
# Full-length runs of the reference programs in FAST mode (tests/test_gpu_fast_parity.py, fixtures made from the compiled
# reference by tests/golden/make_golden_long.py): rounding drift grows with the step count and with the limiter's branch
# flips, so long runs get their own entries, bounding max |p - p_ref| / (1 + |p_ref|) over the primitive fields.  Measured on
# B200 (round 2, profiles/r02_fast_long_runs.log) next to the reference's OWN sensitivity to FMA contraction on the same run
# (the same C built -ffp-contract=fast, as a Graviton build does): 3.7e-6 vs 4.8e-6 (1024^2, 820 steps), 3.05e-5 vs 3.1e-5
# (256^2 2nd order, 1639 steps), 7.3e-6 vs 5.8e-6 (base-omp 128^2, 308 steps) -- FAST mode moves the result as much as the
# compiler flag does.  Every tolerance is within 5x of the larger of the two.  README.md and bench.py's `arith_mode` quote
# this table.
FAST_TOL_LONG = {
    # case (tests/golden/<case>.npz, made by tests/golden/make_golden_long.py from the compiled reference): tolerance
    "long_2d_o1_1024": 3e-5,    # base_shll_2d.c, 1024^2, 820 steps to t = 0.1 (README Table 9 size)
    "long_2d_o2_256": 1e-4,     # 2nd_order_base_shll.c as checked in: 256^2, 1639 steps to t = 0.8
    "long_omp_o2_128": 3e-5,
    "long_1d_o2_65536": 1e-4,   # BASELINE.json configs[1] (65 536 cells, 104 858 steps): NOT a max-norm bound at this run length -- the level
                                # at which the share of cells is reported; the assertions are FAST_LONG_1D_* below    # base-omp/2nd_order_base_shll.c (MC limiter, configuration 6), 128^2, 308 steps to t = 0.3
}


# configs[1] run to completion in FAST mode (104 858 steps) against the full-length fixture.  Over 1e5 steps the different but
# equivalent FP32 formulas of FAST mode (reciprocal instead of divisions, FP32 temperature, face-flux form) drift apart from the
# reference at the 1e-4 level in the smooth parts, and a discontinuity that lands one cell to the side is an O(jump) pointwise
# difference -- the max norm says nothing.  Measured on B200 (profiles/r02_fast_long_runs.log): mean |dp|/(1+|p|) 6.9e-5, 33.5 % of
# the cells above 1e-4, 12.2 % above 3e-4, 0.13 % above 1e-3, 0.027 % above 1e-2, max 0.11.  The reference's OWN sensitivity to FMA
# contraction on this run (restatement built -ffp-contract=fast): mean 1.0e-5, 0.48 % above 1e-4, 0.04 % above 1e-3, max 0.064.
# So for runs of this length FAST is a throughput mode, not a substitute for the bit-exact STRICT mode (1.46 vs 0.99 us per step here).
FAST_LONG_1D_MEAN_TOL = 2e-4          # mean over cells and fields
FAST_LONG_1D_SHARE_ABOVE = {1e-3: 0.005, 1e-2: 0.001}   # error level -> largest share of cells that may exceed it


def load_golden(case: str):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    return z["u"], z["p"], int(z["steps"])


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def problem_from_manifest(entry: dict):
    """MANIFEST entry -> shll_sve_cfd_b200.programs.Problem"""
    from shll_sve_cfd_b200 import capi, programs
    bc = capi.BC_REFLECT if entry["bc"] == "reflect" else capi.BC_OUTFLOW
    lim = capi.LIM_MC if entry.get("limiter") == "mc" else capi.LIM_MINMOD
    dims = entry["dims"]
    tform = capi.TFORM_AUTO
    if dims == 1 and entry["order"] == 2:
        tform = capi.TFORM_2D
    return programs.Problem(
        name=entry["ref_binary"], dims=dims, nx=entry["nx"], ny=(entry["ny"] if dims == 2 else 1), order=entry["order"],
        bc=bc, limiter=lim, alpha=entry.get("alpha", 1.25), ic=entry["ic"], total_time=entry["total"], tform=tform)


def oracle_cfg_for(O, pb, nthreads=1):
    from shll_sve_cfd_b200 import capi, programs
    _, _, _, dtdx, dtdy = programs.time_constants(pb)
    tform = pb.tform
    if tform == capi.TFORM_AUTO:
        tform = capi.TFORM_1D if (pb.dims == 1 and pb.order == 1) else capi.TFORM_2D
    return O.make_cfg(pb.dims, pb.nx, pb.ny, order=pb.order, bc=pb.bc, limiter=pb.limiter, tform=tform, alpha=pb.alpha,
                      dt_on_dx=float(dtdx), dt_on_dy=float(dtdy), nthreads=nthreads)


ORACLE_IC = {"sod_1d": 0, "implosion": 1, "four_shock": 2, "config6": 3, "sod_x": 4}
