"""GPU diagnostic: how far do the FAST kernels (window form / face-flux accumulate form) drift from the STRICT oracle,
as a function of the step count?  A bug shows up at step 1 at a definite place; rounding drift grows smoothly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import oracle_cfg_for  # noqa: E402
from oracle import oracle as O  # noqa: E402
from shll_sve_cfd_b200 import capi, programs  # noqa: E402


def run(pb, u0, steps, mode, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with programs.make_solver(pb, mode) as s:
            s.upload_u(u0)
            s.run(steps)
            return s.download_u(), s.variant
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    cases = {
        "2d_o2_outflow": programs.SECOND_ORDER_2D,
        "2d_o2_reflect": programs.Problem("x", 2, 64, 64, order=2, bc=capi.BC_REFLECT, ic="four_shock"),
        "2d_o2_outflow_mc": programs.BASE_OMP_2D,
        "2d_o1_reflect": programs.BASE_SHLL_2D.resized(64, 64),
    }
    for name, pb in cases.items():
        u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
        total = programs.count_steps(pb)
        print(f"== {name}: {pb.nx}x{pb.ny}, {total} steps to t_end")
        for steps in (1, 2, 4, 16, 64, 256, total):
            ref = O.run(oracle_cfg_for(O, pb, nthreads=4), u0, steps).astype(np.float64)
            row = [f"steps={steps:5d}"]
            for label, env in (("win1", {"SHLL_ACC": "0", "SHLL_VEC": "1"}), ("win2", {"SHLL_ACC": "0", "SHLL_VEC": "2"}),
                               ("acc", {"SHLL_ACC": "1", "SHLL_VEC": "2"})):
                got, variant = run(pb, u0, steps, capi.MODE_FAST, env)
                err = np.abs(got.astype(np.float64) - ref)
                k, c = np.unravel_index(np.argmax(err), err.shape)
                row.append(f"{label}: {err.max():.2e} @comp{k} ({c // pb.ny},{c % pb.ny})")
            print("   ".join(row), flush=True)


if __name__ == "__main__":
    main()
