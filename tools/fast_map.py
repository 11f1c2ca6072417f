"""GPU diagnostic: per-region error of the FAST kernels after a few steps from a smooth random state (in ulps of the field scale)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import oracle_cfg_for  # noqa: E402
from oracle import oracle as O  # noqa: E402
from shll_sve_cfd_b200 import capi, programs  # noqa: E402
from test_gpu_parity import _random_state  # noqa: E402
from fast_drift import run  # noqa: E402


def main():
    nx, ny = 130, 128
    for name, pb0 in (("o2_outflow", programs.SECOND_ORDER_2D), ("o2_mc", programs.BASE_OMP_2D),
                      ("o2_reflect", programs.Problem("x", 2, 64, 64, order=2, bc=capi.BC_REFLECT, ic="four_shock"))):
        pb = pb0.resized(nx, ny)
        u0 = _random_state(pb, seed=3)
        for steps in (1, 3, 20):
            ref = O.run(oracle_cfg_for(O, pb, nthreads=4), u0, steps).astype(np.float64).reshape(4, nx, ny)
            for label, env in (("win2", {"SHLL_ACC": "0", "SHLL_VEC": "2", "SHLL_ROWS_PER_CHUNK": "24"}),
                               ("acc ", {"SHLL_ACC": "1", "SHLL_VEC": "2", "SHLL_ROWS_PER_CHUNK": "24"})):
                got, variant = run(pb, u0, steps, capi.MODE_FAST, env)
                err = np.abs(got.astype(np.float64).reshape(4, nx, ny) - ref).max(axis=0) / 1.2e-7
                regs = {
                    "rows0-1": err[0:2], "rows2-3": err[2:4], "rows-2..": err[-2:], "rows-4..-3": err[-4:-2],
                    "cols0-1": err[4:-4, 0:2], "cols-2..": err[4:-4, -2:], "cols58-61": err[4:-4, 58:62], "cols118-121": err[4:-4, 118:122],
                    "chunk rows 20-27": err[20:28, 4:-4], "interior": err[30:100, 4:56],
                }
                print(f"{name} steps={steps:2d} {label}: " + "  ".join(f"{k}={v.max():.1f}/{v.mean():.2f}" for k, v in regs.items()), flush=True)


if __name__ == "__main__":
    main()
