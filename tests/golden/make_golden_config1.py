#!/usr/bin/env python3
"""Golden fixture for BASELINE.json configs[1] run to completion: the derived 1D 2nd-order program (x-sweep of
base-c/2nd_order_base_shll.c, minmod, outflow, Sod tube) at 65 536 cells, 104 858 steps to t = 0.2.

The compiled reference for this program is the y-uniform NY = 4 run of the 2D file (oracle/_ref/ref_1d_o2_slice_65536),
2.7e10 cell-updates = hours on one core; the fixture is therefore produced by the oracle restatement (oracle/shll_oracle.c,
4 threads, ~4 minutes), which is pinned to that compiled reference bit for bit on the same program at 1024 / 4096 cells run
to completion and at 65 536 cells on a step-capped prefix (tests/test_oracle_golden.py, tests/test_oracle_vs_ref.py) -- and
this script re-checks the 65 536-cell prefix against the reference binary before writing anything.

Run in the build container:  python oracle/build_ref.py && python tests/golden/make_golden_config1.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import oracle_cfg_for  # noqa: E402
from oracle import oracle as O  # noqa: E402
from shll_sve_cfd_b200 import programs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build_lib()
    n = 65536
    pb = programs.SECOND_ORDER_1D.resized(n)
    steps = programs.count_steps(pb)
    assert steps == 104858
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    cfg = oracle_cfg_for(O, pb, nthreads=4)
    # prefix against the compiled reference (NY = 4 y-uniform run of the 2D program)
    cap = 300
    r = O.run_ref("ref_1d_o2_slice_65536", 4, n * 4, step_cap=cap, raw=True)
    ref = r["u"].reshape(4, n, 4)[[0, 1, 3], :, 0]
    got = O.run(cfg, u0, cap)
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32)), "restatement != compiled reference on the prefix"
    u = O.run(cfg, u0, steps)
    p, _ = programs.prim_from_cons(pb, u)
    np.savez_compressed(os.path.join(HERE, "long_1d_o2_65536.npz"), u=u, p=p, steps=np.int64(steps), prefix_steps_checked_vs_reference=np.int64(cap))
    print("long_1d_o2_65536", steps, os.path.getsize(os.path.join(HERE, "long_1d_o2_65536.npz")))


if __name__ == "__main__":
    main()
