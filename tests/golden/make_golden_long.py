#!/usr/bin/env python3
"""Long-run golden fixtures for the FAST-mode tolerance table (tests/conftest.py: FAST_TOL_LONG).

Full-length runs of the compiled reference programs in oracle/_ref (built by oracle/build_ref.py from /root/reference) --
too long to recompute inside the GPU test run, so their final PRIMITIVE fields (what Save_Results would dump) are
committed, the 1024^2 one as a stride-4 sample in both directions to stay small.  Alongside, the reference's own
sensitivity to FMA contraction on the same run (oracle restatement built -ffp-contract=fast vs the plain build), which is
the scale the FAST tolerance is anchored to.

Run in the build container:  python oracle/build_ref.py && python tests/golden/make_golden_long.py
"""
import json
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

# case -> (reference binary, program, nx, ny, sample stride)
CASES = {
    "long_2d_o2_256": ("ref_2d_o2_256", "2nd_order_base_shll", 256, 256, 1),
    "long_omp_o2_128": ("ref_omp_o2_128", "base_omp_2nd_order", 128, 128, 1),
    "long_2d_o1_1024": ("ref_2d_o1_1024", "base_shll_2d", 1024, 1024, 4),
}


def main():
    O.build_lib()
    out = {}
    for case, (exe, prog, nx, ny, stride) in CASES.items():
        r = O.run_ref(exe, 4, nx * ny, save=True, threads=4 if "omp" in exe else None)
        p = r["p"].reshape(4, nx, ny)
        pb = programs.PROGRAMS[prog].resized(nx, ny)
        assert r["steps"] == programs.count_steps(pb)
        # the reference's own FMA sensitivity on this run (restatement, both builds)
        u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
        cfg = oracle_cfg_for(O, pb, nthreads=8)
        a = O.run(cfg, u0, r["steps"])
        assert np.array_equal(a.view(np.uint32), r["u"].view(np.uint32)), "restatement != reference binary"
        b = O.run_fma_contracted(cfg, u0, r["steps"])
        pa, _ = programs.prim_from_cons(pb, a)
        pf, _ = programs.prim_from_cons(pb, b)
        sens = float((np.abs(pa.astype(np.float64) - pf) / (1.0 + np.abs(pa.astype(np.float64)))).max())
        np.savez_compressed(os.path.join(HERE, case + ".npz"), p=np.ascontiguousarray(p[:, ::stride, ::stride]),
                            steps=np.int64(r["steps"]), stride=np.int64(stride), ref_fma=np.float64(sens))
        out[case] = dict(ref_binary=exe, program=prog, nx=nx, ny=ny, steps=int(r["steps"]), stride=stride, ref_fma_sensitivity=sens)
        print(case, out[case], os.path.getsize(os.path.join(HERE, case + ".npz")), flush=True)
    with open(os.path.join(HERE, "MANIFEST_LONG.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
