#!/usr/bin/env python3
"""GPU sweep helper: time the 2D FAST kernels at the benched sizes under environment-variable variations.
usage: tools/sweep2d.py <o1|o2> "VAR=a,b,c" ["VAR2=x,y"] ...   (cartesian product; 30 warm-up + STEPS timed steps, best of 3)"""
import itertools
import os
import sys
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shll_sve_cfd_b200 import capi, programs  # noqa: E402


def main():
    which = sys.argv[1]
    axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:]]
    steps = int(os.environ.get("SWEEP_STEPS", "200"))
    if which == "o1":
        pb = programs.BASE_SHLL_2D.resized(4096, 4096)
    elif which.startswith("o1:") or which.startswith("o2:"):      # e.g. o1:1024 -> the reference's own sizes
        n = int(which.split(":")[1])
        pb = (programs.BASE_SHLL_2D if which[1] == "1" else programs.SECOND_ORDER_2D).resized(n, n)
    else:
        pb = replace(programs.SECOND_ORDER_2D.resized(2048, 16384), lx=2048 / 16384, ly=1.0)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    ref = None
    for combo in itertools.product(*[v for _, v in axes]):
        for (k, _), v in zip(axes, combo):
            if v == "-":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        try:
            with programs.make_solver(pb, capi.MODE_STRICT if os.environ.get("SWEEP_MODE") == "strict" else capi.MODE_FAST) as s:
                s.upload_u(u0)
                s.run(30)
                s.sync()
                ms = min(s.run_timed(steps) for _ in range(3)) / steps
                name = s.variant
                chk = float(s.conserved_sums()[0])
        except Exception as ex:
            print(dict(zip([k for k, _ in axes], combo)), "FAILED", str(ex)[:100], flush=True)
            continue
        if ref is None:
            ref = chk
        gcu = pb.ncells / (ms * 1e-3) / 1e9
        print(" ".join(f"{k}={v}" for (k, _), v in zip(axes, combo)), f"| {ms * 1e3:8.2f} us/step {gcu:7.1f} Gcu/s frac {gcu * 32 / 6554.9:.3f} | {name} | mass {'same' if chk == ref else 'DIFFERENT'}", flush=True)


if __name__ == "__main__":
    main()
