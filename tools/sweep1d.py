#!/usr/bin/env python3
"""GPU sweep helper: 1D FAST/STRICT 2nd-order streaming kernel at a given size under environment-variable variations.
usage: tools/sweep1d.py <ncells> <fast|strict> "VAR=a,b,c" ...  (200 timed steps, best of 3)"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shll_sve_cfd_b200 import capi, programs  # noqa: E402


def main():
    n = int(sys.argv[1])
    mode = capi.MODE_FAST if sys.argv[2] == "fast" else capi.MODE_STRICT
    axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[3:]]
    os.environ["SHLL_PERSIST"] = "0"
    os.environ["SHLL_GRAPH"] = "0"
    pb = programs.SECOND_ORDER_1D.resized(n)
    u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
    for combo in itertools.product(*[v for _, v in axes]):
        for (k, _), v in zip(axes, combo):
            if v == "-":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        with programs.make_solver(pb, mode) as s:
            s.upload_u(u0)
            s.run(30)
            s.sync()
            ms = min(s.run_timed(200) for _ in range(3)) / 200
        gcu = n / (ms * 1e-3) / 1e9
        print(" ".join(f"{k}={v}" for (k, _), v in zip(axes, combo)), f"| {ms * 1e3:8.2f} us/step {gcu:7.1f} Gcu/s frac {gcu * 24 / 6554.9:.3f}", flush=True)


if __name__ == "__main__":
    main()
