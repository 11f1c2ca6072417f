#!/usr/bin/env python3
"""B200 timings of the runs the reference's README tabulates (README.md Tables 1-17: whole-program run times per grid size),
through the C ABI exactly as the C host programs use it: upload + all time steps to the program's end time + download
of the primitives, wall clock.  Writes profiles/r02_readme_tables.json, which tools/generate_figures.py draws next to the
README's published Graviton4 numbers (cited there, not re-measured: the SVE variants cannot be built on x86).

usage (GPU box): python tools/readme_tables.py [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shll_sve_cfd_b200 import capi, programs  # noqa: E402

CASES = (
    [("base_shll", n, 1) for n in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536)] +          # README Tables 1-6
    [("base_shll_2d", n, n) for n in (256, 512, 1024, 2048)] +                                          # Tables 7-10
    [("2nd_order_base_shll", n, n) for n in (256, 512, 1024)] +                                         # Tables 11-15
    [("base_omp_2nd_order", n, n) for n in (512, 1024)])                                               # Tables 16-17 (t = 0.3, MC limiter)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_readme_tables.json")
    rows = []
    for prog, nx, ny in CASES:
        pb = programs.PROGRAMS[prog].resized(nx, ny) if ny > 1 else programs.PROGRAMS[prog].resized(nx)
        steps = programs.count_steps(pb)
        u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
        row = {"program": prog, "nx": nx, "ny": ny, "steps": steps}
        for mode_name, mode in (("strict", capi.MODE_STRICT), ("fast", capi.MODE_FAST)):
            with programs.make_solver(pb, mode) as s:
                s.upload_u(u0); s.run(min(steps, 64)); s.sync()        # warm-up (graph capture, first launches)
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    s.upload_u(u0)
                    s.run(steps)
                    s.download_p()
                    best = min(best, time.perf_counter() - t0)
                row[mode_name + "_s"] = best
                row[mode_name + "_kernel"] = s.variant
        rows.append(row)
        print(f"{prog:22s} {nx:6d} x {ny:5d} {steps:7d} steps: strict {row['strict_s'] * 1e3:9.3f} ms  fast {row['fast_s'] * 1e3:9.3f} ms", flush=True)
    json.dump({"what": "whole-program run time on one B200 through the C ABI (upload + all steps + download of the primitives), best of 3",
               "rows": rows}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
