#!/usr/bin/env python3
"""Bar charts of the reference's published timings with the B200 bars next to them -- the counterpart of the reference's
graphics/generate_figures.py (which hard-codes the README's timing tables and draws speed-up bars, :13-53).

Published numbers (cited, not re-measured: Graviton4, and the SVE sources do not build on x86) are the averages of runs A/B of
README.md Tables 2, 4 (1D, GCC -O3: base C, SVE intrinsics, :86-133), 10, 7 (2D 1st order, :186-196, :170-178), 14, 12 (2D 2nd order,
:236-244, :218-226) and the 16-thread rows of Tables 16 / 17 (OpenMP / OpenMP + SVE at 1024^2, :256-281).  B200 numbers come from
profiles/r02_readme_tables.json (tools/readme_tables.py: the same runs through the C ABI, upload and download included).

Writes figures/*.svg (no matplotlib needed: the image has none; with matplotlib installed, --show also opens the figures) and
figures/speedups.md.    usage: python tools/generate_figures.py [tables.json] [outdir]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PUBLISHED = {   # seconds, (run A + run B) / 2
    "1d": {"sizes": [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536],
           "base C -O3 (Table 2)": [0.002, 0.007, 0.025, 0.099, 0.4035, 1.596, 6.316, 29.4805, 121.825],
           "SVE -O3 (Table 4)": [0.001, 0.004, 0.014, 0.057, 0.2265, 0.892, 3.483, 17.826, 71.2685]},
    "2d_o1": {"sizes": [256, 512, 1024, 2048],
              "base C -O3 (Table 10)": [1.195, 14.9315, 79.378, 645.4915],
              "SVE -O3 (Table 7)": [0.546, 5.976, 40.058, 315.8025]},
    "2d_o2": {"sizes": [256, 512, 1024],
              "base C -O3 (Table 14)": [26.3675, 315.1165, 1941.9065],
              "SVE -O3 (Table 12)": [16.0195, 195.8205, 1321.828]},
    "omp_1024": {"sizes": [1024], "OpenMP 16 threads (Table 16)": [47.39], "OpenMP + SVE 16 threads (Table 17)": [33.959]},
}
PROGRAM_OF = {"1d": "base_shll", "2d_o1": "base_shll_2d", "2d_o2": "2nd_order_base_shll", "omp_1024": "base_omp_2nd_order"}


def svg_bars(path, title, labels, series, ylabel):
    """Grouped log-scale bar chart as plain SVG.  series: list of (name, values)."""
    import math
    W, H, L, B, T = 980, 520, 90, 70, 50
    vals = [v for _, vs in series for v in vs if v and v > 0]
    lo, hi = math.floor(math.log10(min(vals))), math.ceil(math.log10(max(vals)))
    y = lambda v: T + (H - T - B) * (1 - (math.log10(v) - lo) / (hi - lo))
    colors = ["#7a7a7a", "#4c78a8", "#76b900", "#b5e550"]
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="13">',
           f'<rect width="{W}" height="{H}" fill="white"/><text x="{W / 2}" y="24" text-anchor="middle" font-size="16">{title}</text>']
    for e in range(lo, hi + 1):
        out.append(f'<line x1="{L}" x2="{W - 20}" y1="{y(10 ** e):.1f}" y2="{y(10 ** e):.1f}" stroke="#ddd"/>'
                   f'<text x="{L - 8}" y="{y(10 ** e) + 4:.1f}" text-anchor="end">1e{e}</text>')
    gw = (W - L - 20) / len(labels)
    bw = gw * 0.8 / len(series)
    for i, lab in enumerate(labels):
        out.append(f'<text x="{L + gw * (i + 0.5):.1f}" y="{H - B + 18}" text-anchor="middle">{lab}</text>')
        for k, (_, vs) in enumerate(series):
            if vs[i] and vs[i] > 0:
                x = L + gw * i + gw * 0.1 + bw * k
                out.append(f'<rect x="{x:.1f}" y="{y(vs[i]):.1f}" width="{bw * 0.92:.1f}" height="{H - B - y(vs[i]):.1f}" fill="{colors[k % 4]}"/>')
    for k, (name, _) in enumerate(series):
        out.append(f'<rect x="{L + 10 + 235 * k}" y="{H - 28}" width="12" height="12" fill="{colors[k % 4]}"/>'
                   f'<text x="{L + 27 + 235 * k}" y="{H - 17}">{name}</text>')
    out.append(f'<text x="18" y="{H / 2}" transform="rotate(-90 18 {H / 2})" text-anchor="middle">{ylabel}</text></svg>')
    open(path, "w").write("\n".join(out))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tables = args[0] if args else os.path.join(ROOT, "profiles", "r02_readme_tables.json")
    outdir = args[1] if len(args) > 1 else os.path.join(ROOT, "figures")
    os.makedirs(outdir, exist_ok=True)
    rows = json.load(open(tables))["rows"]
    b200 = {(r["program"], r["nx"]): r for r in rows}
    md = ["| problem | cells | steps | published (Graviton4, README) | B200 strict (bit-exact) | B200 fast | speed-up, fast vs best published |", "|---|---|---|---|---|---|---|"]
    for key, pub in PUBLISHED.items():
        sizes = pub["sizes"]
        series = [(name, vals) for name, vals in pub.items() if name != "sizes"]
        strict = [b200.get((PROGRAM_OF[key], n), {}).get("strict_s") for n in sizes]
        fast = [b200.get((PROGRAM_OF[key], n), {}).get("fast_s") for n in sizes]
        series += [("B200 strict (bit-exact)", strict), ("B200 fast", fast)]
        labels = [str(n) if key == "1d" else f"{n}x{n}" for n in sizes]
        svg_bars(os.path.join(outdir, f"timings_{key}.svg"), f"Whole-program run time, {PROGRAM_OF[key]} (log scale)", labels, series, "seconds")
        for i, n in enumerate(sizes):
            best_name, best = min(((nm, vs[i]) for nm, vs in series[:-2]), key=lambda t: t[1])
            r = b200.get((PROGRAM_OF[key], n))
            if r:
                md.append(f"| {PROGRAM_OF[key]} | {labels[i]} | {r['steps']} | {best:g} s ({best_name}) | {r['strict_s'] * 1e3:.2f} ms | {r['fast_s'] * 1e3:.2f} ms | {best / r['fast_s']:.0f}x |")
    open(os.path.join(outdir, "speedups.md"), "w").write(
        "B200 run times include the upload of the initial state and the download of the primitives (tools/readme_tables.py).\n"
        "README's 2D 1st-order step counts (205 at 256^2 ...) correspond to t = 0.1, the value checked into base_shll_2d.c; the B200 runs use the same.\n"
        "README's 256^2 2nd-order entry '1649 steps' is a typo for 1639 (SURVEY.md section 8c).  The OpenMP rows are the 16-thread timings of Tables 16 / 17.\n\n"
        + "\n".join(md) + "\n")
    print("\n".join(md))
    if "--show" in sys.argv:
        try:
            import matplotlib.pyplot as plt  # noqa: F401
            print("matplotlib present: open the SVGs under", outdir)
        except ImportError:
            print("matplotlib is not installed; SVGs written to", outdir)


if __name__ == "__main__":
    main()
