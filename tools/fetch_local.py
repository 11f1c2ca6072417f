#!/usr/bin/env python3
"""Look at the output of a run: the counterpart of the reference's base-c/fetch_local.py.

Reads what the host programs (host/shll_main.c) write --
  results.dat               the reference's text format: 1D `x rho u T`, 2D `x y rho ux uy T`, one cell per line, i-major
                            (base_shll.c:180-192, base_shll_2d.c:322-340)
  results.bin / snapshot_<step>.bin   the same primitives as raw float32 planes behind a 64-byte header (SHLL_SAVE_BIN=1,
                            SHLL_SNAPSHOT_EVERY=k)
-- and, unlike the reference script (which hard-codes NX = NY = 1024), takes the grid shape from the file.  Prints the range
of every field; draws the reference's figures (30 density contours in 2D, the three profiles in 1D) when matplotlib is
installed (`--save fig.png` writes the figure instead of opening a window).

    python tools/fetch_local.py results.dat
    python tools/fetch_local.py snapshot_00000400.bin --save rho.png
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def read_bin(path):
    """-> dict(dims, nx, ny, steps, fields={name: array (nx, ny) or (nx,)})"""
    raw = open(path, "rb").read()
    if raw[:8] != b"SHLLBIN1":
        raise ValueError(f"{path}: not a results.bin / snapshot file")
    dims, nx, ny, ncomp, steps = (int(v) for v in np.frombuffer(raw[8:28], dtype="<i4"))
    planes = np.frombuffer(raw[64:], dtype="<f4")
    if planes.size != ncomp * nx * ny:
        raise ValueError(f"{path}: {planes.size} floats, header says {ncomp} x {nx} x {ny}")
    planes = planes.reshape(ncomp, nx, ny) if dims == 2 else planes.reshape(ncomp, nx)
    names = ["rho", "ux", "uy", "T"] if dims == 2 else ["rho", "u", "T"]
    return dict(dims=dims, nx=nx, ny=ny, steps=steps, fields=dict(zip(names, planes)))


def read_dat(path):
    """The reference's results.dat; the shape comes from the coordinate columns (cell centres (i + 0.5) * DX, i-major)."""
    data = np.genfromtxt(path)
    if data.ndim != 2 or data.shape[1] not in (4, 6):
        raise ValueError(f"{path}: expected 4 (1D) or 6 (2D) columns")
    if data.shape[1] == 4:
        return dict(dims=1, nx=data.shape[0], ny=1, steps=None, x=data[:, 0],
                    fields=dict(rho=data[:, 1], u=data[:, 2], T=data[:, 3]))
    ny = int(np.argmax(data[:, 0] != data[0, 0])) or data.shape[0]   # x is constant along a row of ny cells
    nx = data.shape[0] // ny
    if nx * ny != data.shape[0]:
        raise ValueError(f"{path}: {data.shape[0]} lines do not form rows of {ny} cells")
    g = lambda c: data[:, c].reshape(nx, ny)
    return dict(dims=2, nx=nx, ny=ny, steps=None, x=g(0), y=g(1), fields=dict(rho=g(2), ux=g(3), uy=g(4), T=g(5)))


def read_any(path):
    with open(path, "rb") as f:
        magic = f.read(8)
    return read_bin(path) if magic == b"SHLLBIN1" else read_dat(path)


def summary(r) -> str:
    shape = f"{r['nx']} x {r['ny']}" if r["dims"] == 2 else f"{r['nx']} cells"
    lines = [f"{r['dims']}D, {shape}" + (f", step {r['steps']}" if r["steps"] is not None else "")]
    for name, a in r["fields"].items():
        lines.append(f"  {name:4s} min {a.min(): .6e}  max {a.max(): .6e}  mean {a.mean(): .6e}")
    return "\n".join(lines)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("path")
    ap.add_argument("--save", help="write the figure to this file instead of opening a window")
    ap.add_argument("--no-plot", action="store_true")
    args = ap.parse_args(argv)
    r = read_any(args.path)
    print(summary(r))
    if args.no_plot:
        return 0
    try:
        import matplotlib
        if args.save:
            matplotlib.use("Agg")
        from matplotlib import pyplot as plt
    except ImportError:
        print("(matplotlib is not installed: no figure)", file=sys.stderr)
        return 0
    if r["dims"] == 2:
        x = r.get("x"); y = r.get("y")
        if x is None:
            x, y = np.meshgrid((np.arange(r["nx"]) + 0.5) / r["nx"], (np.arange(r["ny"]) + 0.5) / r["ny"], indexing="ij")
        plt.contour(x, y, r["fields"]["rho"], levels=30)   # the reference's figure: fetch_local.py:35
        plt.gca().set_aspect("equal")
        plt.title("density")
    else:
        x = r.get("x", (np.arange(r["nx"]) + 0.5) / r["nx"])
        fig, ax = plt.subplots(3, 1, sharex=True)
        for a, (name, v) in zip(ax, r["fields"].items()):
            a.plot(x, v)
            a.set_ylabel(name)
    if args.save:
        plt.savefig(args.save, dpi=150)
    else:
        plt.show()
    return 0


if __name__ == "__main__":
    sys.exit(main())
