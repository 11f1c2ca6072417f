#!/usr/bin/env python3
"""Build the reference's own solver programs into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's problem sizes are compile-time `const`s (base-c/base_shll.c:16-26), so each
size needs its own binary.  This recipe reads the sources where they lie under
/root/reference, writes *parametrised copies* into git-ignored oracle/_ref/gen/ and compiles
them with the reference's own flags (`gcc X.c -lm -O3`, base-c/makefile:2,4;
`-fopenmp` for base-omp/makefile:2,4).  Nothing from /root/reference is ever written
into tracked files.

Edits applied to the copies (string substitutions only, each asserted to match exactly once):
  * grid size / TOTAL_TIME constants
  * a run-time step cap        (env SHLL_REF_STEP_CAP) on the `while (time < TOTAL_TIME)` loop
  * wall-clock timing of the loop, printed on stderr as `REF_TIMING steps=<n> seconds=<s>`
  * optional raw float32 dump of u* then p* arrays (env SHLL_REF_RAW=<path>)
  * Save_Results() made switchable (env SHLL_REF_SAVE) where the reference has it commented out
  * base-omp only: thread count from env SHLL_REF_THREADS; `cell_index = 0`
    (uninitialised in the reference, base-omp/2nd_order_base_shll.c:47,136)
  * y-uniform "1D slice" variant of the 2D 2nd-order file: NY=4,
    H=NY*DX, Sod IC in x -- the oracle for the derived 1D 2nd-order program.

On the GPU box /root/reference does not exist: the binaries built here travel with the repo
snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored) and this script is a no-op.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SHLL_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")

PRELUDE = r'''
/* ---- harness prelude inserted by oracle/build_ref.py (not reference code) ---- */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
static int ref_step_cap = 2147483647;
static int ref_threads_n = 1;
static struct timespec ref_t0;
static void ref_tic(void) {
    const char *s = getenv("SHLL_REF_STEP_CAP");
    if (s) ref_step_cap = atoi(s);
    clock_gettime(CLOCK_MONOTONIC, &ref_t0);
}
static int ref_threads(void) {
    const char *s = getenv("SHLL_REF_THREADS");
    if (s) ref_threads_n = atoi(s);
    return ref_threads_n;
}
static void ref_toc(int steps) {
    struct timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double dt = (t1.tv_sec - ref_t0.tv_sec) + 1e-9 * (t1.tv_nsec - ref_t0.tv_nsec);
    fprintf(stderr, "REF_TIMING steps=%d seconds=%.9f threads=%d\n", steps, dt, ref_threads_n);
}
static void ref_dump_raw(int ncomp, long n, float **u, float **p) {
    const char *path = getenv("SHLL_REF_RAW");
    if (!path) return;
    FILE *f = fopen(path, "wb");
    for (int k = 0; k < ncomp; k++) fwrite(u[k], sizeof(float), n, f);
    for (int k = 0; k < ncomp; k++) fwrite(p[k], sizeof(float), n, f);
    fclose(f);
}
/* ---- end prelude ---- */
'''


def sub1(text: str, old: str, new: str, count: int = 1) -> str:
    n = text.count(old)
    if n != count:
        raise RuntimeError(f"pattern {old!r} matched {n} times, expected {count}")
    return text.replace(old, new)


def gen_1d(n: int) -> str:
    src = open(os.path.join(REF_ROOT, "base-c/base_shll.c")).read()
    src = sub1(src, "const int N = 256;", f"const int N = {n};")
    src = sub1(src, "while (time < TOTAL_TIME) {", "ref_tic();\n    while (time < TOTAL_TIME && NO_STEPS < ref_step_cap) {")
    src = sub1(src, 'printf("Completed in %d steps\\n", NO_STEPS);',
               'ref_toc(NO_STEPS);\n    printf("Completed in %d steps\\n", NO_STEPS);\n'
               '    { float *uu[3] = {u0,u1,u2}; float *pp[3] = {p0,p1,p2}; ref_dump_raw(3, N, uu, pp); }')
    return PRELUDE + src


def _common_2d(src: str, nx: int, ny: int) -> str:
    src = sub1(src, "const int NX = 256;", f"const int NX = {nx};")
    src = sub1(src, "const int NY = 256;", f"const int NY = {ny};")
    return src


def gen_2d_o1(nx: int, ny: int) -> str:
    src = open(os.path.join(REF_ROOT, "base-c/base_shll_2d.c")).read()
    src = _common_2d(src, nx, ny)
    src = sub1(src, "while (time < TOTAL_TIME) {", "ref_tic();\n    while (time < TOTAL_TIME && NO_STEPS < ref_step_cap) {")
    src = sub1(src, 'printf("Completed in %d steps\\n", NO_STEPS);',
               'ref_toc(NO_STEPS);\n    printf("Completed in %d steps\\n", NO_STEPS);\n'
               '    { float *uu[4] = {u0,u1,u2,u3}; float *pp[4] = {p0,p1,p2,p3}; ref_dump_raw(4, N, uu, pp); }')
    src = sub1(src, "    //Save_Results();", '    if (getenv("SHLL_REF_SAVE")) Save_Results();')
    return PRELUDE + src


def gen_2d_o2(nx: int, ny: int, slice1d: bool = False, total_time: str | None = None) -> str:
    src = open(os.path.join(REF_ROOT, "base-c/2nd_order_base_shll.c")).read()
    src = _common_2d(src, nx, ny)
    if total_time is not None:
        src = sub1(src, "const float TOTAL_TIME = 0.8;", f"const float TOTAL_TIME = {total_time};")
    if slice1d:
        # SURVEY.md App. B row "1D 2nd-order oracle": y-uniform run, H = NY*DX, Sod IC in x.
        src = sub1(src, "const float H = 1.0;", f"const float H = {ny}.0/{nx}.0;")
        a = src.index("            if ((i < 0.75*NX) && (j < 0.75*NY)) {")
        b = src.index("            cell_index++;", a)
        ic = ("            if (i < 0.5*NX) {\n"
              "                p0[cell_index] = 10.0; p1[cell_index] = 0.0; p2[cell_index] = 0.0; p3[cell_index] = 1.0;\n"
              "            } else {\n"
              "                p0[cell_index] = 1.0; p1[cell_index] = 0.0; p2[cell_index] = 0.0; p3[cell_index] = 1.0;\n"
              "            }\n")
        src = src[:a] + ic + src[b:]
    src = sub1(src, "while (time < TOTAL_TIME) {", "ref_tic();\n    while (time < TOTAL_TIME && NO_STEPS < ref_step_cap) {")
    src = sub1(src, 'printf("Completed in %d steps\\n", NO_STEPS);',
               'ref_toc(NO_STEPS);\n    printf("Completed in %d steps\\n", NO_STEPS);\n'
               '    { float *uu[4] = {u0,u1,u2,u3}; float *pp[4] = {p0,p1,p2,p3}; ref_dump_raw(4, N, uu, pp); }')
    src = sub1(src, "    // Save_Results();", '    if (getenv("SHLL_REF_SAVE")) Save_Results();')
    return PRELUDE + src


def gen_omp(nx: int, ny: int) -> str:
    src = open(os.path.join(REF_ROOT, "base-omp/2nd_order_base_shll.c")).read()
    src = sub1(src, "const int NX = 1024;", f"const int NX = {nx};")
    src = sub1(src, "const int NY = 1024;", f"const int NY = {ny};")
    src = sub1(src, "size_t alignment = 32; int i, j, cell_index;", "size_t alignment = 32; int i, j, cell_index = 0;")
    src = sub1(src, "omp_set_num_threads(16);", "omp_set_num_threads(ref_threads());\n    ref_tic();")
    src = sub1(src, "while (time < TOTAL_TIME) {", "while (time < TOTAL_TIME && NO_STEPS < ref_step_cap) {")
    src = sub1(src, "    // Save_Results();\n\n    // Free",
               "    ref_toc(-1);\n"
               "    { float *uu[4] = {u0,u1,u2,u3}; float *pp[4] = {p0,p1,p2,p3}; ref_dump_raw(4, N, uu, pp); }\n"
               '    if (getenv("SHLL_REF_SAVE")) Save_Results();\n\n    // Free')
    return PRELUDE + src


# name -> (generator, args, openmp)
SPECS = {}
for n in (256, 1024, 8192, 65536):
    SPECS[f"ref_1d_o1_{n}"] = (gen_1d, (n,), False)
for n in (64, 128, 256, 1024, 4096):
    SPECS[f"ref_2d_o1_{n}"] = (gen_2d_o1, (n, n), False)
SPECS["ref_2d_o1_96x160"] = (gen_2d_o1, (96, 160), False)
for n in (64, 128, 256, 1024, 4096):
    SPECS[f"ref_2d_o2_{n}"] = (gen_2d_o2, (n, n), False)
SPECS["ref_2d_o2_96x160"] = (gen_2d_o2, (96, 160), False)
for n in (1024, 4096, 65536):
    SPECS[f"ref_1d_o2_slice_{n}"] = (gen_2d_o2, (n, 4, True, "0.2"), False)
for n in (64, 128, 1024, 4096):
    SPECS[f"ref_omp_o2_{n}"] = (gen_omp, (n, n), True)


def build(names=None, verbose=True) -> bool:
    """Build the reference binaries. Returns False (and does nothing) when /root/reference is absent."""
    if not os.path.isdir(REF_ROOT):
        if verbose:
            print(f"[build_ref] {REF_ROOT} not present: using prebuilt oracle/_ref/ binaries", file=sys.stderr)
        return False
    os.makedirs(GEN, exist_ok=True)
    for name, (gen, args, omp) in SPECS.items():
        if names and name not in names:
            continue
        exe = os.path.join(OUT, name)
        csrc = os.path.join(GEN, name + ".c")
        text = gen(*args)
        if os.path.exists(csrc) and os.path.exists(exe) and open(csrc).read() == text:
            continue
        with open(csrc, "w") as f:
            f.write(text)
        # Reference flags: base-c/makefile:2,4 ("gcc base_shll.c -lm -O3"); base-omp/makefile:2,4 adds -fopenmp.
        # -mcmodel not needed: arrays are heap-allocated.  -w: the reference has unused variables.
        cmd = ["gcc", csrc, "-O3", "-w", "-o", exe, "-lm"]
        if omp:
            cmd.insert(2, "-fopenmp")
        if verbose:
            print("[build_ref]", " ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    build(sys.argv[1:] or None)
