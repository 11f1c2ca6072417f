"""ctypes binding of the CPU oracle (oracle/shll_oracle.c) and of the compiled reference
programs in oracle/_ref/.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by
the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libshll_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

BC_REFLECT, BC_OUTFLOW = 0, 1
LIM_MINMOD, LIM_MC = 0, 1
TFORM_1D, TFORM_2D = 1, 2
IC_SOD_1D, IC_IMPLOSION, IC_FOUR_SHOCK, IC_CONFIG6, IC_SOD_X_2D = 0, 1, 2, 3, 4


class Cfg(C.Structure):
    _fields_ = [
        ("dims", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("order", C.c_int),
        ("bc", C.c_int), ("limiter", C.c_int), ("tform", C.c_int),
        ("alpha", C.c_float), ("dt_on_dx", C.c_float), ("dt_on_dy", C.c_float),
        ("nthreads", C.c_int),
    ]


def build_lib(force: bool = False) -> str:
    """gcc the restatement into oracle/libshll_oracle.so (no FMA contraction, OpenMP for the sweep)."""
    src = os.path.join(HERE, "shll_oracle.c")
    hdr = os.path.join(HERE, "shll_oracle.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", LIB_PATH, src, "-lm"]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(LIB_PATH)
        FP4 = C.POINTER(C.c_void_p)
        L.shll_oracle_cv.restype = C.c_float
        L.shll_oracle_gamma.restype = C.c_float
        L.shll_oracle_count_steps.restype = C.c_long
        L.shll_oracle_count_steps.argtypes = [C.c_float, C.c_float]
        L.shll_oracle_init.argtypes = [C.POINTER(Cfg), C.c_int, FP4]
        L.shll_oracle_cons_from_prim.argtypes = [C.POINTER(Cfg), FP4, FP4]
        L.shll_oracle_prim_from_cons.argtypes = [C.POINTER(Cfg), FP4, FP4, C.c_void_p]
        L.shll_oracle_run.argtypes = [C.POINTER(Cfg), FP4, C.c_long]
        L.shll_oracle_save_results.argtypes = [C.POINTER(Cfg), FP4, C.c_char_p]
        L.shll_oracle_minmod.restype = C.c_float
        L.shll_oracle_minmod.argtypes = [C.c_float, C.c_float]
        L.shll_oracle_mc.restype = C.c_float
        L.shll_oracle_mc.argtypes = [C.c_float] * 4
        L.shll_oracle_split_flux_2d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.shll_oracle_split_flux_1d.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_cfg(dims, nx, ny=1, order=1, bc=BC_REFLECT, limiter=LIM_MINMOD, tform=None,
             alpha=1.25, dt_on_dx=0.125, dt_on_dy=0.125, nthreads=1) -> Cfg:
    if tform is None:
        tform = TFORM_2D if (dims == 2 or order == 2) else TFORM_1D
    return Cfg(dims, nx, ny if dims == 2 else 1, order, bc, limiter, tform, alpha, dt_on_dx, dt_on_dy, nthreads)


def _ptrs(arr: np.ndarray):
    """arr: (ncomp, ncells) float32 C-contiguous -> void*[4]"""
    assert arr.dtype == np.float32 and arr.flags.c_contiguous and arr.ndim == 2
    p = (C.c_void_p * 4)()
    for k in range(arr.shape[0]):
        p[k] = arr[k].ctypes.data
    return p


def ncomp(cfg: Cfg) -> int:
    return 3 if cfg.dims == 1 else 4


def init_prim(cfg: Cfg, ic: int) -> np.ndarray:
    p = np.zeros((ncomp(cfg), cfg.nx * cfg.ny), np.float32)
    rc = lib().shll_oracle_init(C.byref(cfg), ic, _ptrs(p))
    if rc:
        raise RuntimeError(f"shll_oracle_init rc={rc}")
    return p


def cons_from_prim(cfg: Cfg, p: np.ndarray) -> np.ndarray:
    u = np.zeros_like(p)
    rc = lib().shll_oracle_cons_from_prim(C.byref(cfg), _ptrs(p), _ptrs(u))
    if rc:
        raise RuntimeError(f"cons_from_prim rc={rc}")
    return u


def prim_from_cons(cfg: Cfg, u: np.ndarray, want_a: bool = False):
    p = np.zeros_like(u)
    a = np.zeros(u.shape[1], np.float32) if want_a else None
    rc = lib().shll_oracle_prim_from_cons(C.byref(cfg), _ptrs(u), _ptrs(p), a.ctypes.data if want_a else None)
    if rc:
        raise RuntimeError(f"prim_from_cons rc={rc}")
    return (p, a) if want_a else p


def run(cfg: Cfg, u: np.ndarray, nsteps: int) -> np.ndarray:
    """Returns a new array: u advanced nsteps."""
    out = np.ascontiguousarray(u, dtype=np.float32).copy()
    rc = lib().shll_oracle_run(C.byref(cfg), _ptrs(out), int(nsteps))
    if rc:
        raise RuntimeError(f"shll_oracle_run rc={rc}")
    return out


def run_fma_contracted(cfg: Cfg, u: np.ndarray, nsteps: int):
    """The same restatement built the way an FMA machine builds the reference by default (gcc -O2 -mfma
    -ffp-contract=fast, i.e. what Graviton gives base-c): the size of the difference between its result and run()'s is
    the reference's OWN sensitivity to contraction on that problem, which is the scale the FAST-mode tolerance of the CUDA
    path is stated against (tests/test_oracle_golden.py).  Returns None on a CPU without FMA."""
    try:
        with open("/proc/cpuinfo") as f:
            if " fma" not in f.read():
                return None
    except OSError:
        return None
    path = os.path.join(HERE, "libshll_oracle_fma.so")
    src = os.path.join(HERE, "shll_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=fast", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
                               "-o", path, src, "-lm"])
    L = C.CDLL(path)
    L.shll_oracle_run.argtypes = [C.POINTER(Cfg), C.POINTER(C.c_void_p), C.c_long]
    out = np.ascontiguousarray(u, dtype=np.float32).copy()
    rc = L.shll_oracle_run(C.byref(cfg), _ptrs(out), int(nsteps))
    if rc:
        raise RuntimeError(f"shll_oracle_run (fma build) rc={rc}")
    return out


def count_steps(dt: float, total: float) -> int:
    return int(lib().shll_oracle_count_steps(np.float32(dt), np.float32(total)))


def save_results(cfg: Cfg, p: np.ndarray, path: str) -> None:
    rc = lib().shll_oracle_save_results(C.byref(cfg), _ptrs(p), path.encode())
    if rc:
        raise RuntimeError(f"save_results rc={rc}")


def split_flux_2d(u4, direction: int):
    u = np.ascontiguousarray(u4, np.float32)
    fp = np.zeros(4, np.float32)
    fm = np.zeros(4, np.float32)
    lib().shll_oracle_split_flux_2d(u.ctypes.data, direction, fp.ctypes.data, fm.ctypes.data)
    return fp, fm


def split_flux_1d(u3, tform: int = TFORM_1D):
    u = np.ascontiguousarray(u3, np.float32)
    fp = np.zeros(3, np.float32)
    fm = np.zeros(3, np.float32)
    lib().shll_oracle_split_flux_1d(u.ctypes.data, tform, fp.ctypes.data, fm.ctypes.data)
    return fp, fm


# ----------------------------------------------------------------------------- compiled reference

def ref_available(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


def run_ref(name: str, ncomp_: int, ncells: int, step_cap: int | None = None, threads: int | None = None,
            save: bool = False, raw: bool = True, timeout: float = 3600.0):
    """Run one compiled reference program from oracle/_ref in a scratch directory.

    Returns dict(steps, seconds, u, p, results_dat(bytes|None), stdout).
    """
    exe = os.path.join(REF_DIR, name)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    env = dict(os.environ)
    with tempfile.TemporaryDirectory() as td:
        rawpath = os.path.join(td, "raw.bin")
        if raw:
            env["SHLL_REF_RAW"] = rawpath
        if step_cap is not None:
            env["SHLL_REF_STEP_CAP"] = str(int(step_cap))
        if threads is not None:
            env["SHLL_REF_THREADS"] = str(int(threads))
            env["OMP_NUM_THREADS"] = str(int(threads))
        if save:
            env["SHLL_REF_SAVE"] = "1"
        pr = subprocess.run([exe], cwd=td, env=env, capture_output=True, text=True, timeout=timeout)
        if pr.returncode != 0:
            raise RuntimeError(f"{name} exited {pr.returncode}: {pr.stderr[-400:]}")
        m = re.search(r"REF_TIMING steps=(-?\d+) seconds=([0-9.eE+-]+)", pr.stderr)
        steps_line = re.search(r"Completed in (\d+) steps", pr.stdout)
        out = {
            "steps": int(steps_line.group(1)) if steps_line else None,
            "seconds": float(m.group(2)) if m else None,
            "stdout": pr.stdout,
            "u": None, "p": None, "results_dat": None,
        }
        if raw:
            data = np.fromfile(rawpath, dtype=np.float32)
            assert data.size == 2 * ncomp_ * ncells, (data.size, ncomp_, ncells)
            out["u"] = data[: ncomp_ * ncells].reshape(ncomp_, ncells).copy()
            out["p"] = data[ncomp_ * ncells:].reshape(ncomp_, ncells).copy()
        rd = os.path.join(td, "results.dat")
        if os.path.exists(rd):
            out["results_dat"] = open(rd, "rb").read()
        return out
