"""ctypes binding of libshll_b200.so (include/shll_b200.h).  No CPU fallback: a missing library is an ImportError-like
RuntimeError at first use, a missing GPU is SHLL_E_CUDA from shll_create."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libshll_b200.so")

OK, E_INVAL, E_CUDA, E_NOMEM, E_STATE, E_TIMEOUT = 0, -1, -2, -3, -4, -5
BC_REFLECT, BC_OUTFLOW = 0, 1
LIM_MINMOD, LIM_MC = 0, 1
TFORM_AUTO, TFORM_1D, TFORM_2D = 0, 1, 2
MODE_STRICT, MODE_FAST = 0, 1
IPC_BYTES = 64

# every symbol include/shll_b200.h declares (tests check that the library exports exactly these)
API_SYMBOLS = [
    "shll_abi_version", "shll_last_error", "shll_count_steps", "shll_plan_halo_steps", "shll_create", "shll_destroy", "shll_upload_u",
    "shll_download_u", "shll_download_p", "shll_run", "shll_sync", "shll_run_timed", "shll_max_cfl",
    "shll_conserved_sums", "shll_selftest_exact_division",
    "shll_halo_wait_stats", "shll_launch_count", "shll_variant_name", "shll_peer_export", "shll_peer_connect",
    "shll_group_create", "shll_group_destroy", "shll_group_last_error", "shll_group_size", "shll_group_ctx",
    "shll_group_upload_u", "shll_group_download_u", "shll_group_download_p", "shll_group_run", "shll_group_run_timed",
    "shll_group_max_cfl", "shll_group_conserved_sums", "shll_group_launch_count",
]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dims", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
        ("order", C.c_int32), ("bc", C.c_int32), ("limiter", C.c_int32), ("tform", C.c_int32), ("mode", C.c_int32),
        ("alpha", C.c_float), ("dt_on_dx", C.c_float), ("dt_on_dy", C.c_float),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("variant", C.c_int32), ("halo_steps", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class PeerDesc(C.Structure):
    _fields_ = [
        ("state_handle", C.c_uint8 * IPC_BYTES), ("flag_handle", C.c_uint8 * IPC_BYTES),
        ("pid", C.c_int64), ("state_ptr", C.c_uint64), ("flag_ptr", C.c_uint64),
        ("device", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("dims", C.c_int32), ("order", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class ShllError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libshll_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built -- there is deliberately no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or python shll_sve_cfd_b200/build.py). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        VPP = C.POINTER(C.c_void_p)
        L.shll_abi_version.restype = C.c_int
        L.shll_last_error.restype = C.c_char_p
        L.shll_last_error.argtypes = [C.c_void_p]
        L.shll_count_steps.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_long)]
        L.shll_plan_halo_steps.argtypes = [C.POINTER(Config), C.c_int]
        L.shll_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.shll_destroy.argtypes = [C.c_void_p]
        L.shll_upload_u.argtypes = [C.c_void_p, VPP]
        L.shll_download_u.argtypes = [C.c_void_p, VPP]
        L.shll_download_p.argtypes = [C.c_void_p, VPP, C.c_void_p]
        L.shll_run.argtypes = [C.c_void_p, C.c_long]
        L.shll_sync.argtypes = [C.c_void_p]
        L.shll_run_timed.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_float)]
        L.shll_max_cfl.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.shll_conserved_sums.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.shll_halo_wait_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.shll_selftest_exact_division.argtypes = [C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
        L.shll_launch_count.restype = C.c_long
        L.shll_launch_count.argtypes = [C.c_void_p]
        L.shll_variant_name.restype = C.c_char_p
        L.shll_variant_name.argtypes = [C.c_void_p]
        L.shll_peer_export.argtypes = [C.c_void_p, C.POINTER(PeerDesc)]
        L.shll_peer_connect.argtypes = [C.c_void_p, C.c_int, C.POINTER(PeerDesc)]
        L.shll_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_int, C.POINTER(C.c_int)]
        L.shll_group_destroy.argtypes = [C.c_void_p]
        L.shll_group_last_error.restype = C.c_char_p
        L.shll_group_last_error.argtypes = [C.c_void_p]
        L.shll_group_size.argtypes = [C.c_void_p]
        L.shll_group_ctx.restype = C.c_void_p
        L.shll_group_ctx.argtypes = [C.c_void_p, C.c_int]
        L.shll_group_upload_u.argtypes = [C.c_void_p, VPP]
        L.shll_group_download_u.argtypes = [C.c_void_p, VPP]
        L.shll_group_download_p.argtypes = [C.c_void_p, VPP, C.c_void_p]
        L.shll_group_run.argtypes = [C.c_void_p, C.c_long]
        L.shll_group_run_timed.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_float)]
        L.shll_group_max_cfl.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.shll_group_conserved_sums.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.shll_group_launch_count.restype = C.c_long
        L.shll_group_launch_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def count_steps(dt, total_time) -> int:
    n = C.c_long(0)
    rc = lib().shll_count_steps(C.c_float(np.float32(dt)), C.c_float(np.float32(total_time)), C.byref(n))
    if rc:
        raise ShllError(rc, lib().shll_last_error(None).decode())
    return n.value


def plan_halo_steps(dims, nx_global, ny, order, mode, nslabs, bc=BC_REFLECT, limiter=LIM_MINMOD) -> int:
    """shll_plan_halo_steps: the halo_steps value every slab of a domain cut into `nslabs` must carry."""
    cfg = Config()
    cfg.struct_size = C.sizeof(Config)
    cfg.dims, cfg.nx, cfg.ny, cfg.order, cfg.mode, cfg.bc, cfg.limiter = dims, nx_global, (ny if dims == 2 else 1), order, mode, bc, limiter
    cfg.dt_on_dx = cfg.dt_on_dy = 0.125
    cfg.nranks = 1
    return int(lib().shll_plan_halo_steps(C.byref(cfg), int(nslabs)))


def selftest_exact_division(npairs: int, seed: int = 1, device: int = 0) -> dict:
    """Device self-test of the STRICT exactness shortcuts against IEEE division (csrc/selftest.cu)."""
    c = (C.c_ulonglong * 9)()
    rc = lib().shll_selftest_exact_division(int(device), int(npairs), int(seed), c)
    if rc:
        raise ShllError(rc, "shll_selftest_exact_division failed")
    keys = ("div_rn_shared_mismatch", "div_rn_spec_mismatch", "div_by_cv_mismatch", "div_by_cv_spec_mismatch",
            "float_flagged", "double_flagged", "float_slow_path", "float_pairs", "doubles")
    return dict(zip(keys, [int(v) for v in c]))


def _ptrs(arr: np.ndarray):
    assert arr.dtype == np.float32 and arr.ndim == 2 and arr.flags.c_contiguous
    p = (C.c_void_p * 4)()
    for k in range(arr.shape[0]):
        p[k] = arr[k].ctypes.data
    return p


class Solver:
    """One slab on one GPU.  Mirrors the C usage: create -> upload_u -> run(nsteps) -> download_u/p -> destroy."""

    def __init__(self, dims, nx, ny=1, order=1, bc=BC_REFLECT, limiter=LIM_MINMOD, tform=TFORM_AUTO,
                 mode=MODE_STRICT, alpha=1.25, dt_on_dx=0.125, dt_on_dy=0.125, device=0, rank=0, nranks=1, variant=0, halo_steps=0):
        self.cfg = Config()
        self.cfg.struct_size = C.sizeof(Config)
        self.cfg.dims, self.cfg.nx, self.cfg.ny = dims, nx, (ny if dims == 2 else 1)
        self.cfg.order, self.cfg.bc, self.cfg.limiter, self.cfg.tform, self.cfg.mode = order, bc, limiter, tform, mode
        self.cfg.alpha, self.cfg.dt_on_dx, self.cfg.dt_on_dy = alpha, dt_on_dx, dt_on_dy
        self.cfg.device, self.cfg.rank, self.cfg.nranks, self.cfg.variant = device, rank, nranks, variant
        self.cfg.halo_steps = halo_steps
        self.ncomp = 3 if dims == 1 else 4
        self.ncells = nx * (ny if dims == 2 else 1)
        self._h = C.c_void_p(None)
        rc = lib().shll_create(C.byref(self._h), C.byref(self.cfg))
        if rc:
            raise ShllError(rc, lib().shll_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise ShllError(rc, lib().shll_last_error(self._h).decode())

    def upload_u(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float32)
        assert u.shape == (self.ncomp, self.ncells), (u.shape, self.ncomp, self.ncells)
        self._ck(lib().shll_upload_u(self._h, _ptrs(u)))

    def download_u(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.ncomp, self.ncells), np.float32)
        self._ck(lib().shll_download_u(self._h, _ptrs(out)))
        return out

    def download_p(self, want_a=False):
        p = np.empty((self.ncomp, self.ncells), np.float32)
        a = np.empty(self.ncells, np.float32) if want_a else None
        self._ck(lib().shll_download_p(self._h, _ptrs(p), a.ctypes.data if want_a else None))
        return (p, a) if want_a else p

    def run(self, nsteps: int):
        self._ck(lib().shll_run(self._h, int(nsteps)))

    def sync(self):
        self._ck(lib().shll_sync(self._h))

    def run_timed(self, nsteps: int) -> float:
        ms = C.c_float(0)
        self._ck(lib().shll_run_timed(self._h, int(nsteps), C.byref(ms)))
        return ms.value

    def max_cfl(self) -> float:
        v = C.c_float(0)
        self._ck(lib().shll_max_cfl(self._h, C.byref(v)))
        return v.value

    def conserved_sums(self) -> np.ndarray:
        """FP64 sums of the conserved components over the owned cells (mass, momentum, energy): a monitor, like max_cfl."""
        v = (C.c_double * 4)()
        self._ck(lib().shll_conserved_sums(self._h, v))
        return np.array(v[:], dtype=np.float64)

    def halo_wait_stats(self) -> dict:
        """Seconds the edge warps spent spinning on the neighbours' halo flags since creation (multi-GPU attribution)."""
        v = (C.c_double * 3)()
        self._ck(lib().shll_halo_wait_stats(self._h, v))
        return {"lower_s": v[0], "upper_s": v[1], "waits": int(v[2])}

    @property
    def launches(self) -> int:
        return lib().shll_launch_count(self._h)

    @property
    def variant(self) -> str:
        return lib().shll_variant_name(self._h).decode()

    def peer_export(self) -> bytes:
        d = PeerDesc()
        self._ck(lib().shll_peer_export(self._h, C.byref(d)))
        return bytes(d)

    def peer_connect(self, side: int, desc: bytes):
        d = PeerDesc.from_buffer_copy(desc)
        self._ck(lib().shll_peer_connect(self._h, side, C.byref(d)))

    def close(self):
        if self._h:
            lib().shll_destroy(self._h)
            self._h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """Single-process multi-GPU: the whole domain behind one handle (shll_group_*), slabs along x on `devices`.

    Arrays are the GLOBAL SoA arrays (ncomp, nx*ny); `devices` may repeat a device (slabs then share it)."""

    def __init__(self, dims, nx, ny=1, ngpus=1, devices=None, order=1, bc=BC_REFLECT, limiter=LIM_MINMOD, tform=TFORM_AUTO,
                 mode=MODE_STRICT, alpha=1.25, dt_on_dx=0.125, dt_on_dy=0.125, variant=0):
        self.cfg = Config()
        self.cfg.struct_size = C.sizeof(Config)
        self.cfg.dims, self.cfg.nx, self.cfg.ny = dims, nx, (ny if dims == 2 else 1)
        self.cfg.order, self.cfg.bc, self.cfg.limiter, self.cfg.tform, self.cfg.mode = order, bc, limiter, tform, mode
        self.cfg.alpha, self.cfg.dt_on_dx, self.cfg.dt_on_dy = alpha, dt_on_dx, dt_on_dy
        self.cfg.device, self.cfg.rank, self.cfg.nranks, self.cfg.variant = 0, 0, 1, variant
        self.ncomp = 3 if dims == 1 else 4
        self.ncells = nx * (ny if dims == 2 else 1)
        self._h = C.c_void_p(None)
        dev = None
        if devices is not None:
            if len(devices) != ngpus:
                raise ValueError("len(devices) != ngpus")
            dev = (C.c_int * ngpus)(*devices)
        rc = lib().shll_group_create(C.byref(self._h), C.byref(self.cfg), int(ngpus), dev)
        if rc:
            raise ShllError(rc, lib().shll_group_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise ShllError(rc, lib().shll_group_last_error(self._h).decode())

    def upload_u(self, u: np.ndarray):
        u = np.ascontiguousarray(u, dtype=np.float32)
        assert u.shape == (self.ncomp, self.ncells), (u.shape, self.ncomp, self.ncells)
        self._ck(lib().shll_group_upload_u(self._h, _ptrs(u)))

    def download_u(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.ncomp, self.ncells), np.float32)
        self._ck(lib().shll_group_download_u(self._h, _ptrs(out)))
        return out

    def download_p(self, want_a=False):
        p = np.empty((self.ncomp, self.ncells), np.float32)
        a = np.empty(self.ncells, np.float32) if want_a else None
        self._ck(lib().shll_group_download_p(self._h, _ptrs(p), a.ctypes.data if want_a else None))
        return (p, a) if want_a else p

    def run(self, nsteps: int):
        self._ck(lib().shll_group_run(self._h, int(nsteps)))

    def run_timed(self, nsteps: int) -> float:
        ms = C.c_float(0)
        self._ck(lib().shll_group_run_timed(self._h, int(nsteps), C.byref(ms)))
        return ms.value

    def max_cfl(self) -> float:
        v = C.c_float(0)
        self._ck(lib().shll_group_max_cfl(self._h, C.byref(v)))
        return v.value

    def conserved_sums(self) -> np.ndarray:
        v = (C.c_double * 4)()
        self._ck(lib().shll_group_conserved_sums(self._h, v))
        return np.array(v[:], dtype=np.float64)

    @property
    def size(self) -> int:
        return lib().shll_group_size(self._h)

    @property
    def launches(self) -> int:
        return lib().shll_group_launch_count(self._h)

    def variant(self, slab: int = 0) -> str:
        return lib().shll_variant_name(lib().shll_group_ctx(self._h, slab)).decode()

    def close(self):
        if self._h:
            lib().shll_group_destroy(self._h)
            self._h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
