"""Host-side logic of the reference programs, restated with explicit float32/float64 steps.

Each reference solver is a C file with compile-time constants, an initial condition, Compute_U_from_P, a float
clock and (optionally) Save_Results; only the three calls inside the time loop run on the GPU.  This module is the
Python twin of the C host programs in host/ (same names, same arithmetic) so tests and bench.py can drive
libshll_b200.so without compiling a C main per problem size.

Arithmetic notes: unsuffixed C literals are double, `float op double` is double, assignment
to float rounds once.  numpy float32/float64 scalars and arrays reproduce that exactly when every step is
spelled out, which is what the functions below do.  tests/test_host_logic.py checks them bit for bit against
the CPU oracle, which is itself pinned to the compiled reference.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

from . import capi

f32, f64 = np.float32, np.float64

# base_shll.c:17-22
R = f32(1.0)
GAMMA = f32(1.4)
CV = f32(f64(R) / (f64(GAMMA) - 1.0))
CFL = f32(0.25)


@dataclass(frozen=True)
class Problem:
    """One reference program = scheme + initial condition + end time."""
    name: str
    dims: int
    nx: int
    ny: int = 1
    order: int = 1
    bc: int = capi.BC_REFLECT
    limiter: int = capi.LIM_MINMOD
    alpha: float = 1.25
    ic: str = "sod_1d"
    total_time: float = 0.2
    tform: int = capi.TFORM_AUTO
    lx: float = 1.0
    ly: float = 1.0

    @property
    def ncomp(self) -> int:
        return 3 if self.dims == 1 else 4

    @property
    def ncells(self) -> int:
        return self.nx * (self.ny if self.dims == 2 else 1)

    def resized(self, nx: int, ny: int | None = None) -> "Problem":
        return replace(self, nx=nx, ny=(ny if ny is not None else (nx if self.dims == 2 else 1)))


# The four in-scope reference programs (as checked in) and the derived 1D 2nd-order program.
BASE_SHLL = Problem("base_shll", 1, 256, ic="sod_1d", total_time=0.2)                                   # base-c/base_shll.c
BASE_SHLL_2D = Problem("base_shll_2d", 2, 256, 256, ic="implosion", total_time=0.1)                     # base-c/base_shll_2d.c
SECOND_ORDER_2D = Problem("2nd_order_base_shll", 2, 256, 256, order=2, bc=capi.BC_OUTFLOW,
                          ic="four_shock", total_time=0.8)                                              # base-c/2nd_order_base_shll.c
SECOND_ORDER_1D = Problem("2nd_order_base_shll_1d", 1, 256, order=2, bc=capi.BC_OUTFLOW, ic="sod_1d",
                          total_time=0.2, tform=capi.TFORM_2D)                                          # derived: x-sweep of the above
BASE_OMP_2D = Problem("base_omp_2nd_order", 2, 1024, 1024, order=2, bc=capi.BC_OUTFLOW,
                      limiter=capi.LIM_MC, alpha=1.25, ic="config6", total_time=0.3)                    # base-omp/2nd_order_base_shll.c
PROGRAMS = {p.name: p for p in (BASE_SHLL, BASE_SHLL_2D, SECOND_ORDER_2D, SECOND_ORDER_1D, BASE_OMP_2D)}


def time_constants(pb: Problem):
    """DX, DY, DT, DT_ON_DX, DT_ON_DY as Compute_U_from_P sets them (base_shll.c:21,83-84; base_shll_2d.c:30-31,134-136)."""
    dx = f32(pb.lx) / f32(pb.nx)
    dy = f32(pb.ly) / f32(pb.ny) if pb.dims == 2 else f32(1.0)
    dt_on_dx = CFL / (R + f32(1.0))
    dt = f32(dt_on_dx * dx)
    dt_on_dy = f32(dt / dy) if pb.dims == 2 else dt_on_dx
    return dx, dy, dt, f32(dt_on_dx), dt_on_dy


def count_steps(pb: Problem) -> int:
    """NO_STEPS of `while (time < TOTAL_TIME) time += DT` with the float clock (base_shll.c:200-217)."""
    _, _, dt, _, _ = time_constants(pb)
    t, n = f32(0.0), 0
    total = f32(pb.total_time)
    while t < total:
        tn = f32(t + dt)
        if tn == t:
            raise ValueError(f"float clock stalls at t={t} (N={pb.nx}); use a fixed step count (SURVEY.md T4)")
        t = tn
        n += 1
    return n


def initial_primitives(pb: Problem, i0: int = 0, nx_local: int | None = None, nx_global: int | None = None) -> np.ndarray:
    """Allocate_and_Init_Memory's initial condition; rows [i0, i0+nx_local) of a global grid of nx_global rows."""
    NX = nx_global if nx_global is not None else pb.nx
    nl = nx_local if nx_local is not None else pb.nx
    i = np.arange(i0, i0 + nl, dtype=np.int64)
    if pb.dims == 1:
        assert pb.ic == "sod_1d"
        p = np.empty((3, nl), f32)
        p[0] = np.where(i < 0.5 * NX, f32(10.0), f32(1.0))  # base_shll.c:55-59
        p[1] = 0.0
        p[2] = 1.0
        return p
    NY = pb.ny
    ii = i[:, None].astype(f64)
    jj = np.arange(NY, dtype=np.int64)[None, :].astype(f64)
    p = np.empty((4, nl, NY), f32)
    if pb.ic == "implosion":  # base_shll_2d.c:96-100
        inside = (ii > 0.2 * NX) & (ii < 0.8 * NX) & (jj > 0.2 * NY) & (jj < 0.8 * NY)
        p[0] = np.where(inside, f32(1.0), f32(10.0))
        p[1] = 0.0
        p[2] = 0.0
        p[3] = 1.0
    elif pb.ic in ("four_shock", "config6"):
        if pb.ic == "four_shock":  # 2nd_order_base_shll.c:137-145
            frac = 0.75
            states = [(0.138, 1.206, 1.206, 0.029), (0.5323, 0.0, 1.206, 0.3), (0.5323, 1.206, 0.0, 0.3), (1.5, 0.0, 0.0, 1.5)]
        else:  # base-omp/2nd_order_base_shll.c:149-167
            frac = 0.5
            states = [(1.0, -0.75, 0.5, 1.0), (3.0, -0.75, -0.5, 1.0), (2.0, 0.75, 0.5, 1.0), (1.0, 0.75, -0.5, 1.0)]
        c1 = (ii < frac * NX) & (jj < frac * NY)
        c2 = (ii > frac * NX) & (jj < frac * NY) & ~c1
        c3 = (ii < frac * NX) & (jj > frac * NY) & ~c1 & ~c2
        sel = np.where(c1, 0, np.where(c2, 1, np.where(c3, 2, 3)))
        for k in range(3):
            p[k] = np.choose(sel, [f32(s[k]) for s in states])
        # p3 = (pressure/(p0*R)): double divide of the literal by the float product
        pr = np.choose(sel, [f64(s[3]) for s in states])
        p[3] = (pr / (p[0] * R).astype(f64)).astype(f32)
    elif pb.ic == "sod_x":  # y-uniform Sod (base-omp/2nd_order_base_shll.c:138-142, commented out there)
        p[0] = np.where(ii < 0.5 * NX, f32(10.0), f32(1.0)) + np.zeros((1, NY), f32)
        p[1] = 0.0
        p[2] = 0.0
        p[3] = 1.0
    else:
        raise ValueError(pb.ic)
    return p.reshape(4, nl * NY)


def cons_from_prim(pb: Problem, p: np.ndarray) -> np.ndarray:
    """Compute_U_from_P (base_shll.c:76-80; base_shll_2d.c:126-131)."""
    u = np.empty_like(p)
    if pb.dims == 1:
        p0, p1, p2 = p
        u[0] = p0
        u[1] = p0 * p1
        inner = (p2 * CV).astype(f64) + (0.5 * p1.astype(f64)) * p1.astype(f64)
        u[2] = (p0.astype(f64) * inner).astype(f32)
    else:
        p0, p1, p2, p3 = p
        u[0] = p0
        u[1] = p0 * p1
        u[2] = p0 * p2
        ke = (p1 * p1 + p2 * p2).astype(f32)
        inner = (p3 * CV).astype(f64) + 0.5 * ke.astype(f64)
        u[3] = (p0.astype(f64) * inner).astype(f32)
    return u


def prim_from_cons(pb: Problem, u: np.ndarray, tform: int | None = None):
    """Compute_P_from_U on the host (base_shll.c:171-176; base_shll_2d.c:312-318).  Returns (p, a)."""
    p = np.empty_like(u)
    with np.errstate(all="ignore"):
        if pb.dims == 1:
            tf = tform if tform is not None else pb.tform
            if tf == capi.TFORM_AUTO:
                tf = capi.TFORM_1D if pb.order == 1 else capi.TFORM_2D
            u0, u1, u2 = u
            p[0] = u0
            p[1] = u1 / u0
            e = (u2 / u0).astype(f64)
            if tf == capi.TFORM_1D:
                num = e - (0.5 * p[1].astype(f64)) * p[1].astype(f64)
            else:
                num = e - 0.5 * (p[1] * p[1]).astype(f64)
            p[2] = (num / f64(CV)).astype(f32)
            T = p[2]
        else:
            u0, u1, u2, u3 = u
            p[0] = u0
            p[1] = u1 / u0
            p[2] = u2 / u0
            ke = (p[1] * p[1] + p[2] * p[2]).astype(f32)
            p[3] = (((u3 / u0).astype(f64) - 0.5 * ke.astype(f64)) / f64(CV)).astype(f32)
            T = p[3]
        a = np.sqrt((GAMMA * R * T).astype(f64)).astype(f32)
    return p, a


def save_results(pb: Problem, p: np.ndarray, path: str) -> None:
    """Save_Results: results.dat, `%e` tab separated (base_shll.c:186-189; base_shll_2d.c:329-336)."""
    dx, dy, _, _, _ = time_constants(pb)
    with open(path, "w") as f:
        if pb.dims == 1:
            cx = ((np.arange(pb.nx) + 0.5) * f64(dx)).astype(f32)
            for i in range(pb.nx):
                f.write("%e\t%e\t%e\t%e\n" % (cx[i], p[0][i], p[1][i], p[2][i]))
        else:
            cx = ((np.arange(pb.nx) + 0.5) * f64(dx)).astype(f32)
            cy = ((np.arange(pb.ny) + 0.5) * f64(dy)).astype(f32)
            c = 0
            for i in range(pb.nx):
                for j in range(pb.ny):
                    f.write("%e\t%e\t%e\t%e\t%e\t%e\n" % (cx[i], cy[j], p[0][c], p[1][c], p[2][c], p[3][c]))
                    c += 1


def halo_steps_for(pb: Problem, nranks: int, mode: int = capi.MODE_STRICT) -> int:
    """shll_config.halo_steps for a domain of pb.nx rows / cells cut into nranks balanced slabs (include/shll_b200.h:
    shll_plan_halo_steps): the same value on every rank."""
    if nranks <= 1:
        return 1
    return capi.plan_halo_steps(pb.dims, pb.nx, pb.ny, pb.order, mode, nranks, bc=pb.bc, limiter=pb.limiter)


def make_solver(pb: Problem, mode: int = capi.MODE_STRICT, device: int = 0, rank: int = 0, nranks: int = 1,
                nx_local: int | None = None, variant: int = 0, halo_steps: int = 0) -> capi.Solver:
    _, _, _, dt_on_dx, dt_on_dy = time_constants(pb)
    return capi.Solver(pb.dims, nx_local if nx_local is not None else pb.nx, pb.ny, order=pb.order, bc=pb.bc,
                       limiter=pb.limiter, tform=pb.tform, mode=mode, alpha=pb.alpha, dt_on_dx=float(dt_on_dx),
                       dt_on_dy=float(dt_on_dy), device=device, rank=rank, nranks=nranks, variant=variant, halo_steps=halo_steps)


def run_program(pb: Problem, mode: int = capi.MODE_STRICT, nsteps: int | None = None, device: int = 0, variant: int = 0):
    """The reference main() (base_shll.c:198-225) with the time loop on the GPU.

    Returns dict(steps, u, p, a): final conserved and primitive arrays, exactly what the C program would hold in
    u0..u3 / p0..p3 / a when it prints `Completed in %d steps`.
    """
    steps = count_steps(pb) if nsteps is None else int(nsteps)
    p0 = initial_primitives(pb)              # Allocate_and_Init_Memory
    u0 = cons_from_prim(pb, p0)              # Compute_U_from_P
    with make_solver(pb, mode, device, variant=variant) as s:
        s.upload_u(u0)
        s.run(steps)                         # the whole `while (time < TOTAL_TIME)` loop
        u = s.download_u()
        p, a = s.download_p(want_a=True)     # the last Compute_P_from_U
        info = dict(variant=s.variant, launches=s.launches)
    return dict(steps=steps, u=u, p=p, a=a, **info)
