"""One-process-per-GPU slab decomposition.

The domain is cut into contiguous slabs along x (the slow axis: rows of ny contiguous floats, base_shll_2d.c:157).
torch.distributed is plumbing only: it carries the 184-byte CUDA-IPC descriptors between neighbouring ranks and
provides barriers.  The data path has no collective: every step, the edge warps of the step kernel store `order`
rows straight into the neighbour GPU's halo rows over NVLink and raise a flag (csrc/halo_sync.cuh).

The partition / descriptor logic is independent of CUDA and is exercised on CPU with the gloo backend
(tests/test_slabs_cpu.py) by substituting the solver factory.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Slab:
    rank: int
    nranks: int
    i0: int        # first global row owned
    nx_local: int  # rows owned
    nx_global: int

    @property
    def lower(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def upper(self):
        return self.rank + 1 if self.rank < self.nranks - 1 else None


def partition(nx_global: int, nranks: int, rank: int) -> Slab:
    """Balanced contiguous partition: slab sizes differ by at most one row."""
    if nranks < 1 or not (0 <= rank < nranks):
        raise ValueError("bad rank/nranks")
    i0 = (rank * nx_global) // nranks
    i1 = ((rank + 1) * nx_global) // nranks
    return Slab(rank, nranks, i0, i1 - i0, nx_global)


def exchange_descriptors(my_desc: bytes, slab: Slab, dist, device=None) -> dict:
    """All-gather the fixed-size peer descriptors; returns {-1: lower's bytes, +1: upper's bytes} (missing at walls).

    `dist` is torch.distributed (initialised).  Works with gloo (CPU tensors) and nccl (CUDA tensors)."""
    import torch
    n = len(my_desc)
    dev = device if device is not None else "cpu"
    mine = torch.frombuffer(bytearray(my_desc), dtype=torch.uint8).to(dev)
    gathered = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(slab.nranks)]
    dist.all_gather(gathered, mine)
    out = {}
    if slab.lower is not None:
        out[-1] = bytes(gathered[slab.lower].cpu().numpy().tobytes())
    if slab.upper is not None:
        out[+1] = bytes(gathered[slab.upper].cpu().numpy().tobytes())
    return out


class SlabSolver:
    """A slab of a global problem on this rank's GPU, connected to its neighbours.

    solver_factory(pb, mode, device, rank, nranks, nx_local) must return an object with the capi.Solver interface;
    the default is programs.make_solver (the CUDA library)."""

    def __init__(self, pb, mode, dist, rank: int, nranks: int, device: int, solver_factory=None, gather_device=None):
        from . import programs
        self.pb = pb
        self.dist = dist
        self.slab = partition(pb.nx, nranks, rank)
        # 1D: K steps per halo exchange round, agreed from the GLOBAL problem so that every rank uses the same value
        self.halo_steps = programs.halo_steps_for(pb, nranks, mode) if solver_factory is None else 1
        factory = solver_factory or (lambda pb_, mode_, dev_, r_, n_, nl_: programs.make_solver(
            pb_, mode_, device=dev_, rank=r_, nranks=n_, nx_local=nl_, halo_steps=self.halo_steps))
        self.solver = factory(pb, mode, device, rank, nranks, self.slab.nx_local)
        if nranks > 1:
            descs = exchange_descriptors(self.solver.peer_export(), self.slab, dist, gather_device)
            for side, d in descs.items():
                self.solver.peer_connect(side, d)
            dist.barrier()

    def initial_state(self) -> np.ndarray:
        """This slab's rows of the global initial condition, as conserved variables."""
        from . import programs
        p = programs.initial_primitives(self.pb, i0=self.slab.i0, nx_local=self.slab.nx_local, nx_global=self.pb.nx)
        return programs.cons_from_prim(self.pb, p)

    def upload(self, u_local: np.ndarray):
        # all ranks must be idle before halos of a new state are pushed into their buffers
        if self.slab.nranks > 1:
            self.dist.barrier()
        self.solver.upload_u(u_local)
        if self.slab.nranks > 1:
            self.dist.barrier()

    def close(self):
        if self.slab.nranks > 1:
            self.dist.barrier()
        self.solver.close()
