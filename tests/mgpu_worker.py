"""Worker for the multi-GPU tests: launched by torch.distributed.run, one rank per GPU (nccl) -- or per CPU process
(gloo, with a fake solver) for the host-logic test.

Every rank owns a slab of the global problem, runs the fused step kernel with in-kernel halo exchange, and rank 0
compares the gathered result bit for bit with (a) a single-GPU run of the same problem and (b) the CPU oracle.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from shll_sve_cfd_b200 import capi, programs, slabs

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gdev = torch.device("cuda", local)

    cases = [
        ("2d_o1", replace(programs.BASE_SHLL_2D.resized(64 * world + 3, 128), lx=(64 * world + 3) / 128), 37),
        ("2d_o2", replace(programs.SECOND_ORDER_2D.resized(48 * world, 128), lx=(48 * world) / 128), 33),
        ("2d_o2_thin", replace(programs.SECOND_ORDER_2D.resized(4 * world + 1, 64), lx=(4 * world + 1) / 64), 9),
        ("2d_o2_ldg", replace(programs.SECOND_ORDER_2D.resized(24 * world, 90), lx=(24 * world) / 90), 21),   # ny % 4 != 0 -> LDG kernel
        ("1d_o1", programs.BASE_SHLL.resized(1000 * world + 7), 45),
        ("1d_o2", programs.SECOND_ORDER_1D.resized(4096 * world), 45),
        ("1d_o2_tiny", programs.SECOND_ORDER_1D.resized(5 * world), 12),
    ]
    results = {}
    for name, pb, steps in cases:
        for mode in (capi.MODE_STRICT, capi.MODE_FAST):
            ss = slabs.SlabSolver(pb, mode, dist, rank, world, local, gather_device=gdev)
            u_loc = ss.initial_state()
            ss.upload(u_loc)
            ss.solver.run(steps)
            out = ss.solver.download_u()
            variant = ss.solver.variant
            ss.close()
            # gather slabs on rank 0
            ncomp = pb.ncomp
            row = pb.ny if pb.dims == 2 else 1
            parts = [None] * world
            dist.all_gather_object(parts, (ss.slab.i0, out))
            if rank == 0:
                full = np.concatenate([p[1].reshape(ncomp, -1, row) for p in sorted(parts, key=lambda t: t[0])], axis=1).reshape(ncomp, -1)
                u0 = programs.cons_from_prim(pb, programs.initial_primitives(pb))
                with programs.make_solver(pb, mode, device=local) as s1:
                    s1.upload_u(u0)
                    s1.run(steps)
                    single = s1.download_u()
                same_single = bool(np.array_equal(full.view(np.uint32), single.view(np.uint32)))
                same_oracle = None
                if mode == capi.MODE_STRICT:
                    from conftest import oracle_cfg_for
                    from oracle import oracle as O
                    ref = O.run(oracle_cfg_for(O, pb, nthreads=4), u0, steps)
                    same_oracle = bool(np.array_equal(full.view(np.uint32), ref.view(np.uint32)))
                results[f"{name}:{'strict' if mode == capi.MODE_STRICT else 'fast'}"] = dict(
                    same_as_single_gpu=same_single, same_as_oracle=same_oracle, variant=variant,
                    maxdiff=float(np.abs(full - single).max()))
            dist.barrier()
    if rank == 0:
        print("MGPU_RESULTS " + json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
