"""CPU (gloo, world_size 2 and 3): the host side of the slab decomposition -- partition, slab initial conditions,
descriptor exchange and neighbour wiring -- with a fake solver in place of the CUDA library."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from shll_sve_cfd_b200 import programs, slabs


def test_partition_is_balanced_and_contiguous():
    for nx in (7, 64, 4096, 16384 + 5):
        for n in (1, 2, 3, 8):
            parts = [slabs.partition(nx, n, r) for r in range(n)]
            assert parts[0].i0 == 0 and parts[-1].i0 + parts[-1].nx_local == nx
            for a, b in zip(parts, parts[1:]):
                assert a.i0 + a.nx_local == b.i0
            sizes = [p.nx_local for p in parts]
            assert max(sizes) - min(sizes) <= 1
            assert parts[0].lower is None and parts[-1].upper is None
            if n > 1:
                assert parts[0].upper == 1 and parts[-1].lower == n - 2
    with pytest.raises(ValueError):
        slabs.partition(10, 2, 2)


WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    from shll_sve_cfd_b200 import capi, programs, slabs

    class FakeSolver:
        """Stands in for capi.Solver: records the wiring, 'exports' a descriptor that encodes its rank."""
        def __init__(self, pb, mode, dev, rank, nranks, nx_local):
            self.rank, self.nranks, self.nx_local, self.connected, self.uploaded = rank, nranks, nx_local, {{}}, None
        def peer_export(self):
            d = capi.PeerDesc(); d.pid = 1000 + self.rank; d.nx = self.nx_local; d.device = self.rank
            return bytes(d)
        def peer_connect(self, side, desc):
            d = capi.PeerDesc.from_buffer_copy(desc); self.connected[side] = (int(d.pid) - 1000, int(d.nx))
        def upload_u(self, u): self.uploaded = u.shape
        def close(self): pass

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    pb = programs.SECOND_ORDER_2D.resized(50, 24)
    ss = slabs.SlabSolver(pb, capi.MODE_STRICT, dist, rank, world, 0, solver_factory=lambda *a: FakeSolver(*a))
    u = ss.initial_state()
    ss.upload(u)
    full = programs.cons_from_prim(pb, programs.initial_primitives(pb)).reshape(4, 50, 24)
    ok_ic = bool(np.array_equal(u.reshape(4, -1, 24).view(np.uint32), full[:, ss.slab.i0:ss.slab.i0 + ss.slab.nx_local].view(np.uint32)))
    out = dict(rank=rank, i0=ss.slab.i0, nx=ss.slab.nx_local, connected={{str(k): v for k, v in ss.solver.connected.items()}},
               uploaded=list(ss.solver.uploaded), ok_ic=ok_ic)
    ss.close()
    allr = [None] * world
    dist.all_gather_object(allr, out)
    if rank == 0:
        print("SLAB_RESULTS " + json.dumps(allr))
    dist.destroy_process_group()
''')


@pytest.mark.parametrize("world", [2, 3])
def test_descriptor_exchange_and_wiring_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + world), str(script)]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert pr.returncode == 0, pr.stdout[-2000:] + pr.stderr[-2000:]
    import json
    line = [l for l in pr.stdout.splitlines() if l.startswith("SLAB_RESULTS ")][0]
    res = sorted(json.loads(line[len("SLAB_RESULTS "):]), key=lambda r: r["rank"])
    assert sum(r["nx"] for r in res) == 50
    for r in res:
        assert r["ok_ic"] and r["uploaded"] == [4, r["nx"] * 24]
        want = {}
        if r["rank"] > 0:
            want["-1"] = [r["rank"] - 1, res[r["rank"] - 1]["nx"]]
        if r["rank"] < world - 1:
            want["1"] = [r["rank"] + 1, res[r["rank"] + 1]["nx"]]
        assert r["connected"] == want
