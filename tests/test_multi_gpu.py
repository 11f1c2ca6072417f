"""GPU (-m gpu, needs >= 2 GPUs): slab decomposition with the in-kernel NVLink halo exchange.

The N-GPU result must be bitwise equal to the 1-GPU result in BOTH arithmetic modes (the update has no reductions, and
every cell goes through the same device function whatever tile / slab it sits in), and equal to the oracle in STRICT mode."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_bitwise_equal_to_single_gpu(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", SHLL_HALO_TIMEOUT_MS="4000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert pr.returncode == 0, pr.stdout[-3000:] + pr.stderr[-3000:]
    line = [l for l in pr.stdout.splitlines() if l.startswith("MGPU_RESULTS ")]
    assert line, pr.stdout[-2000:]
    res = json.loads(line[0][len("MGPU_RESULTS "):])
    assert len(res) == 14
    for name, r in res.items():
        assert r["same_as_single_gpu"], f"{name}: {world}-GPU result differs from 1-GPU ({r})"
        if name.endswith(":strict"):
            assert r["same_as_oracle"], f"{name}: differs from the oracle ({r})"
