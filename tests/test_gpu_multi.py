"""GPU (>= 2 B200s): one map sharded by block hash over the GPUs — frames pulled out of rank 0's ring by the copy engines (or NCCL
broadcast), rays split across the GPUs with keys routed to their owners' inboxes, peer-memory marching-cubes halos, mesh gather — in six
modes (tests/multi_gpu_worker.py); the union must equal the single-GPU engine bit for bit. Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from test_multi_host import free_port

pytestmark = pytest.mark.gpu


def run_world(n):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_map_equals_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = run_world(world)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-5000:]
