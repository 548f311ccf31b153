"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU (block-hash sharded) map."""
import os
import socket
import subprocess
import sys

from conftest import ROOT


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_partition_and_mesh_merge_world2(vh):
    vh.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "multi_host_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "MULTI_HOST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
