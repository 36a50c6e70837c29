"""Multi-process coverage of the N > 1 path: world_size-2 gloo on CPU for the host logic,
and (on a box with >= 2 GPUs) NCCL for the bit-identity of a sharded chain."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_mp_shard_worker.py")


def _torchrun(nproc, mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_gloo_world2_host_logic():
    out = _torchrun(2, "cpu", 29511)
    assert "CPU_SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_gloo_world4_two_chains_of_two_shards():
    out = _torchrun(4, "groups", 29513)
    assert "CPU_GROUPS_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.gpu
def test_nccl_sharded_chain_bit_identical():
    import instruct_b200
    if instruct_b200.load().ig_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _torchrun(2, "gpu", 29512)
    assert "GPU_SHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
