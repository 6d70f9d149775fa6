"""z-slab solves over two processes / two GPUs (NCCL) against the single-GPU solve of the same volume.
Skipped on a one-GPU box; the host-side partitioning logic is covered on CPU by test_dist_cpu.py (gloo)."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["peer-memory", "nccl"])
@pytest.mark.parametrize("n", [96, 128])
def test_two_rank_slab_solve_matches_single_gpu(n, transport):
    """One process per GPU (torchrun): halos and reductions through peer memory (CUDA IPC + NVLink stores issued by
    the solver's kernels) and, as the cross-check, through NCCL."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + n % 50), os.path.join(ROOT, "tests", "helpers", "slab_check.py"), str(n)]
    env = dict(os.environ)
    env.pop("KEFFB200_SLAB_NCCL", None)
    if transport == "nccl":
        env["KEFFB200_SLAB_NCCL"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"transport: {transport}" in r.stdout, r.stdout[-2000:]
    slab = re.search(r"'keff': ([0-9.eE+-]+).*'iters': (\d+), 'status': (\d+)", r.stdout)
    single = re.search(r"keff rel diff ([0-9.eE+-]+) \| T maxdiff ([0-9.eE+-]+)", r.stdout)
    assert slab and single, r.stdout[-2000:]
    assert int(slab.group(3)) == 0 and int(slab.group(2)) < 40          # converged, multigrid iteration counts
    assert float(single.group(1)) < 1e-7 and float(single.group(2)) < 1e-6


def test_one_process_drives_all_gpus():
    """keffb200_init_multi: the host-buffer batch call shards its images over the GPUs, the host-buffer volume call
    cuts the volume into z-slabs linked through peer memory; both must reproduce the one-GPU results."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "multi_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "multi ok" in r.stdout, r.stdout[-2000:]
