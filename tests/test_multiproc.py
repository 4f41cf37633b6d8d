"""Multi-process paths: world_size-2 gloo test on CPU (host-side rendezvous + sharding logic) and,
on a box with >= 2 GPUs, the sharded-covariance parity run (one rank per GPU over NCCL)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, extra, timeout=900, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                          env=dict(os.environ, **(env or {})))


def test_gloo_world2_host_logic():
    out = _torchrun(2, ["--cpu"])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("cpu checks ok") == 2


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ekf_parity(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _torchrun(world, ["--what", "ekf"])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    assert out.stdout.count("sharded EKF parity ok") == world


@pytest.mark.gpu
def test_sharded_ekf_parity_nccl_exchange_fallback():
    """Without the CUDA-IPC peer mapping (CSLAM_EKF_PEER=0) the column exchange is a packed snapshot + NCCL
    all-reduce; the fused group-gain kernel then reads plain columns instead of flagged cells.  And the
    one-kernel-per-observation chain (CSLAM_GAIN_FUSED=0) over the flag-based peer push."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    for env in ({"CSLAM_EKF_PEER": "0"}, {"CSLAM_GAIN_FUSED": "0"}):
        out = _torchrun(2, ["--what", "ekf"], env=env)
        assert out.returncode == 0, str(env) + out.stdout[-3000:] + out.stderr[-6000:]
        assert out.stdout.count("sharded EKF parity ok") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_pf_parity(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _torchrun(world, ["--what", "pf"])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    assert out.stdout.count("sharded PF parity ok") == world
