"""Multi-GPU parity ON HARDWARE: ShardedStore at world size 2 and 4, one process per GPU (torchrun over NCCL / NVLink),
both exchange forms, against the CPU oracle over the unsharded corpus (tests/_sharded_gpu_worker.py).

Skipped when fewer GPUs are visible than the world size (the driver's 1-GPU tier); run with
`gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu -q` -- the log of such a run is kept under
profiles/.  The N > 1 host protocol is also covered on CPU (gloo) by tests/test_sharded_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus() -> int:
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_search_matches_the_oracle_on_hardware(world, exchange):
    if _gpus() < world:
        pytest.skip(f"{world} GPUs needed, {_gpus()} visible")
    env = dict(os.environ, MX_EXCHANGE=exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "_sharded_gpu_worker.py"), "--epochs", "6"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=800)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    oks = [ln for ln in r.stdout.splitlines() if " ok: " in ln]
    assert len(oks) == world, r.stdout
    assert all(f"exchange {exchange}" in ln for ln in oks), oks   # the form that was asked for is the one that ran
