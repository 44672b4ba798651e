"""Node-sharded multi-GPU path on 2 GPUs: tools/check_sharded.py under torchrun (needs >= 2 GPUs; the single-GPU
driver run skips it, `gpurun --gpus 2` runs it)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("model", ["ours", "tgn", "jodie", "tgat"])
def test_sharded_equals_single_gpu_on_two_ranks(model):
    port = 29600 + os.getpid() % 300
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "check_sharded.py"), model],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert "== single GPU" in p.stdout, tail
    print(p.stdout.strip().splitlines()[-1])
