"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): torchrun launches tests/mg_worker.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("split", ["0", "1"], ids=["plain_sequence", "exchange_hidden_behind_interior"])
def test_strip_partitioned_gpus_equal_single_gpu(split):
    """split = PFEM2_P2P_SPLIT: the A/B form of the P2P step that moves / reduces the boundary layers first."""
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tests", "mg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, PFEM2_P2P_SPLIT=split))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MG_OK") == 4 and r.stdout.count("MG_SPILL_OK") == 2, r.stdout[-2000:]
