"""Multi-GPU correctness on the record: when the box has >= 2 GPUs, run
tests/multi/check_sharded.py under torchrun on up to 8 of them.  The script asserts that the
sharded device ensemble (NCCL all-gather transport, replicated-state transport with and
without NVSwitch multicast) and the sharded public sampler reproduce the single-GPU chain,
log-probabilities, blob records and acceptance counts BITWISE.
bench.py --gpus N repeats the same check before its timed region."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_chain_is_bitwise_equal_to_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU)")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi",
                                                            "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert "sharded == single-GPU chain (bitwise) on %d ranks: OK" % n in r.stdout, tail
    for line in ("transport nccl", "transport fused (multicast False)"):
        assert line in r.stdout, tail
