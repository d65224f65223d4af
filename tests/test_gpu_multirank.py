"""N-GPU parity (needs >= 2 visible B200s; skipped otherwise): torchrun launches
tests/mrank_worker.py, one process per GPU, NCCL halo exchange inside the library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_multirank_equals_oracle(n, transport):
    """transport: "peer" = CUDA IPC peer memory over NVLink with the exchanges fused into the five PCG kernels and the
    iteration replayed as a CUDA graph (the default); "nccl" = PF_HALO=nccl, send/recv + all-gather.  Same bits."""
    if ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    env = dict(os.environ)
    if transport == "nccl":
        env["PF_HALO"] = "nccl"
    else:
        env.pop("PF_HALO", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + n + (20 if transport == "nccl" else 0)), os.path.join(ROOT, "tests", "mrank_worker.py"),
           "hex20", "hex20_thin", "hex8", "p123", "p123_fixed",
           "hex20:sym", "hex8:sym", "p123:sym", "hex20_thin:mf2", "hex20:mf1", "hex8:mf2",
           "hex20_psize", "hex20_psize:sym", "p124", "p124_fixed", "p125", "hex20_mat",
           "hex20_shuffled", "hex20_shuffled:sym", "hex20_shuffled:mf2", "p123_fixed_shuffled", "tet4", "tet4_scalar",
           "p129", "p122", "p1210", "p1210:mf"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0 and "MRANK_OK" in res.stdout, res.stdout[-4000:] + res.stderr[-4000:]
