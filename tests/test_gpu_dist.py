"""One process per GPU over NCCL (torchrun): the sharded index build + probe-sharded search give the families of a lone
build, bit for bit, and the same digest on every rank. Skipped on boxes with fewer than two GPUs (the same code runs
with host threads as members on one GPU in test_gpu_parity.py::test_sharded_index_*)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("msd_min", ["", "0"])
def test_two_ranks_equal_a_lone_build(msd_min):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    if msd_min:
        env["ASGART_B200_MSD_MIN"] = msd_min      # force the MSD form of the initial sort on the small test genome
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_index_check.py"), "2", "4000000"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DIST INDEX CHECK: OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    digests = {line.split()[-1] for line in r.stdout.splitlines() if "families digest" in line}
    assert len(digests) == 1
