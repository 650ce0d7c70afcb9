"""GPU, >= 2 devices: map sharded over ranks + all-reduce of the normal equations == unsharded result, for both exchange
paths (NVLink peer memory inside the LM kernel; ncclAllReduce between per-evaluation kernels).  The rank logs are kept as an
artefact under gpurun_out/ (and copied to profiles/ by the builder) so that the run can be audited."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("peer", [True, False])
def test_sharded_mapping_matches_unsharded(peer):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ)
    if not peer:
        env["ALOAM_NO_PEER"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29611" if peer else "29612", os.path.join(ROOT, "tests", "multi_gpu_mapping_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    sys.stdout.write(r.stdout[-3000:]); sys.stderr.write(r.stderr[-3000:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "multi_gpu_mapping_%dgpu_%s.log" % (world, "peer" if peer else "nccl")), "w") as f:
        f.write(r.stdout)
    assert r.returncode == 0 and "MULTI_GPU_MAPPING_OK" in r.stdout
    assert ("peer memory" in r.stdout) == peer
