"""Data-parallel correctness over NCCL on two GPUs: the reduced gradients are the mean of the two ranks'
own gradients, with the ResNet-18 style encoder of BASELINE.json configs[2].  Skipped on a single-GPU box; the report of a
2-GPU run is committed under profiles/ (r02_dp_nccl_check.json)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("encoder", ["resnet18", "vgg"])
def test_reduced_gradients_are_the_mean_of_the_rank_gradients(encoder, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dp_nccl_worker.py"), str(out), encoder]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.loads(out.read_text())
    print("\n" + json.dumps(rep))
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, f"dp_nccl_check_{encoder}.json"), "w") as f:
            json.dump(rep, f, indent=1)
    assert rep["ok_all_ranks"], rep
    assert rep["subnets"]["gen"]["tensors_without_grad"] > 0       # dead tensors are skipped, not reduced
