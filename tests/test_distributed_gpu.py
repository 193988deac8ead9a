"""-m gpu: the data-parallel product path (TrainStep + NCCL all-reduce inside the captured step).

With one visible GPU the script runs as a world of one (replay == product path bit for bit, which pins
`local_gradient` + `_enqueue_update` to the graphed step); with two or more it is launched under torchrun on
two GPUs and checks replica identity and the two-gradient single-GPU replay (tests/dist_trainstep_check.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "dist_trainstep_check.py")


def _run(cmd, extra_env=None):
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    env.update(extra_env or {})
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    return json.loads(lines[-1])


def test_trainstep_world_of_one_replay_is_bit_identical():
    rep = _run([sys.executable, SCRIPT], {"GCB_DP_BATCH": "512", "GCB_DP_STEPS": "4", "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
    assert rep["ok"] and rep["replay_bit_identical"] and rep["moments_bit_identical"], rep


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run through `gpurun --gpus 2`)")
def test_trainstep_two_gpus_nccl():
    rep = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29533", SCRIPT])
    assert rep["ok"] and rep["replicas_bit_identical"] and rep["replay_bit_identical"], rep
