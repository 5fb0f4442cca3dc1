"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`; skipped on one):
W processes, one GPU each, the engine's own NCCL communicator.  Every mode must reproduce the single-process run on the
same GLOBAL batch: data-parallel coordinate steps (one fused all-reduce per step), data-parallel momentum-space steps
(all-reduce of the kernel-space block per iteration) and the frequency-bin sharded momentum-space step (data-parallel
forward, all-to-all of the spectrum slabs, bin-sharded training with the multiobjective term split over the ranks)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_np as O
from conftest import ROOT

pytestmark = pytest.mark.gpu
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def _run(mode, world, tmp_path):
    outs = [str(tmp_path / f"{mode}_{world}_{r}.npz") for r in range(world)]
    procs = [subprocess.Popen([sys.executable, WORKER, mode, str(r), str(world), str(tmp_path / f"id_{mode}_{world}"), outs[r]],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    for p in procs:
        so, se = p.communicate(timeout=600)
        assert p.returncode == 0, se[-2000:]
    return [dict(np.load(o)) for o in outs]


@pytest.mark.parametrize("mode", ["coord", "fft_dp", "fft_bins"])
def test_two_ranks_equal_one_rank_on_the_global_batch(mode, tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    one = _run(mode, 1, tmp_path)[0]
    two = _run(mode, 2, tmp_path)
    for k in one:
        if k == "mse":
            continue
        assert np.array_equal(two[0][k], two[1][k]), k            # replicas stay identical without any broadcast
        assert O.rel_l2(two[0][k], one[k]) < 2e-5, (mode, k, O.rel_l2(two[0][k], one[k]))
    assert np.allclose(two[0]["mse"], one["mse"], rtol=1e-4), (two[0]["mse"], one["mse"])
