"""BASELINE configs[0]: BPR on ml-100k through the reference's own trainer, reference BPR vs FusedBPR on
the same device with the same seed.  Because the fused sampler reproduces torch.randint's CUDA stream
and the data loader / initialisation use the same CPU generator, the two runs see the same batches and
the same negatives: metrics must agree to fp32 noise.  Skipped when the reference is not installed under
baseline/_ref (it is git-ignored; install recipe in DESIGN.md section 11)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def test_bpr_ml100k_reference_trainer_drop_in():
    if not os.path.isdir(os.path.join(REPO, "baseline", "_ref", "recstudio")):
        pytest.skip("reference not installed under baseline/_ref")
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "integration", "run_ml100k.py")],
                       capture_output=True, text=True, timeout=1500)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and line, (r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads(line[-1][7:])
    try:                                           # keep the numbers as evidence (gpurun merges this directory back)
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(REPO, "gpurun_out", "ml100k_dropin.json"), "w"), indent=1)
    except OSError:
        pass
    ref, fd, fs = res["reference"], res["fused_dense"], res["fused_sparse"]
    assert fd["encoder"] == "FusedEmbedding" and fd["sampler"] == "FusedUniformSampler" and fd["fused_ws"], fd
    assert ref["encoder"] == "Embedding" and not ref["fused_ws"]
    # same batches + same negatives + dense Adam on dense grads  =>  same model up to fp32 noise
    assert abs(fd["item_norm"] - ref["item_norm"]) <= 2e-3 * ref["item_norm"], (fd, ref)
    for k, v in ref["test"].items():
        assert abs(fd["test"][k] - v) <= 0.01, (k, fd["test"][k], v)
    # device-resident batch producer: same epoch permutations => same model
    fl = res["fused_device_loader"]
    assert abs(fl["item_norm"] - ref["item_norm"]) <= 2e-3 * ref["item_norm"], (fl, ref)
    for k, v in ref["test"].items():
        assert abs(fl["test"][k] - v) <= 0.01, (k, fl["test"][k], v)
    # a sampling method the kernels do not fuse: reference code path over the standalone plugins, same stream of
    # negatives => same model up to the summation-order noise of the standalone scorer kernels
    rd, fdn = res["reference_dns"], res["fused_dns"]
    assert abs(fdn["item_norm"] - rd["item_norm"]) <= 5e-3 * rd["item_norm"], (fdn, rd)
    for k, v in rd["test"].items():
        assert abs(fdn["test"][k] - v) <= 0.02, (k, fdn["test"][k], v)
    # sparse gradients + SparseAdam is a different optimizer (moments of untouched rows do not decay):
    # it must train to a comparable quality, not to identical numbers
    assert fs["test"]["ndcg@10"] > 0.3 * ref["test"]["ndcg@10"], (fs["test"], ref["test"])
