#!/usr/bin/env python
"""BASELINE configs[0] end to end: BPR on ml-100k under the REFERENCE'S OWN trainer
(TripletDataset -> fit -> evaluate, recstudio/quickstart/run.py:6-61), once with the reference's BPR
and once with FusedBPR (same seed, same CUDA generator stream => the same negatives), both on cuda:0.
Needs the unmodified reference importable from baseline/_ref (pip --target install, git-ignored) and the
two import-only stand-ins under baseline/shim.  Prints one JSON line."""
import json
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)     # recstudio_b200.iface resolves the reference (baseline/_ref) and the nni / torchmetrics stand-ins
os.chdir(tempfile.mkdtemp(prefix="rs_ml100k_"))

import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import logging  # noqa: E402

import torch  # noqa: E402

from recstudio_b200 import iface  # noqa: E402,F401  (puts baseline/_ref on sys.path and imports recstudio.model first)
from recstudio.data.dataset import TripletDataset  # noqa: E402
from recstudio.model.mf.bpr import BPR  # noqa: E402
from recstudio.utils import get_model  # noqa: E402

from recstudio_b200.retriever import FusedBPR  # noqa: E402

EPOCHS = int(os.environ.get("RSB_EPOCHS", "8"))


def run(kind, grad_mode="dense", learner="adam", sampling_method="none", negative_count=1, device_loader=False):
    conf = get_model("BPR")[1]
    conf["train"].update({"gpu": [0], "epochs": EPOCHS, "seed": 2022, "learner": learner, "early_stop_patience": 100,
                          "sampling_method": sampling_method, "negative_count": negative_count})
    data_conf = {"user_feat_name": None}                 # pandas-3 CoW workaround; BPR uses ids only
    data_conf.update(conf["data"])
    model = BPR(conf) if kind == "reference" else FusedBPR(conf, fused_grad=grad_mode, device_loader=device_loader)
    datasets = TripletDataset(name="ml-100k", config=data_conf).build(**conf["data"])
    logging.getLogger("recstudio").setLevel(logging.ERROR)
    val = model.fit(*datasets[:2], run_mode="light")
    test = model.evaluate(datasets[-1])
    fused_steps = getattr(model, "_fused_steps", None)
    out = {"val": {k: float(v) for k, v in (val or {}).items()}, "test": {k: float(v) for k, v in test.items()},
           "item_norm": float(model.item_encoder.weight.norm()), "user_norm": float(model.query_encoder.weight.norm()),
           "encoder": type(model.item_encoder).__name__, "sampler": type(model.sampler).__name__,
           "loss": type(model.loss_fn).__name__, "fused_ws": bool(getattr(model, "_fused_ws_cache", None))}
    return out


if __name__ == "__main__":
    res = {"reference": run("reference"), "fused_dense": run("fused", "dense"),
           "fused_sparse": run("fused", "sparse", learner="sparse_adam"),
           "fused_device_loader": run("fused", "dense", device_loader=True),      # 8(f)-3: batches sliced on the GPU
           # sampling_method 'dns' (baseretriever.py:313-343) is outside the fused combination: the reference's own
           # forward / sampling code runs on top of the standalone CUDA plugins (gather, scorer, loss)
           "reference_dns": run("reference", sampling_method="dns", negative_count=[8, 2]),
           "fused_dns": run("fused", "dense", sampling_method="dns", negative_count=[8, 2]),
           "epochs": EPOCHS, "torch": torch.__version__}
    print("RESULT " + json.dumps(res))
