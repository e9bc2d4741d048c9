"""The reference's plugin interfaces for the retriever hot path: the REAL ``recstudio`` classes.

Every replacement in this package subclasses the reference class it replaces so that it passes the
reference's own ``isinstance`` gates (``baseretriever.py:17-46``, ``recommender.py:48-54``,
``init.py:6,24,37``).  There are no local mirrors: ``recstudio`` (ustcml/RecStudio, unmodified) must be
importable.  It is resolved in this order:

  1. whatever ``import recstudio`` finds (a user's own installation);
  2. ``<repo>/baseline/_ref`` -- the ``pip install --target`` copy that ``__graft_entry__.build()`` makes
     from ``/root/reference`` (git-ignored, travels to the GPU box with the snapshot).

The reference hard-imports two packages this image lacks and never calls on this path, ``nni``
(``recommender.py:10``, ``utils.py:8``) and ``torchmetrics`` (``eval/__init__.py:6``); when they are
missing, the import-only stand-ins under ``<repo>/baseline/shim`` are appended to ``sys.path``.

NB (import-order hazard in the reference): ``recstudio.model`` must be imported before
``recstudio.ann.sampler`` -- sampler.py:6 imports ``recstudio.model.scorer`` whose package
``__init__`` reaches baseretriever.py:9 (``from recstudio.ann.sampler import *``) while sampler.py is
half-initialised.  Importing ``recstudio.utils`` creates ``./log`` and ``./.recstudio`` in the CWD
(``utils.py:27-31``).
"""
from __future__ import annotations

import importlib.util
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(_REPO, "baseline", "_ref")
SHIM_DIR = os.path.join(_REPO, "baseline", "shim")


def _have(mod: str) -> bool:
    try:
        return importlib.util.find_spec(mod) is not None
    except (ImportError, ValueError):
        return False


def _resolve():
    if not _have("recstudio") and os.path.isdir(os.path.join(REF_DIR, "recstudio")):
        sys.path.append(REF_DIR)
    if not _have("recstudio"):
        raise ImportError(
            "recstudio_b200 plugs into ustcml/RecStudio, which is not importable.  Install the unmodified reference "
            "(`python -c 'import __graft_entry__ as g; g.build()'` puts it under %s) or add your own checkout to "
            "sys.path.  The low-level ops (recstudio_b200.fused / sampling / topk / sharded) do not need it." % REF_DIR)
    if not (_have("nni") and _have("torchmetrics")):
        sys.path.append(SHIM_DIR)           # appended: a real installation always wins


_resolve()

import recstudio.model  # noqa: E402,F401  (must come first, see above)
from recstudio.ann.sampler import MaskedUniformSampler as RefMaskedUniformSampler  # noqa: E402,F401
from recstudio.ann.sampler import PopularSamplerModel as RefPopularSamplerModel  # noqa: E402,F401
from recstudio.ann.sampler import Sampler  # noqa: E402,F401
from recstudio.ann.sampler import UniformSampler as RefUniformSampler  # noqa: E402,F401
from recstudio.model.basemodel import BaseRetriever  # noqa: E402,F401
from recstudio.model.loss_func import BPRLoss as RefBPRLoss  # noqa: E402,F401
from recstudio.model.loss_func import FullScoreLoss, PairwiseLoss, PointwiseLoss  # noqa: E402,F401
from recstudio.model.loss_func import SampledSoftmaxLoss as RefSampledSoftmaxLoss  # noqa: E402,F401
from recstudio.model.loss_func import SoftmaxLoss as RefSoftmaxLoss  # noqa: E402,F401
from recstudio.model.scorer import CosineScorer, EuclideanScorer, InnerProductScorer  # noqa: E402,F401

HAVE_RECSTUDIO = True      # kept for callers that used to branch on it; always true now
