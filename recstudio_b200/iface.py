"""The reference's plugin interfaces for the retriever hot path.

If ``recstudio`` (ustcml/RecStudio) is importable, the REAL base classes are used so
that every replacement passes the reference's own ``isinstance`` gates
(``baseretriever.py:17-46``, ``recommender.py:48-54``, ``init.py:6,24,37``).  When it
is not (the GPU box: the Python reference cannot travel), local mirrors with the
same names, signatures and semantics are defined -- interface declarations only
(no arithmetic), so that the plugin layer and its tests read the same either way.

NB (import-order hazard in the reference): ``recstudio.model`` must be imported
before ``recstudio.ann.sampler`` -- sampler.py:6 imports ``recstudio.model.scorer``
whose package ``__init__`` reaches baseretriever.py:9 (``from recstudio.ann.sampler
import *``) while sampler.py is half-initialised.
"""
from __future__ import annotations

import torch

HAVE_RECSTUDIO = False
try:  # pragma: no cover - depends on the environment
    import recstudio.model  # noqa: F401  (must come first, see above)
    from recstudio.ann.sampler import MaskedUniformSampler as RefMaskedUniformSampler
    from recstudio.ann.sampler import PopularSamplerModel as RefPopularSamplerModel
    from recstudio.ann.sampler import Sampler
    from recstudio.ann.sampler import UniformSampler as RefUniformSampler
    from recstudio.model.basemodel import BaseRetriever
    from recstudio.model.loss_func import BPRLoss as RefBPRLoss
    from recstudio.model.loss_func import FullScoreLoss, PairwiseLoss, PointwiseLoss
    from recstudio.model.loss_func import SampledSoftmaxLoss as RefSampledSoftmaxLoss
    from recstudio.model.loss_func import SoftmaxLoss as RefSoftmaxLoss
    from recstudio.model.scorer import CosineScorer, EuclideanScorer, InnerProductScorer
    HAVE_RECSTUDIO = True
except Exception:  # recstudio (or one of its hard deps: nni, torchmetrics) is absent
    class Sampler(torch.nn.Module):
        """recstudio/ann/sampler.py:48-58"""

        def __init__(self, num_items, scorer_fn=None):
            super().__init__()
            self.num_items = num_items - 1      # remove padding (sampler.py:51)
            self.scorer = scorer_fn

        def update(self, item_embs, max_iter=30):
            pass

        def compute_item_p(self, query, pos_items):
            pass

    class FullScoreLoss(torch.nn.Module):
        """recstudio/model/loss_func.py:6-17"""

        def forward(self, label, pos_score, all_score):
            pass

    class PairwiseLoss(torch.nn.Module):
        """recstudio/model/loss_func.py:20-22"""

        def forward(self, label, pos_score, log_pos_prob, neg_score, log_neg_prob):
            pass

    class PointwiseLoss(torch.nn.Module):
        """recstudio/model/loss_func.py:25-28"""

        def forward(self, label, pos_score):
            raise NotImplementedError

    class InnerProductScorer(torch.nn.Module):
        """recstudio/model/scorer.py:5-17 (interface only; arithmetic lives in plugins.py)"""

        def forward(self, query, items):
            raise NotImplementedError

    class EuclideanScorer(InnerProductScorer):
        """recstudio/model/scorer.py:28-34"""

    class CosineScorer(InnerProductScorer):
        """recstudio/model/scorer.py:19-25 (marker: the samplers only test isinstance to normalise their inputs)"""

    # marker types so that `type(x) in (...)` checks read the same in both environments
    class RefUniformSampler(Sampler):
        pass

    class RefPopularSamplerModel(Sampler):
        pass

    class RefMaskedUniformSampler(Sampler):
        pass

    class RefBPRLoss(PairwiseLoss):
        pass

    class RefSampledSoftmaxLoss(PairwiseLoss):
        pass

    class RefSoftmaxLoss(FullScoreLoss):
        pass

    BaseRetriever = None
