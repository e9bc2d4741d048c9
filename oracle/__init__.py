"""CPU oracle for the RecStudio retriever hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's hot path
(``recstudio.ann.sampler`` / ``recstudio.model.scorer`` /
``recstudio.model.loss_func`` / ``BaseRetriever.forward|training_step|topk`` /
``recstudio.eval``).  Every function cites the reference file:line it follows.

Nothing in the product (``recstudio_b200/``) may import it.  The only legal
importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or
as the CPU arm being timed, never as the thing shipped.

Pinning status
--------------
* loss / score / gradient / top-k / metric functions are pinned against golden
  vectors produced by importing the *unmodified* reference from
  ``/root/reference`` (``tests/golden/make_golden.py``; fixtures committed under
  ``tests/golden/``) and against SURVEY.md Appendix A.
* the Philox stream (S1/S2) restates a third-party dependency that is not in
  ``/root/reference``: PyTorch/ATen's CUDA generator (reference pin
  ``pytorch=1.12.1`` in environment.yml:97; this image: torch 2.11.0+cu128).
  It is pinned (a) against the Random123 known-answer vectors for
  Philox4x32-10 and (b), on a GPU box, live against ``torch.randint`` /
  ``torch.rand`` on CUDA (tests/test_gpu_sampler.py).  There is no GPU in the
  authoring container, so (b) is only ever checked by the ``-m gpu`` suite.
"""
