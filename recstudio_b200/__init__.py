"""recstudio_b200 -- B200 (sm_100a) implementation of RecStudio's retriever training-step
hot path behind the reference's own plugin surfaces.

Layout
  csrc/          hand-written CUDA kernels + the C ABI (include/rsb200.h) -> librsb200.so
  _lib.py        ctypes binding (no torch types cross the boundary)
  fused.py       host side of the fused gather-score-loss-scatter step (PairWorkspace, pair_step, GraphedPairStep)
  sampling.py    UniformSampler / PopularSamplerModel / MaskedUniformSampler draws (torch-CUDA-identical Philox stream)
  plugins.py     drop-in Sampler / scorer / loss_func / nn.Embedding plugins, FusedRetrieverMixin (fused training_step,
                 sampling() methods, top-k)
  retriever.py   FusedRetriever / FusedBPR (BaseRetriever subclasses), MiniRetriever mirror, synthetic builders
  midx.py        k-means, construct_index and the MIDX / Cluster samplers (index build on csrc/midx.cu)
  attention.py   FusedSASRecQueryEncoder (tcgen05 attention block), sequence pooling
  topk.py        full-catalog top-k;  rank_metrics.py  the reference's rank metrics
  rowopt.py      FusedRowOptimizer (touched-row SGD / Adagrad / SparseAdam, also fused into the scatter epilogue)
  sharded.py     row-sharded item table: row lookup and owner-compute steps over torch.distributed
  loader.py      device-resident batch loader

There is no CPU fallback: every op raises if librsb200.so or a CUDA device is missing.
"""
__version__ = "0.1.0"
