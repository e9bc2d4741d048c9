"""recstudio_b200 -- B200 (sm_100a) implementation of RecStudio's retriever training-step
hot path behind the reference's own plugin surfaces.

Layout
  csrc/          hand-written CUDA kernels + the C ABI (include/rsb200.h) -> librsb200.so
  _lib.py        ctypes binding (no torch types cross the boundary)
  fused.py       host side of the fused gather-score-loss-scatter step
  sampling.py    UniformSampler / PopularSamplerModel draws (torch-CUDA-identical Philox stream)
  plugins.py     drop-in Sampler / scorer / loss_func / nn.Embedding / BaseRetriever subclasses

There is no CPU fallback: every op raises if librsb200.so or a CUDA device is missing.
"""
__version__ = "0.1.0"
