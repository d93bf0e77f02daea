"""TEST INFRASTRUCTURE ONLY — CPU oracle for the FFT-convolution hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker / the thing the GPU path is compared
with.  The product (``cuda-fft-convolution_b200/``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: the reference (chrischoy/CUDA-FFT-Convolution) ships no tests,
golden vectors or known-answer fixtures, and its arithmetic lives in closed-source
cuFFT (CUDA 6.0, ``compile.m:2``; call sites ``src/cudaConvolutionFFT.cu:128-142,
167,255,273``).  The oracle is therefore pinned by
  (1) a float64 direct convolution (``demoCudaConvolutionFFT.m:91-96``),
  (2) a numpy restatement of the reference pipeline, step for step,
  (3) golden outputs of the reference's OWN device kernels + cuFFT 11.4 replayed on
      a B200 (``oracle/ref_replay.cu`` -> ``oracle/_ref/``; fixtures committed under
      ``tests/golden/`` by ``tests/golden/make_golden.py``).
"""
from .oracle import *  # noqa: F401,F403
