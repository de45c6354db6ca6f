"""nbss_b200 — B200-native (sm_100a) SpatialNet hot path behind the NBSS module surface.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI library), ``_lib`` (ctypes loader) and
the host-side mirrors of the reference interface (``SpatialNet``, ``STFT``, ``Norm``).
"""
__version__ = "0.1.0"
