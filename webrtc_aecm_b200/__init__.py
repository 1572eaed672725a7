"""webrtc_aecm_b200 -- B200-native batched WebRTC AECM engine.

The product is the C-ABI shared library ``libaecm_b200.so`` (CUDA kernels for sm_100a + a plain-C host
layer, sources under ``csrc/``, headers under ``include/``).  This package is only a thin ctypes mirror of
that ABI for tests, tooling and ``bench.py``; it contains no arithmetic and no CPU fallback: importing
:mod:`webrtc_aecm_b200.api` without the built library raises.
"""
from .api import (AECM_BAD_PARAMETER_ERROR, AECM_BAD_PARAMETER_WARNING, AECM_NULL_POINTER_ERROR,  # noqa: F401
                  AECM_UNINITIALIZED_ERROR, AECM_UNSPECIFIED_ERROR, AecmBatch, AecmConfig, InstanceState, lib,
                  library_path, scalar_api)
