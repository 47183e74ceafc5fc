"""protosam_b200 -- B200-native coarse-segmentation hot path of ProtoSAM.

Only what the path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI
in include/psam_b200.h), a ctypes loader, and the host-side mirrors of the
reference interfaces (``MultiProtoAsConv``, the prompt helpers, the batched
volume engine).  There is no CPU fallback: every entry point raises if the CUDA
library is missing or the inputs are not CUDA tensors.
"""
__version__ = "0.1.0"
