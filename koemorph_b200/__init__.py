"""koemorph_b200: B200-native (sm_100a) audio -> ARKit-blendshape inference path of KoeMorph.

Mirrors the reference's ``src/model`` and ``src/features`` module surface for the dual-stream path
(SURVEY.md section 8b).  All compute is in hand-written CUDA kernels behind the C ABI declared in
``include/koemorph_b200.h``; PyTorch supplies device memory, streams and ``torch.distributed`` only.
"""
from .model.dual_stream_attention import (ARKIT_BLENDSHAPES, EXPRESSION_INDICES, MOUTH_BLENDSHAPES, MOUTH_INDICES,
                                          DualStreamCrossAttention)
from .model.sequential_dual_stream_model import SequentialDualStreamModel
from .model.simplified_dual_stream_model import SimplifiedDualStreamModel

__all__ = ["DualStreamCrossAttention", "SimplifiedDualStreamModel", "SequentialDualStreamModel",
           "MOUTH_INDICES", "EXPRESSION_INDICES", "ARKIT_BLENDSHAPES", "MOUTH_BLENDSHAPES"]
__version__ = "0.1.0"
