"""B200-native MHAda forward hot path: a drop-in for the reference's `network` package surface on that
path (MHAdaSTr/network/__init__.py:1-3, MHAdaSTr/network/adaDecoder.py).

    from mhada_style_transfer_b200 import AdaAttnTransformerMultiHead, AdaAttnMultiHead

Same constructors, forward signatures, attribute names and state_dict keys as the reference classes;
the forward runs hand-written sm_100a kernels through the C ABI in include/mhada_b200.h.
"""
from .network import (AdaAttN, AdaAttnForLoss, AdaAttnMultiHead, AdaAttnTransformer,  # noqa: F401
                      AdaAttnTransformerMultiHead, CosineSimilarity, Decoder, Softmax, StyleCache)

from .vit import VisionTransformer  # noqa: F401,E402

__all__ = ["VisionTransformer", "AdaAttnTransformer", "AdaAttnTransformerMultiHead", "AdaAttnForLoss", "AdaAttnMultiHead", "AdaAttN",
           "Decoder", "Softmax", "CosineSimilarity", "StyleCache"]
