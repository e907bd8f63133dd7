"""avsum_b200 -- B200-native drop-in for the AudioVidSum inference hot path.

The public Python surface mirrors the reference (``models.av_model.AVBiLSTMModel``,
``models.attention.MultiHeadSelfAttention``, ``features.fusion``, ``utils.*``,
``evaluation.*``); all arithmetic on the hot path runs in hand-written sm_100a
CUDA kernels behind the C ABI declared in ``include/avsum_b200.h``
(``csrc/`` -> ``libavsum_b200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
