"""dcasr_b200 — B200-native drop-in for the hot path of anshulk-cmu/H-Net-Mamba-ASR (`dcasr`).

Public names mirror ``dcasr.models`` and ``mamba_ssm``:
    Mamba2, MambaBlock, MambaStack, reverse_sequences,
    ChunkOutput, RoutingModule, DynamicChunker, FixedPoolChunker, ratio_loss,
    DCASREncoder, EncoderOutput, ConvSubsampling4, build_chunker,
    CTCHead (``dcasr.decoders.ctc``)
``install()`` makes the reference's own ``dcasr`` package resolve to these.
Importing the package does not need a GPU; calling any op without the CUDA library raises.
"""
from ._lib import HnbError, launch_count, reset_launch_count  # noqa: F401
from .ctc import CTCHead, ctc_greedy_collapse  # noqa: F401
from .encoder import (ConvSubsampling4, DCASREncoder, EncoderOutput, build_chunker,  # noqa: F401
                      register_chunker)
from .fixed_pool import FixedPoolChunker  # noqa: F401
from .hnet_chunk import ChunkOutput, DynamicChunker, RoutingModule, ratio_loss  # noqa: F401
from .install import install  # noqa: F401
from .mamba_block import Mamba2, MambaBlock, MambaStack, reverse_sequences  # noqa: F401

__all__ = ["Mamba2", "MambaBlock", "MambaStack", "reverse_sequences", "ChunkOutput", "RoutingModule",
           "DynamicChunker", "FixedPoolChunker", "ratio_loss", "DCASREncoder", "EncoderOutput", "ConvSubsampling4",
           "build_chunker", "register_chunker", "install", "HnbError", "CTCHead", "ctc_greedy_collapse"]
