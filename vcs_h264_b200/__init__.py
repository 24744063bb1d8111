"""vcs_h264_b200 -- B200-native (sm_100a) implementation of the VCS-h264 interframe hot path:
block-matching motion estimation, motion-compensated residual, 8x8 DCT / quantise / dequantise /
IDCT reconstruction, behind the reference's own class surface (MotionProcessor, DCTCompressor,
Encoder, Decoder, Frame) plus a clip-level batched API (ClipEncoder).  CUDA only: importing works
anywhere, every operation needs libvcs_b200.so and a GPU."""
from . import _capi
from ._capi import (COEF_F64, COEF_F64_RINT, COEF_I16_RINT, COEF_I8_RINT, METRIC_SAD, METRIC_WRAP8, ME_AUTO,
                    ME_GENERIC, ME_TILED, VcsError)
from .frame import Frame
from .motion import MotionProcessor
from .DCTcompressor import DCTCompressor
from .encoder import Encoder
from .decoder import Decoder
from .clip import ClipDecoder, ClipEncoder, flip_counters, sparsity_device

__all__ = ["MotionProcessor", "DCTCompressor", "Encoder", "Decoder", "Frame", "ClipEncoder", "ClipDecoder", "sparsity_device", "flip_counters",
           "VcsError", "_capi"]
