"""B200-native caption generation behind the reference's VideoCaptioningModel / Predictor API.

Import as ``video_captioning_b200`` (the importable alias of this ``video-captioning_b200/`` directory).
Compute = hand-written sm_100a CUDA in ``csrc/`` behind the C ABI of ``include/vc_b200.h``; this package
is the host-side mirror of the reference's Python interface.  No CPU / PyTorch fallback exists.
"""
from ._native import LIB_PATH, build_library, load_library  # noqa: F401
from .attention import (BahdanauAttention, LuongAttention, MultiHeadAttention,  # noqa: F401
                        create_attention_mechanism)
from .decoder import CaptionDecoder  # noqa: F401
from .encoder import VideoEncoder  # noqa: F401
from .predictor import (BatchPredictor, VideoCaptionPredictor, save_batch_results,  # noqa: F401
                        save_multiple_captions, save_single_result)
from .sharding import ShardedCaptioner, shard_bounds  # noqa: F401
from .video_captioning_model import VideoCaptioningModel  # noqa: F401
from .vocabulary import Vocabulary  # noqa: F401

__all__ = [
    "VideoCaptioningModel", "VideoEncoder", "CaptionDecoder", "BahdanauAttention", "LuongAttention",
    "MultiHeadAttention", "create_attention_mechanism", "VideoCaptionPredictor", "BatchPredictor", "Vocabulary",
    "ShardedCaptioner", "shard_bounds", "build_library", "load_library", "LIB_PATH", "save_single_result", "save_batch_results",
    "save_multiple_captions",
]
