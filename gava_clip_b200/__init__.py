"""gava_clip_b200: B200-native (sm_100a) Vita-CLIP / GaVA-CLIP video encoder forward + logit head.

Drop-in for the reference's `training/VitaCLIP_model.py` API:
    from gava_clip_b200 import VitaCLIP, CLIPVisionEncoder
"""
from .model import VitaCLIP  # noqa: F401
from .input_stage import EvalInputStage, aggregate_views  # noqa: F401
from .streaming import ClipStreamer  # noqa: F401
from .text_encoder import CLIPTextEncoder, TextPromptLearner, tokenize  # noqa: F401
from .vision_encoder import CLIPVisionEncoder  # noqa: F401

__all__ = ["VitaCLIP", "ClipStreamer", "EvalInputStage", "aggregate_views", "CLIPVisionEncoder", "CLIPTextEncoder", "TextPromptLearner", "tokenize"]
