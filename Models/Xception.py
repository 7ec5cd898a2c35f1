"""Drop-in for the reference's Models/Xception.py (Xception.py:26,37-213): same public names."""
from multimodal_deepfake_detection_b200.modules import Block, SeparableConv2d, Xception, model_urls, xception  # noqa: F401

__all__ = ["xception"]
