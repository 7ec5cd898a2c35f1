"""Drop-in for the reference's Models/XceptionLSTMA.py (XceptionLSTMA.py:5-59)."""
from multimodal_deepfake_detection_b200.modules import XceptionLSTMA  # noqa: F401
