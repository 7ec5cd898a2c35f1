"""Drop-in for the reference's Models/XceptionLSTMV.py (XceptionLSTMV.py:9-70)."""
from multimodal_deepfake_detection_b200.modules import XceptionLSTMV  # noqa: F401
