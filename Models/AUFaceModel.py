"""`Models.AUFaceModel` as imported by train_au_face.py:409 / test_au_face.py:13 (absent from the reference tree;
interface inferred from the call sites)."""
from multimodal_deepfake_detection_b200.modules import AUFaceCrossDetector, FusionHead  # noqa: F401
