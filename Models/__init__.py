"""Import surface of the reference (`from Models.XceptionLSTMV import XceptionLSTMV`, train_visual.py:450,
train_audio.py:4).  The classes live in multimodal_deepfake_detection_b200.modules."""
