"""Input side of the path: the reference's `.npy` loaders (tuple layouts kept) plus synthetic generators used when the
datasets are not on disk (this environment has none)."""
