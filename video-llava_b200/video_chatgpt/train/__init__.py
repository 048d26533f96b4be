"""Training-side READER of the offline CLIP features (the trainer itself is out of scope)."""
from .feature_reader import collate_video_features, load_video_features  # noqa: F401
