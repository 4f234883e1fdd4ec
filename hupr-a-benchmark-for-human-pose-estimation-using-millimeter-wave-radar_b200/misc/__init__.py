"""Host-side mirror of /root/reference/misc for the hot-path pieces (loss, target synthesis, keypoint decode)."""
from .losses import LossComputer, generateTarget, get_max_preds  # noqa: F401
