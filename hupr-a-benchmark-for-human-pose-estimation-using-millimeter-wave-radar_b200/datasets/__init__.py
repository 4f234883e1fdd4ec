"""Host-side mirror of /root/reference/datasets (``getDataset``, ``HuPR3D_horivert``)."""
from .dataset import HuPR3D_horivert, generateGTAnnot, getDataset, window_frame_indices  # noqa: F401
