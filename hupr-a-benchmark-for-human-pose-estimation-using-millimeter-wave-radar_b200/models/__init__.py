"""Host-side mirror of /root/reference/models: same class name (``HuPRNet``), constructor argument (``cfg``), forward
signature and 255-entry ``state_dict`` — every arithmetic op is a call into libhupr_b200.so (sm_100a)."""
from .networks import HuPRNet  # noqa: F401
