"""Host-side mirror of /root/reference/tools (``Runner``)."""
from .run import Runner  # noqa: F401
