"""pymunk.shape_filter stand-in (oracle only)."""
from . import ShapeFilter  # noqa: F401
