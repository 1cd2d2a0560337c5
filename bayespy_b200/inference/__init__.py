"""``from bayespy_b200.inference import VB`` (bayespy/inference/__init__.py:34)."""
from .vb import VB      # noqa: F401
