"""Limiter ids (src/pyclaw/limiters/__init__.py, limiters/tvd.py:74-79)."""
from . import tvd

__all__ = ["tvd"]
