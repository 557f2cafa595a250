"""Drop-in `gaussian_renderer` package (reference: /root/reference gaussian_renderer/__init__.py:8-50)."""
from gaussianavatar_b200.renderer import render_batch  # noqa: F401

__all__ = ["render_batch"]
