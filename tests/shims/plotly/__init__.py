"""Stand-in for plotly (training/train_muse.py:28 imports plotly.express for one optional histogram)."""
from . import express  # noqa: F401
