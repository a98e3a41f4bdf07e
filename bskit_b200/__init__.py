"""bskit_b200 — B200-native FFT bispectrum estimator behind the bskit.main API.

Like the reference package (``bskit/__init__.py:6``) the public names live in
``bskit_b200.main`` and are re-exported here.
"""
from .main import *  # noqa: F401,F403
from . import main  # noqa: F401

__version__ = "0.1.0"
