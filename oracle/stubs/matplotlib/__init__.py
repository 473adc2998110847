"""Empty stand-in for matplotlib (main.py:9, plotOnline.py:4-5); plotting is out of scope."""
from . import pyplot, cm  # noqa: F401
