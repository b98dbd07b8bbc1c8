"""`altcorr` operator API (devo/altcorr/correlation.py:51-72): corr, patchify."""
from .correlation import corr, patchify, CorrLayer, PatchLayer  # noqa: F401
