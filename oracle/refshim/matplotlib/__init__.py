"""Empty matplotlib stub (plots are outside the hot path)."""
from . import pyplot  # noqa: F401
