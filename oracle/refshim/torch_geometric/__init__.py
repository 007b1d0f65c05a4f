"""Stub of torch_geometric (test infrastructure; see oracle/refshim/README.md)."""
from . import utils, data, nn  # noqa: F401
