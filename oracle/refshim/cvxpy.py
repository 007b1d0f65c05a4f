"""Empty cvxpy stub: competitive assignment is outside the hot path."""
