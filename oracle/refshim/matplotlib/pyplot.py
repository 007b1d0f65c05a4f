"""Empty matplotlib.pyplot stub."""
