"""genie_b200 — B200-native (sm_100a) implementation of GENIE's product-graph GNN front end.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C-ABI library
`libgenie_b200.so`), `capi.py` (ctypes binding of `include/genie_b200.h`), `plan.py` (graph plans),
`module.py` (drop-in `GCN_Detection_Network_extended`), `process_utils.py` (input/graph assembly mirrors) and
`synth.py` (seeded synthetic networks for tests and the benchmark).
"""
__version__ = '0.1.0'
