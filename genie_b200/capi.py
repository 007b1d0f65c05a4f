"""ctypes binding of include/genie_b200.h (libgenie_b200.so).

This is the stub a maintainer of the reference would add to call the B200 library from `module.py` /
`process_utils.py`.  torch is used only to own device memory and streams: every call receives raw
device pointers (`tensor.data_ptr()`) and the current CUDA stream handle.

There is no CPU fallback: if the library is missing it is built with nvcc; if that is impossible, or no CUDA device is
present when an op is called, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libgenie_b200.so')
_lib = None

GRAPH_CARTESIAN, GRAPH_EXPLICIT = 0, 1
ABI_VERSION = 8
PEER_HANDLE_BYTES = 64         # GENIE_PEER_HANDLE_BYTES
HEADS_PROJ_LD = 160            # GENIE_HEADS_PROJ_LD
EDGE_TERM_LD = 48           # GENIE_EDGE_TERM_LD
STORAGE_FP32, STORAGE_BF16 = 0, 1

c_f32p = ctypes.c_void_p   # device pointers are passed as opaque addresses


class GraphDesc(ctypes.Structure):
    _fields_ = [('mode', ctypes.c_int32), ('n_sta', ctypes.c_int32), ('n_grid', ctypes.c_int32),
                ('sta_max_deg', ctypes.c_int32), ('n_prod', ctypes.c_int64),
                ('sta_rowptr', ctypes.c_void_p), ('sta_col', ctypes.c_void_p),
                ('src_rowptr', ctypes.c_void_p), ('src_col', ctypes.c_void_p),
                ('grid_rowptr', ctypes.c_void_p), ('grid_col', ctypes.c_void_p),
                ('grid_outdeg', ctypes.c_void_p), ('prod_grid', ctypes.c_void_p),
                ('grid_order', ctypes.c_void_p),
                ('n_sta_tiles', ctypes.c_int32), ('n_grid_groups', ctypes.c_int32),
                ('sta_tile_rows', ctypes.c_void_p), ('sta_tile_meta', ctypes.c_void_p),
                ('sta_tile_nbr', ctypes.c_void_p), ('sta_tile_invdeg', ctypes.c_void_p),
                ('grid_grp_ptr', ctypes.c_void_p), ('grid_grp_nodes', ctypes.c_void_p),
                ('n_grid_owned', ctypes.c_int32)]


class Linear(ctypes.Structure):
    _fields_ = [('weight', ctypes.c_void_p), ('bias', ctypes.c_void_p)]


class SAWeights(ctypes.Structure):
    _fields_ = [('fc1', Linear), ('fc2', Linear), ('fglobal', Linear),
                ('activate1', ctypes.c_void_p), ('activate2', ctypes.c_void_p), ('activate3', ctypes.c_void_p)]


class FrontendWeights(ctypes.Structure):
    _fields_ = [('da_init_trns', Linear), ('da_l1_t1_2', Linear), ('da_l1_t2_2', Linear),
                ('da_l2_t1_1', Linear), ('da_l2_t2_1', Linear), ('da_l2_t1_2', Linear), ('da_l2_t2_2', Linear),
                ('da_activate', ctypes.c_void_p), ('da_activate11', ctypes.c_void_p),
                ('da_activate12', ctypes.c_void_p), ('da_activate1', ctypes.c_void_p),
                ('da_activate21', ctypes.c_void_p), ('da_activate22', ctypes.c_void_p),
                ('da_activate2', ctypes.c_void_p),
                ('ri_fc1', Linear), ('ri_fc2', Linear),
                ('ri_activate1', ctypes.c_void_p), ('ri_activate2', ctypes.c_void_p),
                ('sa', SAWeights * 3)]


class NearestParams(ctypes.Structure):
    _fields_ = [('offset_per_batch', ctypes.c_double), ('offset_per_station', ctypes.c_double),
                ('kernel_sig_t', ctypes.c_double), ('n_all', ctypes.c_int64), ('n_p', ctypes.c_int64),
                ('n_s', ctypes.c_int64), ('n_batch', ctypes.c_int32), ('n_grid', ctypes.c_int32),
                ('n_locs', ctypes.c_int32), ('n_sta_use', ctypes.c_int32)]


class InputParams(ctypes.Structure):
    _fields_ = [('t0', ctypes.c_double), ('max_t', ctypes.c_double), ('kernel_sig_t', ctypes.c_double),
                ('dt', ctypes.c_double), ('ref0', ctypes.c_double), ('ref_step', ctypes.c_double),
                ('n_ts', ctypes.c_int32), ('n_extra', ctypes.c_int32), ('n_locs', ctypes.c_int32),
                ('n_sta_use', ctypes.c_int32), ('use_sign_input', ctypes.c_int32), ('reserved_', ctypes.c_int32)]


class MlpDesc(ctypes.Structure):
    """genie_mlp_desc_t."""
    _fields_ = [('n_rows', ctypes.c_int64), ('n_parts', ctypes.c_int32), ('n_out', ctypes.c_int32),
                ('width', ctypes.c_int32 * 4), ('ld', ctypes.c_int32 * 4), ('x', ctypes.c_void_p * 4),
                ('weight', ctypes.c_void_p), ('bias', ctypes.c_void_p), ('slope', ctypes.c_void_p)]


class WindowParams(ctypes.Structure):
    """genie_window_params_t: the device-resident per-window block of genie_window_fwd."""
    _fields_ = [('prm', InputParams), ('pick_lo', ctypes.c_int64), ('pick_hi', ctypes.c_int64)]


# name -> (restype, argtypes); every symbol include/genie_b200.h declares
_P = ctypes.c_void_p
SIGNATURES = {
    'genie_last_error': (ctypes.c_char_p, []),
    'genie_abi_version': (ctypes.c_int, []),
    'genie_launch_count': (ctypes.c_int64, []),
    'genie_timing_enable': (ctypes.c_int, [ctypes.c_int]),
    'genie_timing_kernel_count': (ctypes.c_int, []),
    'genie_timing_kernel_name': (ctypes.c_char_p, [ctypes.c_int]),
    'genie_timing_collect': (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                            ctypes.c_int]),
    'genie_debug_trace': (ctypes.c_int, [_P, ctypes.c_int]),
    'genie_plan_create': (ctypes.c_int, [ctypes.POINTER(GraphDesc), ctypes.POINTER(_P)]),
    'genie_plan_destroy': (None, [_P]),
    'genie_plan_workspace_bytes': (ctypes.c_size_t, [_P]),
    'genie_plan_set_storage': (ctypes.c_int, [_P, ctypes.c_int32]),
    'genie_peer_alloc': (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    'genie_peer_open': (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    'genie_peer_close': (ctypes.c_int, [_P]),
    'genie_peer_free': (ctypes.c_int, [_P]),
    'genie_plan_set_halo_export': (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    'genie_plan_set_edge_terms': (ctypes.c_int, [_P, _P, _P]),
    'genie_plan_set_init_terms': (ctypes.c_int, [_P, _P, _P]),
    'genie_frontend_packed_floats': (ctypes.c_size_t, []),
    'genie_frontend_pack_weights': (ctypes.c_int, [ctypes.POINTER(FrontendWeights), _P, _P]),
    'genie_heads_packed_floats': (ctypes.c_size_t, []),
    'genie_heads_layout': (ctypes.c_int, [_P, ctypes.c_int]),
    'genie_heads_grid_fwd': (ctypes.c_int, [_P, _P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, _P, _P, _P]),
    'genie_heads_query_fwd': (ctypes.c_int, [_P, _P, ctypes.c_int, _P, ctypes.c_int, _P, _P, _P, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float, _P, _P, _P]),
    'genie_kron_spmm_fwd': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, _P, _P, _P, _P, ctypes.c_int,
                                           ctypes.c_int, _P, ctypes.c_int, _P]),
    'genie_node_mlp_partial_rows': (ctypes.c_int, []),
    'genie_node_mlp_fwd': (ctypes.c_int, [ctypes.POINTER(MlpDesc), _P, ctypes.c_int32, _P, _P]),
    'genie_node_mlp_bwd': (ctypes.c_int, [ctypes.POINTER(MlpDesc), _P, ctypes.c_int32, _P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.POINTER(ctypes.c_int32), _P, _P]),
    'genie_stack_output_fwd': (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_float, _P, ctypes.c_int64,
                                              _P]),
    'genie_knn_fwd': (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_int, _P, _P]),
    'genie_assoc_packed_floats': (ctypes.c_size_t, []),
    'genie_assoc_layout': (ctypes.c_int, [_P, ctypes.c_int]),
    'genie_assoc_workspace_bytes': (ctypes.c_size_t, [_P]),
    'genie_assoc_set_terms': (ctypes.c_int, [_P, _P, _P, _P, _P]),
    'genie_assoc_product_fwd': (ctypes.c_int, [_P, _P, _P, ctypes.c_int, _P, ctypes.c_int, ctypes.c_float, _P, _P, _P, _P,
                                               _P, _P, ctypes.POINTER(ctypes.c_void_p), _P]),
    'genie_assoc_collapse_fwd': (ctypes.c_int, [_P, _P, ctypes.c_int64, _P, _P, ctypes.c_int64, _P, _P, _P, _P,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_float, _P, _P]),
    'genie_input_nearest_fwd': (ctypes.c_int, [ctypes.POINTER(NearestParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    'genie_input_scatter_fwd': (ctypes.c_int, [_P, ctypes.POINTER(InputParams), _P, ctypes.c_int64, _P, _P, _P, _P, _P,
                                               _P, _P, _P, _P, _P]),
    'genie_data_aggregation_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    'genie_bipartite_readin_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    'genie_spatial_aggregation_fwd': (ctypes.c_int, [_P, _P, ctypes.c_int32, _P, _P, ctypes.c_float, _P, _P, _P]),
    'genie_da_layer1_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    'genie_workspace_region': (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.POINTER(ctypes.c_size_t)]),
    'genie_da_layer2_readin_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    'genie_frontend_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, ctypes.c_float, _P, _P, _P, _P, _P]),
    'genie_window_fwd': (ctypes.c_int, [_P, _P, _P, ctypes.c_int64, ctypes.c_int32, _P, _P, _P, _P, _P, ctypes.c_int32, _P, _P, ctypes.c_float,
                                        _P, _P, _P, _P, _P, _P, _P]),
}


class GenieError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load(build_if_missing=True):
    """Load libgenie_b200.so (building it in-tree first if it is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        if not build_if_missing:
            raise GenieError('libgenie_b200.so is missing: run `python -m genie_b200.build`')
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.genie_abi_version() != ABI_VERSION:
        raise GenieError('libgenie_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GenieError('libgenie_b200: %s (status %d)' % (load().genie_last_error().decode(), rc))


def dptr(t, dtype=None, name='tensor'):
    """Raw device address of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GenieError('%s must live on a CUDA device (there is no CPU path)' % name)
    if not t.is_contiguous():
        raise GenieError('%s must be contiguous' % name)
    if dtype is not None and t.dtype != dtype:
        raise GenieError('%s must be %s, got %s' % (name, dtype, t.dtype))
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(load().genie_launch_count())


def timing_enable(on=True):
    check(load().genie_timing_enable(1 if on else 0))


def timing_collect(reset=True):
    """{kernel name: (total device ms, launches)} of the launches recorded since the last reset."""
    lib = load()
    n = lib.genie_timing_kernel_count()
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    check(lib.genie_timing_collect(ms, cnt, 1 if reset else 0))
    return {lib.genie_timing_kernel_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}
