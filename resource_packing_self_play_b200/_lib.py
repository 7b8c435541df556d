"""ctypes binding of the C ABI declared in include/bpp_b200.h (csrc/libbpp_b200.so).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
Build the library with `python __graft_entry__.py build` (or `python -m resource_packing_self_play_b200.build`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbpp_b200.so")

OK = 0
STUB = {"U": 1, "V": 2, "H": 3, "D": 4}
DTYPE_F32, DTYPE_F64 = 0, 1
CHOOSE_ARGMAX_FIRST, CHOOSE_SAMPLE, CHOOSE_GREEDY = 0, 1, 2
REC_WORDS, REC_REM = 32, 28


class BppError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bpp_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("G", C.c_int32), ("num_sims", C.c_int32),
        ("cpuct", C.c_double), ("node_cap", C.c_int32), ("edge_cap", C.c_int64), ("device", C.c_int32),
    ]


_vp, _i32, _i64, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "bpp_last_error": [],
    "bpp_version": [],
    "bpp_env_valid_moves": [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "bpp_items_generate": [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "bpp_env_planes": [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "bpp_env_next_state": [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "bpp_env_game_ended": [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_create": [C.POINTER(Config), C.POINTER(_vp)],
    "bpp_engine_destroy": [_vp],
    "bpp_engine_device_bytes": [_vp],
    "bpp_engine_edge_cap": [_vp, C.POINTER(_i64)],
    "bpp_engine_reset": [_vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_reset_host": [_vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_set_roots": [_vp, _vp, _vp],
    "bpp_engine_begin_move": [_vp, _vp],
    "bpp_engine_set_max_h": [_vp, _vp, _vp],
    "bpp_engine_set_num_sims": [_vp, _i32],
    "bpp_engine_last_values": [_vp, _vp, _vp],
    "bpp_engine_set_select_cap": [_vp, _i32],
    "bpp_engine_unfinished": [_vp, C.POINTER(_i32)],
    "bpp_engine_select": [_vp, _vp],
    "bpp_engine_leaf_count": [_vp, C.POINTER(_i32), _vp],
    "bpp_engine_leaf_buffers": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "bpp_engine_leaf_planes": [_vp, _vp, _vp],
    "bpp_engine_expand_backup": [_vp, _vp, _i32, _vp, _i32, _vp],
    "bpp_engine_expand_select": [_vp, _vp, _i32, _vp, _i32, _vp],
    "bpp_engine_leaf_count_async": [_vp, _vp, _vp],
    "bpp_engine_search_stub": [_vp, _i32, _vp],
    "bpp_engine_root_counts": [_vp, _vp, _vp],
    "bpp_engine_root_counts_host": [_vp, _vp, _vp],
    "bpp_engine_choose": [_vp, _i32, _u64, _vp, _vp],
    "bpp_engine_advance": [_vp, _vp, _vp],
    "bpp_engine_status": [_vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_roots": [_vp, _vp, _vp],
    "bpp_engine_play_stub": [_vp, _i32, _i32, _u64, _i32, _vp, _vp, C.POINTER(_i32), _vp],
    "bpp_engine_play_stub_host": [_vp, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_play_stub_stream": [_vp, _i32, _i32, _u64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_play_stub_stream_host": [_vp, _i32, _i32, _u64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_engine_set_auto_play": [_vp, _i32, _u64, _vp, _vp, _vp],
    "bpp_engine_progress_async": [_vp, _vp, _vp],
    "bpp_engine_play_net": [_vp, _vp, _i32, _u64, _vp, _vp, _vp, C.POINTER(_i32), _vp],
    "bpp_engine_play_net_host": [_vp, _vp, _i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 C.POINTER(_i32), _vp],
    "bpp_engine_play_net_stream": [_vp, _vp, _i32, _u64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   C.POINTER(_i32), _vp],
    "bpp_engine_play_net_stream_host": [_vp, _vp, _i32, _u64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        C.POINTER(_i32), _vp],
    "bpp_engine_set_profile": [_vp, _i32],
    "bpp_engine_profile": [_vp, C.POINTER(C.c_double)],
    "bpp_engine_stats": [_vp, C.POINTER(_u64), _i32, _vp],
    "bpp_engine_check": [_vp, _vp],
    "bpp_engine_export_game": [_vp, _i32, _vp, _i32, _vp, _i64, C.POINTER(_i32), C.POINTER(_i64), _vp],
    "bpp_engine_graph_sizes": [_vp, _vp, _vp, _vp],
    "bpp_net_create": [_i32, _i32, _i32, _i32, _i32, C.POINTER(_vp)],
    "bpp_net_destroy": [_vp],
    "bpp_net_set_param": [_vp, C.c_char_p, _vp, _i64],
    "bpp_net_commit": [_vp, _vp],
    "bpp_net_set_precision": [_vp, _i32],
    "bpp_net_profile": [_vp, C.POINTER(_i64)],
    "bpp_net_profile_roles": [_vp, C.POINTER(_i64)],
    "bpp_net_grid_row": [_vp],
    "bpp_net_forward": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_learner_create": [_i32, _i32, _i32, _i32, _i32, C.POINTER(_vp)],
    "bpp_learner_destroy": [_vp],
    "bpp_learner_num_params": [_vp, C.POINTER(_i64)],
    "bpp_learner_param_offset": [_vp, C.c_char_p, C.POINTER(_i64), C.POINTER(_i64)],
    "bpp_learner_grad": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bpp_learner_adam": [_i64, _vp, _vp, _vp, _vp, _i32, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _vp],
}
_RESTYPES = {"bpp_last_error": C.c_char_p, "bpp_engine_device_bytes": C.c_int64}

_lib = None


def load():
    """Load the shared library once; raise if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BppError(-100, f"{LIB_PATH} not found: build it first (python __graft_entry__.py build). "
                             "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise BppError(rc, load().bpp_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(load(), name)(*args))
