"""ctypes binding of include/mindrl_b200.h (the drop-in C ABI).

There is no CPU fallback: if the CUDA library is missing this module raises at first use, and
every compute call on a machine without a CUDA device raises RuntimeError(MRL_ERR_CUDA).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmindrl_b200.so")

MRL_OK, MRL_ERR_INVALID, MRL_ERR_CUDA, MRL_ERR_HANDLE, MRL_ERR_STATE = 0, 1, 2, 3, 4
TIME_MAJOR, BATCH_MAJOR = 0, 1
SCAN_AUTO, SCAN_SEQUENTIAL, SCAN_PARALLEL = 0, 1, 2
MAX_COLS = 32

vp, i64, i32, u64, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_float, C.c_size_t

# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/mindrl_b200.h 1:1
SIGNATURES = {
    "mrl_version": [],
    "mrl_last_error": [],
    "mrl_device_count": [vp],
    "mrl_set_device": [C.c_int],
    "mrl_get_device": [vp],
    "mrl_launch_count": [],
    "mrl_set_option": [C.c_char_p, i64],
    "mrl_malloc": [vp, sz],
    "mrl_free": [vp],
    "mrl_host_alloc": [vp, sz],
    "mrl_host_free": [vp],
    "mrl_memcpy_h2d": [vp, vp, sz, vp],
    "mrl_memcpy_d2h": [vp, vp, sz, vp],
    "mrl_memcpy_d2d": [vp, vp, sz, vp],
    "mrl_memset": [vp, C.c_int, sz, vp],
    "mrl_stream_sync": [vp],
    "mrl_fill_rows": [vp, i64, i64, u64, vp],
    "mrl_urb_create": [i64, i32, vp, vp, u64, vp],
    "mrl_urb_destroy": [i64],
    "mrl_urb_columns": [i64, vp],
    "mrl_urb_append": [i64, i64, vp, vp],
    "mrl_urb_append_host": [i64, i64, vp, vp],
    "mrl_transpose01": [vp, vp, i64, i64, i64, vp],
    "mrl_urb_write_at": [i64, i64, i64, vp, vp],
    "mrl_urb_write_at_host": [i64, i64, i64, vp, vp],
    "mrl_urb_get_item": [i64, i64, vp, vp],
    "mrl_urb_get_item_host": [i64, i64, vp, vp],
    "mrl_urb_gather": [i64, vp, i64, vp, vp],
    "mrl_urb_sample": [i64, i64, vp, vp, vp],
    "mrl_urb_sample_host": [i64, i64, vp, vp, vp],
    "mrl_urb_reset": [i64, vp],
    "mrl_urb_state": [i64, vp, vp],
    "mrl_urb_set_state": [i64, i32, i32],
    "mrl_prb_create": [i64, f32, i32, vp, u64, u64, vp],
    "mrl_prb_destroy": [i64],
    "mrl_prb_push": [i64, vp, vp],
    "mrl_prb_push_host": [i64, vp, vp],
    "mrl_prb_push_batch": [i64, i64, vp, vp],
    "mrl_prb_push_batch_host": [i64, i64, vp, vp],
    "mrl_prb_sample": [i64, i64, f32, vp, vp, vp, vp, vp],
    "mrl_prb_sample_host": [i64, i64, f32, vp, vp, vp, vp, vp],
    "mrl_prb_update": [i64, vp, vp, i64, vp],
    "mrl_prb_update_host": [i64, vp, vp, i64, vp],
    "mrl_prb_info": [i64, vp, vp, vp],
    "mrl_prb_max_priority": [i64, vp, vp],
    "mrl_prb_tree_dump": [i64, vp, vp, vp],
    "mrl_prb_columns": [i64, vp],
    "mrl_prb_get_state": [i64, vp, vp, vp, vp, vp],
    "mrl_prb_set_state": [i64, i64, i64, u64, f32, vp, vp, vp],
    "mrl_prb_root_stats": [i64, vp, vp],
    "mrl_prb_sample_sharded": [i64, i64, f32, vp, vp, i32, vp, vp, vp, vp],
    "mrl_prb_shard_init": [i64, i32, i32, vp],
    "mrl_prb_shard_set_mode": [i64, i32],
    "mrl_prb_shard_connect": [i64, vp, vp],
    "mrl_prb_shard_wait": [i64, vp],
    "mrl_prb_shard_diag": [i64, vp, vp, vp, vp, i32, vp],
    "mrl_prb_sample_sharded_p2p": [i64, i64, f32, vp, vp, vp, vp, vp],
    "mrl_prb_sample_sharded_p2p_host": [i64, i64, f32, vp, vp, vp, vp, vp],
    "mrl_discounted_return": [vp, vp, vp, vp, i64, i64, i64, f32, i32, i32, vp],
    "mrl_discounted_return_host": [vp, vp, vp, vp, i64, i64, i64, f32, i32, i32, vp],
    "mrl_gae": [vp, vp, vp, vp, vp, vp, vp, i64, i64, f32, f32, i32, i32, vp],
    "mrl_gae_host": [vp, vp, vp, vp, vp, vp, vp, i64, i64, f32, f32, i32, i32, vp],
    "mrl_normalize": [vp, vp, i64, f32, vp],
    "mrl_gae_normalized": [vp, vp, vp, vp, vp, vp, vp, i64, i64, f32, f32, f32, i32, i32, vp],
    "mrl_discounted_return_normalized": [vp, vp, vp, vp, i64, i64, i64, f32, f32, i32, i32, vp],
    "mrl_td_lambda": [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, f32, f32, vp],
    "mrl_vtrace": [vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, i64, i64, i32, vp],
    "mrl_affine_scan": [vp, vp, vp, vp, i64, i64, i32, i32, vp],
}
# include/mindrl_b200_aot.h: the MindSpore "aot" custom-operator entry points and the stand-in AotExtra
AOT_OPS = ["BufferAppend", "BufferGetItem", "BufferSample", "PriorityReplayBufferCreate", "PriorityReplayBufferPush",
           "PriorityReplayBufferSample", "PriorityReplayBufferUpdate", "PriorityReplayBufferDestroy", "DiscountedReturn",
           "GaeScan"]
AOT_SIGNATURES = {
    "mrl_aot_extra_create": [],
    "mrl_aot_extra_destroy": [vp],
    "mrl_aot_extra_set_bool": [vp, C.c_char_p, i32],
    "mrl_aot_extra_set_int": [vp, C.c_char_p, i64],
    "mrl_aot_extra_set_float": [vp, C.c_char_p, f32],
    "mrl_aot_extra_set_str": [vp, C.c_char_p, C.c_char_p],
    "mrl_aot_extra_set_int2d": [vp, C.c_char_p, vp, vp, i32],
}
for _op in AOT_OPS:
    AOT_SIGNATURES[_op] = [C.c_int, vp, vp, vp, vp, vp, vp]
    AOT_SIGNATURES[_op + "Init"] = [vp, vp, vp, vp]
_RESTYPES = {"mrl_last_error": C.c_char_p, "mrl_launch_count": i64, "mrl_aot_extra_create": vp,
             "mrl_aot_extra_destroy": None}

_lib = None


class MindRLError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mindrl_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Loads libmindrl_b200.so (built in-tree by mindrl_b200/build.py). Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mindrl_b200.build` "
                "(there is no CPU fallback for the experience data path)")
        L = C.CDLL(LIB_PATH)
        for name, argtypes in list(SIGNATURES.items()) + list(AOT_SIGNATURES.items()):
            fn = getattr(L, name)  # AttributeError here = header and library out of sync
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc):
    if rc != MRL_OK:
        msg = lib().mrl_last_error().decode("utf-8", "replace")
        if rc == MRL_ERR_INVALID:
            raise ValueError(msg)
        raise MindRLError(rc, msg)


def ptr_table(ptrs):
    """host array of void* for the column tables of the ABI"""
    tab = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        tab[i] = p
    return tab


def set_option(key, value):
    check(lib().mrl_set_option(key.encode(), int(value)))


def launch_count():
    return int(lib().mrl_launch_count())
