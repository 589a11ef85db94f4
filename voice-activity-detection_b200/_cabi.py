"""ctypes binding of libvadb200.so (include/vadb200.h).  Thin: pointers and sizes only."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvadb200.so")

VADB_F32, VADB_BF16 = 0, 1

# every symbol include/vadb200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "vadb_create", "vadb_destroy", "vadb_last_error", "vadb_weight_count", "vadb_load_weights",
    "vadb_reserve", "vadb_forward", "vadb_forward_host", "vadb_predict_probabilities",
    "vadb_predict_probabilities_host", "vadb_attention", "vadb_positional_table",
    "vadb_launch_count", "vadb_version", "vadb_logmel_frames", "vadb_logmel", "vadb_logmel_tables",
    "vadb_predict_audio_host", "vadb_forward_host_async", "vadb_host_wait", "vadb_broadcast_weights", "vadb_forward_ragged",
)


class VadbConfig(C.Structure):
    _fields_ = [("feature_size", C.c_int32), ("num_layers", C.c_int32),
                ("d_model", C.c_int32), ("compute_dtype", C.c_int32)]


_lib = None


def load_library():
    """Load libvadb200.so or fail loudly -- the product path never falls back to CPU code."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C voice-activity-detection_b200/csrc`). vad_b200 has no CPU "
            "fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.vadb_create.argtypes = [C.POINTER(vp), C.POINTER(VadbConfig), i32]
    lib.vadb_create.restype = i32
    lib.vadb_destroy.argtypes = [vp]
    lib.vadb_destroy.restype = None
    lib.vadb_last_error.argtypes = [vp]
    lib.vadb_last_error.restype = C.c_char_p
    lib.vadb_weight_count.argtypes = [C.POINTER(VadbConfig)]
    lib.vadb_weight_count.restype = sz
    lib.vadb_load_weights.argtypes = [vp, vp, sz, i32, vp]
    lib.vadb_load_weights.restype = i32
    lib.vadb_broadcast_weights.argtypes = [vp, vp, i32, vp]
    lib.vadb_broadcast_weights.restype = i32
    lib.vadb_reserve.argtypes = [vp, i32, i32]
    lib.vadb_reserve.restype = i32
    lib.vadb_forward.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp]
    lib.vadb_forward.restype = i32
    lib.vadb_forward_ragged.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp]
    lib.vadb_forward_ragged.restype = i32
    lib.vadb_forward_host.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp]
    lib.vadb_forward_host.restype = i32
    lib.vadb_predict_probabilities.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.vadb_predict_probabilities.restype = i32
    lib.vadb_predict_probabilities_host.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.vadb_predict_probabilities_host.restype = i32
    lib.vadb_attention.argtypes = [vp, vp, vp, vp, vp, i32, vp, i32, i32, vp]
    lib.vadb_attention.restype = i32
    lib.vadb_positional_table.argtypes = [vp, i32, vp]
    lib.vadb_positional_table.restype = i32
    lib.vadb_launch_count.argtypes = [vp]
    lib.vadb_launch_count.restype = C.c_int64
    lib.vadb_logmel_frames.argtypes = [C.c_long, i32]
    lib.vadb_logmel_frames.restype = C.c_long
    lib.vadb_logmel.argtypes = [vp, vp, C.c_long, i32, i32, i32, i32, i32, vp, vp]
    lib.vadb_logmel.restype = i32
    lib.vadb_logmel_tables.argtypes = [i32, i32, i32, i32, vp, vp]
    lib.vadb_logmel_tables.restype = i32
    lib.vadb_predict_audio_host.argtypes = [vp, vp, C.c_long, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.vadb_predict_audio_host.restype = i32
    lib.vadb_forward_host_async.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, C.POINTER(C.c_long)]
    lib.vadb_forward_host_async.restype = i32
    lib.vadb_host_wait.argtypes = [vp, C.c_long]
    lib.vadb_host_wait.restype = i32
    lib.vadb_version.argtypes = []
    lib.vadb_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.vadb_last_error(handle)
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
