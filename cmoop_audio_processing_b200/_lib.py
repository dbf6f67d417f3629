"""ctypes binding of libcmoop_b200.so (the C ABI declared in include/cmoop_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and
every compute entry point returns an error status (-> RuntimeError) when no CUDA
device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcmoop_b200.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


class GpModel(C.Structure):
    _fields_ = [("n_train", C.c_int), ("dim", C.c_int), ("amplitude", C.c_double), ("length_scale", C.c_double),
                ("nu", C.c_double), ("noise", C.c_double), ("y_scale", C.c_double), ("y_shift", C.c_double),
                ("x_train", c_double_p), ("alpha", c_double_p), ("chol_lower", c_double_p)]


class MfccConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("frame_length", C.c_int), ("hop", C.c_int), ("n_fft", C.c_int),
                ("n_mels", C.c_int), ("n_mfcc", C.c_int), ("f_min", C.c_float), ("f_max", C.c_float),
                ("log_floor", C.c_float), ("center", C.c_int)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header.
SIGNATURES = {
    "cmoop_abi_version": (C.c_int, []),
    "cmoop_last_error": (C.c_char_p, []),
    "cmoop_device_count": (C.c_int, []),
    "cmoop_set_device": (C.c_int, [C.c_int]),
    "cmoop_launch_count": (C.c_uint64, []),
    "cmoop_copy_bytes": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cmoop_profile_enable": (C.c_int, [C.c_int]),
    "cmoop_profile_read": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "cmoop_cnn_last_device_ms": (C.c_double, []),
    "cmoop_nds_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cmoop_nds_crowding_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "cmoop_nds_crowding_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmoop_crowding_distance_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                               C.c_void_p]),
    "cmoop_gp_create": (C.c_int, [C.POINTER(GpModel), C.c_int, C.POINTER(C.c_void_p)]),
    "cmoop_gp_destroy": (C.c_int, [C.c_void_p]),
    "cmoop_gp_predict_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "cmoop_gp_predict_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmoop_gp_lml_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_int, C.POINTER(C.c_void_p)]),
    "cmoop_gp_lml_destroy": (C.c_int, [C.c_void_p]),
    "cmoop_gp_lml_n_theta": (C.c_int, [C.c_void_p]),
    "cmoop_gp_lml_eval": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmoop_hypervolume_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cmoop_hypervolume_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "cmoop_hypervolume_workspace_bytes": (C.c_size_t, [C.c_int]),
    "cmoop_front_metrics_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_nondominated_mask_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_coverage_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_mfcc_create": (C.c_int, [C.POINTER(MfccConfig), C.POINTER(C.c_void_p)]),
    "cmoop_mfcc_destroy": (C.c_int, [C.c_void_p]),
    "cmoop_mfcc_n_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "cmoop_mfcc_n_out": (C.c_int, [C.c_void_p]),
    "cmoop_mfcc_fwd_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cmoop_mfcc_fwd_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "cmoop_mfcc_fwd_dev_i16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "cmoop_mfcc_fwd_host_i16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "cmoop_mfcc_set_standardise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmoop_cnn_dataset_create_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cmoop_cnn_dataset_create_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cmoop_cnn_dataset_destroy": (C.c_int, [C.c_void_p]),
    "cmoop_feature_stats_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmoop_cnn_param_count": (C.c_longlong, [C.c_void_p, C.c_void_p]),
    "cmoop_fpr_from_predictions_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                  C.c_void_p]),
    "cmoop_cnn_pop_train_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "cmoop_cnn_debug_init_params": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "cmoop_cnn_debug_permutation": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_cnn_debug_conv": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_cnn_debug_wgrad": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cmoop_cnn_debug_train_steps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
}

_lib = None


class CmoopError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises CmoopError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmoopError(
            f"{LIB_PATH} is missing: build it with `python -m cmoop_audio_processing_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def bind_device() -> None:
    """Make the library use the device torch has selected for this process (one process per GPU)."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        check(load().cmoop_set_device(int(torch.cuda.current_device())), "cmoop_set_device")


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().cmoop_last_error().decode("utf-8", "replace")
        raise CmoopError(f"{what} failed (status {status}): {msg}")


def copy_bytes() -> tuple[int, int]:
    """(host->device, device->host) bytes copied by the library since load."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    load().cmoop_copy_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def profile_table() -> dict:
    """{family: {"launches", "ms", "flops"}} accumulated since cmoop_profile_enable(1)."""
    lib = load()
    need = lib.cmoop_profile_read(None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    lib.cmoop_profile_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, fl = line.rsplit(" ", 3)
        out[name] = {"launches": int(n), "ms": float(ms), "flops": float(fl)}
    return out


def ptr(arr) -> C.c_void_p:
    """void* of a C-contiguous numpy array (or None)."""
    if arr is None:
        return None
    return C.c_void_p(arr.ctypes.data)
