# -*- coding: utf-8 -*-
"""ctypes binding of the C ABI declared in ``include/swn_b200.h``.

The library is hand-written CUDA for sm_100a built in-tree
(``shallow-wavenet_b200/lib/libswn_b200.so``, see ``csrc/Makefile``).  There is no
Python/CPU implementation behind this module: if the library is missing or no B200 is
present, every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWN_B200_LIB selects another build of the same library (kernel A/B timing, tools/ab_build.sh); never a fallback
LIB_PATH = os.environ.get("SWN_B200_LIB") or os.path.join(_HERE, "lib", "libswn_b200.so")

KIND_CSWNV = 0
KIND_DSWNV = 1

# every symbol include/swn_b200.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = ("swn_create", "swn_set_tensor", "swn_finalize", "swn_condition", "swn_generate", "swn_teacher_forced", "swn_forward", "swn_sync",
           "swn_generate_host", "swn_get_info", "swn_philox_uniform", "swn_measure_smem_bandwidth", "swn_noise_shaping", "swn_destroy",
           "swn_last_error")


class SwnConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "kind", "n_quantize", "n_aux", "hid_chn", "skip_chn", "aux_kernel_size", "aux_dilation_size",
        "dilation_depth", "dilation_repeat", "kernel_size", "upsampling_factor", "seg", "lpc",
        "aux_conv2d_flag", "wav_conv_flag", "audio_in_flag")]


class SwnGenArgs(ctypes.Structure):
    _fields_ = [("n_samples", ctypes.c_int64), ("seed", ctypes.c_uint64), ("noise_dev", ctypes.c_void_p),
                ("forced_dev", ctypes.c_void_p), ("out_dev", ctypes.c_void_p), ("head_dev", ctypes.c_void_p),
                ("seed_dev", ctypes.c_void_p), ("batch_tile", ctypes.c_int32), ("poll_mode", ctypes.c_int32),
                ("prof_dev", ctypes.c_void_p), ("prof_start", ctypes.c_int32), ("prof_count", ctypes.c_int32),
                ("kernel_sel", ctypes.c_int32), ("debug_helper_delay_ns", ctypes.c_int32), ("n_samples_list", ctypes.c_void_p)]


class SwnFwdArgs(ctypes.Structure):
    _fields_ = [("audio_dev", ctypes.c_void_p), ("audio_onehot", ctypes.c_int32), ("clip", ctypes.c_int32), ("T_a", ctypes.c_int64),
                ("out_dev", ctypes.c_void_p), ("b_dev", ctypes.c_void_p), ("log_b_dev", ctypes.c_void_p),
                ("b_clip_dev", ctypes.c_void_p), ("log_b_clip_dev", ctypes.c_void_p)]


class SwnNsArgs(ctypes.Structure):
    _fields_ = [("x_dev", ctypes.c_void_p), ("y_dev", ctypes.c_void_p), ("tmp_dev", ctypes.c_void_p), ("stride", ctypes.c_int64),
                ("B", ctypes.c_int32), ("order", ctypes.c_int32), ("n_samples", ctypes.c_void_p), ("n_frames", ctypes.c_void_p),
                ("mcep_host", ctypes.c_void_p), ("mcep_utts", ctypes.c_int32), ("mcep_frames", ctypes.c_int32),
                ("pade", ctypes.c_int32), ("hop", ctypes.c_int32), ("alpha", ctypes.c_double), ("fir_host", ctypes.c_void_p),
                ("fir_taps", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class SwnInfo(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("receptive_field", ctypes.c_int32), ("eff_seg", ctypes.c_int32),
                ("head_dim", ctypes.c_int32), ("n_sm", ctypes.c_int32), ("ctas_per_layer", ctypes.c_int32),
                ("head_ctas", ctypes.c_int32), ("smem_bytes", ctypes.c_int32), ("batch_tile", ctypes.c_int32),
                ("n_slots", ctypes.c_int32), ("weight_bytes", ctypes.c_int64), ("cond_bytes", ctypes.c_int64),
                ("last_cond_ms", ctypes.c_float), ("last_gen_ms", ctypes.c_float), ("kernel_launches", ctypes.c_int64),
                ("grid", ctypes.c_int32), ("kernel_variant", ctypes.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SwnError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library (no fallback: a missing build is an error)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SwnError("%s not found: build it with `python __graft_entry__.py build` "
                       "(or `make -C shallow-wavenet_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    lib.swn_create.argtypes = [ctypes.POINTER(SwnConfig), ctypes.c_int, ctypes.POINTER(vp)]
    lib.swn_set_tensor.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32]
    lib.swn_finalize.argtypes = [vp]
    lib.swn_condition.argtypes = [vp, vp, i32, i32, vp]
    lib.swn_generate.argtypes = [vp, ctypes.POINTER(SwnGenArgs), vp]
    lib.swn_teacher_forced.argtypes = [vp, vp, i64, vp, vp]
    lib.swn_forward.argtypes = [vp, ctypes.POINTER(SwnFwdArgs), vp]
    lib.swn_sync.argtypes = [vp, vp]
    lib.swn_generate_host.argtypes = [vp, vp, i32, i32, i64, u64, vp]
    lib.swn_get_info.argtypes = [vp, ctypes.POINTER(SwnInfo)]
    lib.swn_philox_uniform.argtypes = [u64, i32, i32, i64, i64, i64, vp]
    lib.swn_measure_smem_bandwidth.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.swn_noise_shaping.argtypes = [ctypes.c_int, ctypes.POINTER(SwnNsArgs), vp]
    lib.swn_destroy.argtypes = [vp]
    lib.swn_last_error.argtypes = []
    lib.swn_last_error.restype = ctypes.c_char_p
    for s in SYMBOLS:
        if s != "swn_last_error":
            getattr(lib, s).restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SwnError(load().swn_last_error().decode("utf-8", "replace"))
