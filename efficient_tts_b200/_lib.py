"""ctypes binding of ``libefts_b200.so`` (C ABI: ``include/efts_b200.h``).

There is no CPU or PyTorch fallback behind this module: if the library is missing or cannot be
built, or no sm_100 device is present, every entry point raises.
"""
import ctypes
import os
import threading

from . import build as _build

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_void_p, c_size_t, c_char_p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p


class EftsConfig(ctypes.Structure):
    _fields_ = [("num_symbols", c_i32), ("odim", c_i32), ("n_channels", c_i32), ("k_size", c_i32),
                ("n_text_encoder_layer", c_i32), ("n_mel_encoder_layer", c_i32),
                ("n_decoder_layer", c_i32), ("n_duration_layer", c_i32),
                ("duration_kernel_size", c_i32), ("sigma", c_f32), ("sigma_e", c_f32),
                ("duration_offset", c_f32), ("leaky_relu_slope", c_f32), ("use_masking", c_i32),
                ("device", c_i32)]


class EftsVocoderConfig(ctypes.Structure):
    _fields_ = [("num_mels", c_i32), ("upsample_initial_channel", c_i32), ("num_upsamples", c_i32),
                ("upsample_rates", c_i32 * 8), ("upsample_kernel_sizes", c_i32 * 8), ("num_kernels", c_i32),
                ("resblock_kernel_sizes", c_i32 * 4), ("resblock_dilations", (c_i32 * 3) * 4), ("resblock_type", c_i32), ("num_dilations", c_i32),
                ("device", c_i32)]


class EftsFrontendConfig(ctypes.Structure):
    _fields_ = [("n_fft", c_i32), ("hop_size", c_i32), ("win_size", c_i32), ("num_mels", c_i32), ("device", c_i32)]


# name -> (restype, argtypes); also the export list checked by
# tests/test_host_logic.py::test_library_builds_loads_and_exports_the_header
SIGNATURES = {
    "efts_create": (c_i32, [ctypes.POINTER(EftsConfig), ctypes.POINTER(c_void_p)]),
    "efts_destroy": (None, [c_void_p]),
    "efts_set_weight": (c_i32, [c_void_p, c_char_p, c_void_p, ctypes.POINTER(c_i64), c_i32]),
    "efts_finalize_weights": (c_i32, [c_void_p]),
    "efts_workspace_bytes": (c_size_t, [c_void_p, c_i32, c_i32, c_i32]),
    "efts_forward": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "efts_inference_phase1": (c_i32, [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "efts_inference_phase2": (c_i32, [c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "efts_inference_batch_phase1": (c_i32, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p,
                                            c_size_t, c_void_p]),
    "efts_inference_batch_phase2": (c_i32, [c_void_p, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_void_p]),
    "efts_conv_stack_fwd": (c_i32, [c_void_p, c_i32, c_void_p, c_void_p, c_i32, c_i32, c_void_p,
                                    c_size_t, c_void_p]),
    "efts_duration_predictor_fwd": (c_i32, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p,
                                            c_void_p, c_size_t, c_void_p]),
    "efts_length_regulator_plan": (c_i32, [c_void_p, c_void_p, c_f32, c_i32, c_i32, c_void_p, c_void_p,
                                           c_void_p, c_void_p]),
    "efts_length_regulator_fwd": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32,
                                          c_i64, c_f32, c_void_p, c_void_p, c_void_p]),
    "efts_tap_gemm": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32,
                              c_i32, c_i32, c_void_p, c_size_t, c_void_p]),
    "efts_alignment_fwd": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32,
                                   c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "efts_mask_lengths": (c_i32, [c_void_p, c_i32, c_i32, c_void_p, c_void_p]),
    "efts_index_vector": (c_i32, [c_void_p, c_i32, c_i32, c_void_p, c_void_p]),
    "efts_attention_alpha": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "efts_imv_generator": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "efts_aligned_positions": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_f32, c_void_p,
                                       c_void_p]),
    "efts_reconstruct_alignment": (c_i32, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_f32, c_void_p,
                                           c_void_p]),
    "efts_vocoder_create": (c_i32, [ctypes.POINTER(EftsVocoderConfig), ctypes.POINTER(c_void_p)]),
    "efts_vocoder_finalize": (c_i32, [c_void_p]),
    "efts_vocoder_workspace_bytes": (c_size_t, [c_void_p, c_i32, c_i32]),
    "efts_vocoder_forward": (c_i32, [c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "efts_frontend_create": (c_i32, [ctypes.POINTER(EftsFrontendConfig), ctypes.POINTER(c_void_p)]),
    "efts_frontend_finalize": (c_i32, [c_void_p]),
    "efts_frontend_frames": (c_i32, [c_void_p, c_i64]),
    "efts_frontend_workspace_bytes": (c_size_t, [c_void_p, c_i32, c_i32]),
    "efts_frontend_forward": (c_i32, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "efts_resconv_train_workspace_bytes": (c_size_t, [c_void_p, c_i32, c_i32, c_i32]),
    "efts_resconv_train_fwd": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "efts_resconv_train_bwd": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "efts_duration_train_workspace_bytes": (c_size_t, [c_void_p, c_i32, c_i32, c_i32]),
    "efts_duration_train_fwd": (c_i32, [c_void_p] + [c_void_p] * 9 + [c_i32] * 4 + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "efts_duration_train_bwd": (c_i32, [c_void_p] + [c_void_p] * 8 + [c_i32] * 4 + [c_void_p] * 8 + [c_size_t, c_void_p]),
    "efts_fastspeech_loss_workspace_bytes": (c_size_t, [c_void_p]),
    "efts_fastspeech_loss": (c_i32, [c_void_p] + [c_void_p] * 6 + [c_i32] * 6 + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "efts_scale_by_scalar": (c_i32, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "efts_host_map_transposed": (c_i32, [c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "efts_host_map_grouped": (c_i32, [c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p, ctypes.POINTER(c_i32)]),
    "efts_set_option": (c_i32, [c_void_p, c_char_p, c_i32]),
    "efts_launch_count": (c_i64, [c_void_p]),
    "efts_read_words": (c_i32, [c_void_p, c_void_p, ctypes.POINTER(c_i32), c_i32, c_void_p]),
    "efts_error_flags": (c_i32, [c_void_p, c_void_p, ctypes.POINTER(c_i32)]),
    "efts_profile_enable": (c_i32, [c_void_p, ctypes.c_uint32]),
    "efts_profile_read": (c_i32, [c_void_p, c_i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64)]),
    "efts_profile_stack_trace": (c_i32, [c_void_p, ctypes.POINTER(c_i64), c_i32]),
    "efts_profile_kernel_name": (c_i32, [c_void_p, c_i32, c_char_p, c_size_t]),
    "efts_last_error": (c_char_p, []),
    "efts_version": (c_char_p, []),
}

_lock = threading.Lock()
_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first when the sources are newer and nvcc exists) and type the library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("EFTS_B200_LIB") or _build.LIB_PATH     # A/B experiments: alternate build
        if path == _build.LIB_PATH and _build.is_stale():
            try:
                _build.build_library()
            except Exception as exc:  # no nvcc on this machine: use a prebuilt library if present
                if not os.path.exists(path):
                    raise RuntimeError(
                        "efts_b200: libefts_b200.so is missing and could not be built; there is no "
                        "fallback path (%s)" % exc)
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error():
    return load().efts_last_error().decode("utf-8", "replace")


class EftsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("efts_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise EftsError(rc, last_error())
