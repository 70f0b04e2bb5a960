"""ctypes binding of the C ABI in include/mdt_b200.h (libmdt_b200.so, built from csrc/ for sm_100a)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdt_b200.so")
MDT_ABI_VERSION = 1
MDT_MAX_LEVELS = 4
PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2, "fp16": 3}

EXPORTS = [
    "mdt_last_error", "mdt_abi_version", "mdt_device_count", "mdt_adpm2_scalars", "mdt_aeuler_scalars", "mdt_karras_sigmas",
    "mdt_plan_create", "mdt_plan_destroy", "mdt_plan_device_bytes", "mdt_plan_launch_count", "mdt_plan_sample",
    "mdt_plan_inpaint", "mdt_plan_unet_forward", "mdt_plan_enable_taps", "mdt_plan_read_tap", "mdt_op_linear", "mdt_op_step_update",
    "mdt_op_decode_tokens", "mdt_plan_set_context_mode", "mdt_plan_set_sampler_mode",
]


class MdtConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("length", C.c_int32),
        ("channels", C.c_int32), ("patch_size", C.c_int32), ("num_levels", C.c_int32),
        ("multipliers", C.c_int32 * (MDT_MAX_LEVELS + 1)), ("factors", C.c_int32 * MDT_MAX_LEVELS),
        ("num_blocks", C.c_int32 * MDT_MAX_LEVELS), ("attentions", C.c_int32 * (MDT_MAX_LEVELS + 1)),
        ("pre_transformer", C.c_int32), ("heads", C.c_int32), ("head_features", C.c_int32),
        ("ff_multiplier", C.c_int32), ("resnet_groups", C.c_int32), ("kernel_multiplier_downsample", C.c_int32),
        ("use_skip_scale", C.c_int32), ("mapping_features", C.c_int32), ("ctx_features", C.c_int32),
        ("ctx_max_length", C.c_int32), ("text_embed_dim", C.c_int32), ("embed_dim_position", C.c_int32),
        ("pos_emb_fourier", C.c_int32), ("pos_emb_fourier_add", C.c_int32), ("sigma_data", C.c_float),
        ("precision", C.c_int32), ("max_batch", C.c_int32), ("max_timesteps", C.c_int32),
    ]


class MdtTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class MdtIterScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "sigma", "c_in_a", "c_noise_a", "c_skip_a", "c_out_a", "sigma_mid", "c_in_b", "c_noise_b", "c_skip_b",
        "c_out_b", "dt_mid", "dt_down", "sigma_up")]


class MdtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mdt_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the extension; raises loudly if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C moleculediffusiontransformer_b200/csrc`. There is no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
    lib.mdt_last_error.restype = C.c_char_p
    lib.mdt_abi_version.restype = C.c_int
    lib.mdt_device_count.restype = C.c_int
    lib.mdt_adpm2_scalars.argtypes = [vp, C.c_int, f64, f64, C.POINTER(MdtIterScalars)]
    lib.mdt_aeuler_scalars.argtypes = [vp, C.c_int, f64, C.POINTER(MdtIterScalars)]
    lib.mdt_karras_sigmas.argtypes = [C.c_int, f64, f64, f64, vp]
    lib.mdt_plan_create.argtypes = [C.POINTER(MdtConfig), C.POINTER(MdtTensor), i64, C.c_int, C.POINTER(vp)]
    lib.mdt_plan_destroy.argtypes = [vp]
    lib.mdt_plan_destroy.restype = None
    lib.mdt_plan_device_bytes.argtypes = [vp]
    lib.mdt_plan_device_bytes.restype = i64
    lib.mdt_plan_launch_count.argtypes = [vp]
    lib.mdt_plan_launch_count.restype = i64
    lib.mdt_plan_sample.argtypes = [vp, vp, i32, vp, vp, vp, i32, u64, u64, i64, f32, i32, vp, vp, vp]
    lib.mdt_plan_inpaint.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, u64, u64, i64, f32, vp, vp]
    lib.mdt_plan_unet_forward.argtypes = [vp, vp, f32, vp, i32, i64, f32, vp, vp]
    lib.mdt_plan_set_context_mode.argtypes = [vp, C.c_int]
    lib.mdt_plan_set_sampler_mode.argtypes = [vp, C.c_int, f32, f32]
    lib.mdt_plan_enable_taps.argtypes = [vp, C.c_int]
    lib.mdt_plan_read_tap.argtypes = [vp, C.c_char_p, vp, i64]
    lib.mdt_plan_read_tap.restype = i64
    lib.mdt_op_linear.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.mdt_op_step_update.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.POINTER(MdtIterScalars), f32, i64, i32, i32,
                                       C.c_int, vp]
    lib.mdt_op_decode_tokens.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    if lib.mdt_abi_version() != MDT_ABI_VERSION:
        raise ImportError("libmdt_b200.so ABI version mismatch; rebuild the extension")
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise MdtError(rc, load().mdt_last_error().decode())
    return rc
