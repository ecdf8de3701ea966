"""ctypes binding of libduoformer_sm100.so (the C ABI declared in include/duoformer_sm100.h).

The product path has no CPU / eager fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libduoformer_sm100.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

# enums of include/duoformer_sm100.h
EPI_BF16, EPI_GELU_BF16, EPI_RESIDUAL_F32, EPI_SCATTER_F32, EPI_F32, EPI_SPLIT_BF16, EPI_GELU_SPLIT_BF16 = range(7)
ACT_BF16, ACT_SPLIT, ACT_F32, ACT_F16 = range(4)

# every symbol include/duoformer_sm100.h declares
ABI_VERSION = 4  # duo_abi_version() of the library this binding was written against

EXPORTED_SYMBOLS = (
    "duo_last_error",
    "duo_abi_version",
    "duo_launch_count",
    "duo_launch_count_reset",
    "duo_gemm",
    "duo_layernorm",
    "duo_group_attention",
    "duo_fill_scale_token",
    "duo_add_pos",
    "duo_assemble_patch_tokens",
    "duo_head",
    "duo_convert",
    "duo_im2col3x3",
    "duo_pool_to_slice",
    "duo_maxpool3x3s2",
    "duo_conv2d",
    "duo_stem_pack",
    "duo_stem_conv7x7",
)


class GemmArgs(Structure):
    """struct duo_gemm_args"""

    _fields_ = [
        ("A", c_void_p),
        ("W", c_void_p),
        ("bias", c_void_p),
        ("out", c_void_p),
        ("gamma", c_void_p),
        ("row_map", c_void_p),
        ("pos", c_void_p),
        ("M", c_int64),
        ("lda", c_int64),
        ("ldw", c_int64),
        ("ldo", c_int64),
        ("N", c_int32),
        ("K", c_int32),
        ("split3", c_int32),
        ("epilogue", c_int32),
        ("rows_per_group", c_int32),
        ("dest_rows_per_group", c_int32),
        ("pos_period", c_int32),
        ("ln_eps", c_float),
        ("relu", c_int32),
        ("fp16_operands", c_int32),
        ("xb_out", c_void_p),
        ("stats_out", c_void_p),
        ("ln_stats", c_void_p),
        ("shift_stats", c_void_p),
    ]


class Conv2dArgs(Structure):
    """struct duo_conv2d_args"""

    _fields_ = [
        ("inp", c_void_p),
        ("weight", c_void_p),
        ("bias", c_void_p),
        ("residual", c_void_p),
        ("out", c_void_p),
        ("B", c_int32),
        ("H", c_int32),
        ("W", c_int32),
        ("Cin", c_int32),
        ("Cout", c_int32),
        ("ksize", c_int32),
        ("stride", c_int32),
        ("relu", c_int32),
        ("fp16", c_int32),
        ("out_fp16", c_int32),
        ("in2", c_void_p),
        ("H2", c_int32),
        ("W2", c_int32),
        ("Cin2", c_int32),
        ("stride2", c_int32),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building libduoformer_sm100.so failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and declare the prototypes.  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the DuoFormer sm_100a path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.duo_last_error.restype = c_char_p
    lib.duo_last_error.argtypes = []
    lib.duo_abi_version.restype = c_int32
    if lib.duo_abi_version() != ABI_VERSION:  # a stale build: struct layouts would not match
        raise RuntimeError(f"{LIB_PATH} has ABI version {lib.duo_abi_version()}, this package needs {ABI_VERSION}: "
                           "rebuild with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib.duo_launch_count.restype = c_int64
    lib.duo_launch_count_reset.restype = None
    lib.duo_gemm.restype = c_int32
    lib.duo_gemm.argtypes = [POINTER(GemmArgs), c_void_p]
    lib.duo_layernorm.restype = c_int32
    lib.duo_layernorm.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int64, c_float, c_void_p, c_void_p]
    lib.duo_group_attention.restype = c_int32
    lib.duo_group_attention.argtypes = [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, c_int32, c_float, c_int32, c_int32, c_void_p]
    lib.duo_fill_scale_token.restype = c_int32
    lib.duo_fill_scale_token.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_add_pos.restype = c_int32
    lib.duo_add_pos.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]
    lib.duo_assemble_patch_tokens.restype = c_int32
    lib.duo_assemble_patch_tokens.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_head.restype = c_int32
    lib.duo_head.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_convert.restype = c_int32
    lib.duo_convert.argtypes = [c_void_p, c_int64, c_void_p, c_int32, c_int64, c_int32, c_void_p]
    lib.duo_im2col3x3.restype = c_int32
    lib.duo_im2col3x3.argtypes = [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_pool_to_slice.restype = c_int32
    lib.duo_pool_to_slice.argtypes = [c_void_p, c_int32, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_maxpool3x3s2.restype = c_int32
    lib.duo_maxpool3x3s2.argtypes = [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_conv2d.restype = c_int32
    lib.duo_conv2d.argtypes = [POINTER(Conv2dArgs), c_void_p]
    lib.duo_stem_pack.restype = c_int32
    lib.duo_stem_pack.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.duo_stem_conv7x7.restype = c_int32
    lib.duo_stem_conv7x7.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().duo_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
