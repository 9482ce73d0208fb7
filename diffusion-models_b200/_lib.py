"""ctypes binding of libddm_b200.so (C ABI in include/ddm_b200.h) and its in-tree nvcc build.

The library is the only compute path of this package.  There is no PyTorch / CPU fallback: if the shared object is
missing, or `ddm_init` fails (no B200), every kernel entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libddm_b200.so")
SOURCES = ["conv_tc.cu", "small_kernels.cu", "attention.cu", "attention_tc.cu", "linattn_tc.cu", "linattn_fused.cu", "stem_tc.cu", "stem_umma.cu", "api.cu"]
HEADERS = ["conv_tc.cuh", "kernels.cuh", "ptx.cuh", os.path.join("..", "..", "include", "ddm_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]

MAX_TAPS = 9
ABI_VERSION = 2          # DDM_ABI_VERSION of include/ddm_b200.h


class ConvArgs(C.Structure):
    """Mirror of `ddm_conv_args` (include/ddm_b200.h)."""
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p),
        ("C0", C.c_int), ("C1", C.c_int), ("ld0", C.c_int), ("ld1", C.c_int),
        ("view", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("ntaps", C.c_int),
        ("tap_dy", C.c_int * MAX_TAPS), ("tap_dx", C.c_int * MAX_TAPS), ("tap_p", C.c_int * MAX_TAPS),
        ("weight", C.c_void_p),
        ("N", C.c_int), ("N_pad", C.c_int), ("K_pad", C.c_int),
        ("row_scale", C.c_void_p), ("bias", C.c_void_p), ("norm_g", C.c_void_p), ("scale_shift", C.c_void_p),
        ("ss_stride", C.c_longlong),
        ("act", C.c_int),
        ("residual", C.c_void_p), ("ld_res", C.c_int),
        ("out", C.c_void_p), ("out_f32_nchw", C.c_int), ("ld_out", C.c_int),
        ("OH", C.c_int), ("OW", C.c_int), ("oy", C.c_int), ("ox", C.c_int), ("sy", C.c_int), ("sx", C.c_int),
        ("rnorm_out", C.c_void_p),
        ("rsrc0", C.c_void_p), ("rsrc1", C.c_void_p),
        ("rC0", C.c_int), ("rC1", C.c_int), ("rld0", C.c_int), ("rld1", C.c_int),
        ("rbias", C.c_void_p),
        ("ksplit", C.c_int), ("partial_out", C.c_void_p),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p), ("head_n", C.c_int),
    ]


class LinAttnBlockArgs(C.Structure):
    """Mirror of `ddm_linattn_block_args` (include/ddm_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("out", C.c_void_p),
        ("B", C.c_int), ("n", C.c_int), ("C", C.c_int),
        ("w_qkv", C.c_void_p), ("w_out", C.c_void_p), ("bias_out", C.c_void_p), ("g_out", C.c_void_p),
        ("mem_kv", C.c_void_p), ("k_shift", C.c_void_p),
        ("heads", C.c_int), ("dim_head", C.c_int), ("n_mem", C.c_int),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "ddm_abi_version": (C.c_int, []),
    "ddm_init": (C.c_int, [C.c_int]),
    "ddm_error_string": (C.c_char_p, [C.c_int]),
    "ddm_launch_count": (C.c_longlong, []),
    "ddm_conv2d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "ddm_conv2d_shortcut_supported": (C.c_int, [C.c_int] * 6),
    "ddm_conv2d_suggest_ksplit": (C.c_int, [C.c_longlong, C.c_int, C.c_int]),
    "ddm_conv2d_row_norm_supported": (C.c_int, [C.c_int]),
    "ddm_conv2d_head_supported": (C.c_int, [C.c_int] * 4),
    "ddm_rmsnorm_act_split": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "ddm_debug_conv_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "ddm_stem_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ddm_head_conv1x1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p]),
    "ddm_sinusoidal_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "ddm_small_linear": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ddm_row_rnorm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "ddm_rmsnorm_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "ddm_groupnorm_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.c_void_p]),
    "ddm_linear_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]),
    "ddm_linear_attention_bounded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p]),
    "ddm_linear_attention_block_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ddm_linear_attention_block": (C.c_int, [C.POINTER(LinAttnBlockArgs), C.c_void_p]),
    "ddm_debug_linattn_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "ddm_attention": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ddm_sampler_step": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_void_p]),
    "ddm_sampler_step_learned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_ulonglong, C.c_longlong, C.c_longlong, C.c_void_p]),
    "ddm_sampler_step_guided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_ulonglong, C.c_longlong, C.c_void_p]),
    "ddm_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "ddm_select_row": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ddm_randn": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_longlong, C.c_void_p]),
}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libddm_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if not force and os.path.exists(LIB_PATH):
        lib_m = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(d) <= lib_m for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    dbg = ["-DDDM_CONV_DEBUG_BUILD"] if os.environ.get("DDM_CONV_DEBUG_BUILD") else []     # in-kernel bisection / trace switches
    cmd = [nvcc] + NVCC_FLAGS + dbg + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class DdmError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None
_initialised = set()


def load() -> C.CDLL:
    """dlopen the library and type every export (no device needed).  Raises if the .so has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DdmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(this package has no fallback path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in EXPORTS.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.ddm_abi_version() != ABI_VERSION:
                raise DdmError("libddm_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().ddm_error_string(code).decode()
        raise DdmError(f"{what or 'ddm call'} failed with code {code}: {msg}")


def init(device_index: int) -> C.CDLL:
    """Load + ddm_init(device).  Fails loudly without a B200."""
    lib = load()
    if device_index not in _initialised:
        check(lib.ddm_init(device_index), "ddm_init")
        _initialised.add(device_index)
    return lib


def launch_count() -> int:
    return int(load().ddm_launch_count())
