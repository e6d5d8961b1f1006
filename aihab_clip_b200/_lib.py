"""ctypes binding of libaihab_clip.so (C ABI declared in include/aihab_clip.h).

The shared library is built in-tree by ``python -m aihab_clip_b200.build`` (nvcc, sm_100a only).  There is no
fallback of any kind: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

F32, F16, BF16 = 0, 1, 2
EPI_BIAS_16, EPI_BIAS_GELU_16, EPI_BIAS_RES_32, EPI_PATCH_32, EPI_SCALE_32 = range(5)

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libaihab_clip.so"

c_float_p = C.POINTER(C.c_float)


class VitConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("image_size", "patch_size", "width", "layers", "heads", "dtype", "max_batch")]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln_1_weight", "ln_1_bias", "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
        "ln_2_weight", "ln_2_bias", "c_fc_weight", "c_fc_bias", "c_proj_weight", "c_proj_bias")]


class TextConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("context_length", "vocab_size", "width", "layers", "heads", "dtype", "max_batch")]


class TextWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("token_embedding", "positional_embedding", "ln_final_weight",
                                          "ln_final_bias")] + [("blocks", C.POINTER(BlockWeights))]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "conv1_weight", "class_embedding", "positional_embedding", "ln_pre_weight", "ln_pre_bias",
        "ln_post_weight", "ln_post_bias")] + [("blocks", C.POINTER(BlockWeights))]


# every symbol include/aihab_clip.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "aihab_abi_version": (C.c_int, []),
    "aihab_last_error": (C.c_char_p, []),
    "aihab_kernel_launches": (C.c_uint64, []),
    "aihab_profile_enable": (C.c_int, [C.c_int]),
    "aihab_profile_sites": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                      C.POINTER(C.c_uint64), C.c_int]),
    "aihab_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int]),
    "aihab_vit_create": (C.c_int, [C.POINTER(VitConfig), C.POINTER(VitWeights), C.c_int, C.POINTER(C.c_void_p)]),
    "aihab_vit_destroy": (None, [C.c_void_p]),
    "aihab_vit_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "aihab_preferred_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "aihab_vit_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aihab_vit_encode_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aihab_preprocess_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aihab_score": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float,
                              C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aihab_l2_normalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "aihab_score16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aihab_text_create": (C.c_int, [C.POINTER(TextConfig), C.POINTER(TextWeights), C.c_int, C.POINTER(C.c_void_p)]),
    "aihab_text_destroy": (None, [C.c_void_p]),
    "aihab_text_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aihab_attention_causal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aihab_l2_metrics": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aihab_prototype_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aihab_gemm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "aihab_layernorm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]),
    "aihab_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("AIHAB_CLIP_LIB", LIB_PATH))
    if not path.is_file():
        raise RuntimeError(
            f"{path} not found: the CUDA extension is not built. Run `python -m aihab_clip_b200.build` "
            "(needs nvcc). aihab_clip_b200 has no CPU or PyTorch fallback for the image tower.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if lib.aihab_abi_version() != 1:
        raise RuntimeError("libaihab_clip.so ABI version mismatch; rebuild with `python -m aihab_clip_b200.build`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().aihab_last_error()
        raise RuntimeError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def kernel_launches() -> int:
    return int(load().aihab_kernel_launches())


PROFILE_CLASSES = {"gemm": 0, "attention": 1, "layernorm": 2, "preprocess": 3, "score": 4}


def profile_enable(on: bool) -> None:
    load().aihab_profile_enable(1 if on else 0)


def profile_sites(cls: str, cap: int = 16) -> list:
    """[{'work': algorithmic FLOPs or bytes per launch, 'tag': site tag (GEMM: N), 'ms': summed event time,
    'launches': n}] per launch site of a kernel class.  Non-resetting: call before profile_read(reset=True)."""
    work, tag, ms, n = (C.c_double * cap)(), (C.c_int * cap)(), (C.c_double * cap)(), (C.c_uint64 * cap)()
    g = load().aihab_profile_sites(PROFILE_CLASSES[cls], work, tag, ms, n, cap)
    if g < 0:
        check(1, "aihab_profile_sites")
    return [{"work": work[i], "tag": int(tag[i]), "ms": ms[i], "launches": int(n[i])} for i in range(g)]


def profile_read(cls: str, reset: bool = True) -> dict:
    """{'ms': total event time, 'launches': n, 'work': algorithmic FLOPs or bytes} for one kernel class."""
    ms, n, work = C.c_double(), C.c_uint64(), C.c_double()
    check(load().aihab_profile_read(PROFILE_CLASSES[cls], C.byref(ms), C.byref(n), C.byref(work), 1 if reset else 0),
          "aihab_profile_read")
    return {"ms": ms.value, "launches": int(n.value), "work": work.value}
