"""The fast preprocessing path (csrc/preprocess.cu: normalize_im2col16_kernel) replaces ToTensor + Normalize by ONE fp32
fma per byte, round16(fma(x, A_c, C_c)).  This CPU test pins the constants in the source to the oracle: for all 256 x 3
inputs, fp16 and bf16, the fma form rounds to the same 16-bit value as the reference's divide / subtract / divide chain
(oracle/clip_oracle.py:289-302, data/clip_transforms.py:50-56)."""
import re
from pathlib import Path

import numpy as np

from oracle import clip_oracle as O

SRC = Path(__file__).resolve().parent.parent / "aihab_clip_b200" / "csrc" / "preprocess.cu"


def _consts(name):
    m = re.search(r"constexpr float %s\[3\] = \{([^}]*)\};" % name, SRC.read_text())
    assert m, f"{name} not found in {SRC}"
    vals = [np.float32(float(t.strip().rstrip("f"))) for t in m.group(1).split(",")]
    assert len(vals) == 3
    return vals


def _to_bf16_bits(x32):
    u = x32.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)     # round to nearest even


def test_one_fma_normalisation_equals_the_reference_chain_for_every_byte():
    A, C = _consts("kA"), _consts("kC")
    b = np.arange(256, dtype=np.float32)
    for c in range(3):
        ref = ((b / np.float32(255.0)) - np.float32(O.CLIP_MEAN[c])) / np.float32(O.CLIP_STD[c])   # fp32 ops, as torch evaluates them
        assert ref.dtype == np.float32
        # fma with a single rounding: the product of an 8-bit integer and a 24-bit significand and the sum are exact in fp64
        fma = (b.astype(np.float64) * np.float64(A[c]) + np.float64(C[c])).astype(np.float32)
        np.testing.assert_array_equal(fma.astype(np.float16).view(np.uint16), ref.astype(np.float16).view(np.uint16))
        np.testing.assert_array_equal(_to_bf16_bits(fma), _to_bf16_bits(ref))
        # and the constants are the correctly rounded 1 / (255 std) and -mean / std
        assert A[c] == np.float32(1.0 / (255.0 * float(np.float32(O.CLIP_STD[c]))))
        assert C[c] == np.float32(-float(np.float32(O.CLIP_MEAN[c])) / float(np.float32(O.CLIP_STD[c])))
