"""``clip.load`` / ``clip.tokenize`` drop-ins (reference: clip/clip.py:89-137, 192-228).

``load`` keeps this fork's 3-tuple return ``(state_dict, model, preprocess)`` (ref :137) and the non-JIT branch the
callers use (``jit=False``, aihab_utils/model_init.py:145).  There is no network in the target environment, so a
model *name* is resolved to an already-downloaded file under ``download_root`` instead of being fetched.
"""
from __future__ import annotations

import os
import warnings
from typing import List, Union

import torch

from .model import build_model
from .simple_tokenizer import SimpleTokenizer

__all__ = ["available_models", "load", "tokenize"]

# file names of the OpenAI releases the reference knows (clip/clip.py:29-36); ViT only on this path
_MODEL_FILES = {"ViT-B/32": "ViT-B-32.pt", "ViT-B/16": "ViT-B-16.pt"}

_tokenizer = None


def _get_tokenizer() -> SimpleTokenizer:
    global _tokenizer
    if _tokenizer is None:
        _tokenizer = SimpleTokenizer()
    return _tokenizer


def available_models() -> List[str]:
    return list(_MODEL_FILES.keys())


class ClipPreprocess:
    """``preprocess`` returned by :func:`load`: PIL.Image -> float32 [3,R,R] exactly as ref clip/clip.py:74-81
    (Resize(R, BICUBIC), CenterCrop(R), RGB, ToTensor, Normalize).  Called per image on the host (DataLoader
    workers); batches of raw uint8 arrays take the CUDA kernel instead via :meth:`batch_u8` /
    ``model.encode_image_u8``."""

    def __init__(self, n_px: int):
        from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
        self.n_px = n_px
        self.transforms = Compose([
            Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), lambda im: im.convert("RGB"),
            ToTensor(), Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])

    def __call__(self, image):
        return self.transforms(image)

    def batch_u8(self, images_u8: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        from .. import ops
        return ops.preprocess_u8(images_u8, self.n_px, dtype)

    def __repr__(self):
        return f"ClipPreprocess(n_px={self.n_px})"


def _transform(n_px: int) -> ClipPreprocess:
    return ClipPreprocess(n_px)


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu",
         jit: bool = False, download_root: str = None):
    """Load a CLIP checkpoint.  ``name`` is a key of :func:`available_models` or a path to a checkpoint holding a
    state_dict (or an OpenAI TorchScript archive, whose state_dict is extracted).  Returns
    ``(state_dict, model, preprocess)`` like the reference fork."""
    if name in _MODEL_FILES:
        model_path = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), _MODEL_FILES[name])
        if not os.path.isfile(model_path):
            raise RuntimeError(f"Model {name}: {model_path} does not exist and this build cannot download it "
                               "(no network); place the OpenAI checkpoint there or pass a checkpoint path")
    elif os.path.isfile(name):
        model_path = name
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")

    try:
        archive = torch.jit.load(model_path, map_location="cpu").eval()
        state_dict = archive.state_dict()
    except RuntimeError:
        state_dict = torch.load(model_path, map_location="cpu")
    if jit:
        warnings.warn("jit=True is not supported by the B200 engine; loading the state dict instead")

    model = build_model(state_dict).to(device)
    if str(device) == "cpu":
        model.float()
    return model.state_dict(), model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.LongTensor:
    """ref clip/clip.py:192-228 — [SOT] + BPE(text) + [EOT], zero padded to ``context_length``."""
    if isinstance(texts, str):
        texts = [texts]
    tok = _get_tokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [sot] + tok.encode(text) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        result[i, :len(ids)] = torch.tensor(ids)
    return result
