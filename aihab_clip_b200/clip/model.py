"""CLIP model object with the reference's protocol, image tower executed by the sm_100a engine.

Mirror of the reference ``clip/model.py`` for the ViT path:

* ``CLIP.encode_image`` (ref :335-336) returns the PRE-projection features ``ln_post(x[:, 0, :])`` in
  ``model.dtype`` — ``visual.proj`` is not applied inside (ref :228-235); callers project, normalise and score.
* ``CLIP.encode_text`` (ref :338-353) returns the 2-tuple ``(x_before_proj, x)``; it runs once per class set and
  stays plain PyTorch.
* parameter / state_dict names are the reference's (``visual.conv1.weight``,
  ``visual.transformer.resblocks.N.attn.in_proj_weight`` …) so checkpoints and ``state_dict["visual.proj"]``
  consumers (methods/ProLIP.py:89) keep working.

The image tower holds its parameters as ordinary ``nn.Parameter`` s but never runs a torch op on them: ``forward``
hands the raw device pointers to ``libaihab_clip.so`` (tcgen05 GEMMs, fused attention, LayerNorm kernels).  There is
no fallback: on a CPU tensor or without the built library ``encode_image`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib, ops

_COMPUTE_DTYPES = {"fp16": _lib.F16, "float16": _lib.F16, "bf16": _lib.BF16, "bfloat16": _lib.BF16}


class LayerNorm(nn.LayerNorm):
    """fp32-upcast LayerNorm (ref clip/model.py:151-157)."""

    def forward(self, x: torch.Tensor):
        return super().forward(x.float()).to(x.dtype)


class _Attn(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight/bias, out_proj.weight/bias)."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def forward(self, x: torch.Tensor, causal: bool):  # x: [N, L, D]; text tower only
        n, L, d = x.shape
        q, k, v = F.linear(x, self.in_proj_weight, self.in_proj_bias).chunk(3, dim=-1)
        shape = (n, L, self.n_head, d // self.n_head)
        q, k, v = (t.reshape(shape).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        return self.out_proj(o.transpose(1, 2).reshape(n, L, d))


class ResidualAttentionBlock(nn.Module):
    """ref clip/model.py:165-186.  ``forward`` is used by the text tower; the image tower only reads the params."""

    def __init__(self, d_model: int, n_head: int, causal: bool = False):
        super().__init__()
        self.attn = _Attn(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.causal = causal

    def forward(self, x: torch.Tensor):
        x = x + self.attn(self.ln_1(x), self.causal)
        h = self.mlp.c_fc(self.ln_2(x))
        return x + self.mlp.c_proj(h * torch.sigmoid(1.702 * h))  # QuickGELU, ref :160-162


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, causal: bool = False):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, causal) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


class _Engine:
    """Owns one aihab_vit handle (packed 16-bit weights, TMA descriptors, workspace) for a parameter snapshot."""

    def __init__(self, vt: "VisionTransformer", device: torch.device, dtype_code: int, max_batch: int):
        lib = _lib.load()
        keep = []

        def f32(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return C.c_void_p(t.data_ptr())

        blocks = (_lib.BlockWeights * vt.transformer.layers)()
        for i, b in enumerate(vt.transformer.resblocks):
            blocks[i] = _lib.BlockWeights(
                f32(b.ln_1.weight), f32(b.ln_1.bias), f32(b.attn.in_proj_weight), f32(b.attn.in_proj_bias),
                f32(b.attn.out_proj.weight), f32(b.attn.out_proj.bias), f32(b.ln_2.weight), f32(b.ln_2.bias),
                f32(b.mlp.c_fc.weight), f32(b.mlp.c_fc.bias), f32(b.mlp.c_proj.weight), f32(b.mlp.c_proj.bias))
        w = _lib.VitWeights(f32(vt.conv1.weight), f32(vt.class_embedding), f32(vt.positional_embedding),
                            f32(vt.ln_pre.weight), f32(vt.ln_pre.bias), f32(vt.ln_post.weight), f32(vt.ln_post.bias),
                            blocks)
        cfg = _lib.VitConfig(vt.input_resolution, vt.patch_size, vt.width, vt.transformer.layers, vt.heads,
                             dtype_code, max_batch)
        handle = C.c_void_p()
        torch.cuda.synchronize(device)
        _lib.check(lib.aihab_vit_create(C.byref(cfg), C.byref(w), device.index, C.byref(handle)), "aihab_vit_create")
        self._lib, self.handle, self.device = lib, handle, device
        del keep  # the handle owns packed copies

    def close(self):
        if getattr(self, "handle", None):
            self._lib.aihab_vit_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter teardown order
        try:
            self.close()
        except Exception:
            pass


class _TextEngine:
    """Owns one aihab_text handle: the text transformer of a CLIP model on the tensor-core kernels (causal mask)."""

    def __init__(self, clip: "CLIP", device: torch.device, dtype_code: int, max_batch: int):
        lib = _lib.load()
        keep = []

        def f32(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return C.c_void_p(t.data_ptr())

        tr = clip.transformer
        blocks = (_lib.BlockWeights * tr.layers)()
        for i, b in enumerate(tr.resblocks):
            blocks[i] = _lib.BlockWeights(
                f32(b.ln_1.weight), f32(b.ln_1.bias), f32(b.attn.in_proj_weight), f32(b.attn.in_proj_bias),
                f32(b.attn.out_proj.weight), f32(b.attn.out_proj.bias), f32(b.ln_2.weight), f32(b.ln_2.bias),
                f32(b.mlp.c_fc.weight), f32(b.mlp.c_fc.bias), f32(b.mlp.c_proj.weight), f32(b.mlp.c_proj.bias))
        w = _lib.TextWeights(f32(clip.token_embedding.weight), f32(clip.positional_embedding),
                             f32(clip.ln_final.weight), f32(clip.ln_final.bias), blocks)
        cfg = _lib.TextConfig(clip.context_length, clip.vocab_size, tr.width, tr.layers, tr.width // 64, dtype_code,
                              max_batch)
        handle = C.c_void_p()
        torch.cuda.synchronize(device)
        _lib.check(lib.aihab_text_create(C.byref(cfg), C.byref(w), device.index, C.byref(handle)), "aihab_text_create")
        self._lib, self.handle, self.device, self.width = lib, handle, device, tr.width
        del keep  # the handle owns packed copies

    def encode(self, tokens: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
        tokens = tokens.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(tokens.shape[0], self.width, dtype=out_dtype, device=self.device)
        rc = self._lib.aihab_text_encode(self.handle, C.c_void_p(tokens.data_ptr()), tokens.shape[0],
                                         C.c_void_p(out.data_ptr()), ops.dtype_code(out_dtype),
                                         C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, "aihab_text_encode")
        return out

    def close(self):
        if getattr(self, "handle", None):
            self._lib.aihab_text_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - interpreter teardown order
        try:
            self.close()
        except Exception:
            pass


class VisionTransformer(nn.Module):
    """ref clip/model.py:199-235.  Same constructor, same parameter names; forward runs in libaihab_clip.so."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution, self.patch_size, self.width = input_resolution, patch_size, width
        self.heads, self.output_dim = heads, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        # engine knobs (not part of the reference protocol; defaults keep reference behaviour)
        self.compute_dtype = os.environ.get("AIHAB_CLIP_COMPUTE_DTYPE", "fp16")
        self.max_batch = int(os.environ.get("AIHAB_CLIP_MAX_BATCH", "256"))
        self._engine = None
        self._engine_key = None

    # -- engine lifecycle: rebuilt lazily whenever a parameter was replaced, moved, cast or written in place.
    # `visual.proj` is NOT part of the engine (encode_image returns pre-projection features, ref :228-235), so
    # projector updates (ProLIP training) never trigger a rebuild.  In-place writes through `.data`
    # (p.data.copy_(), EMA loops, custom loaders) do not bump `_version`: call invalidate_engine() after them.
    def _engine_params(self):
        ps = self.__dict__.get("_engine_param_list")
        if ps is None:
            ps = [p for n, p in self.named_parameters() if n != "proj"]
            self.__dict__["_engine_param_list"] = ps
        return ps

    def _fingerprint(self, device):
        return (device, self.compute_dtype, self.max_batch,
                tuple((p.data_ptr(), p._version) for p in self._engine_params()))

    def invalidate_engine(self):
        """Drop the packed 16-bit weight copy; the next forward repacks from the current parameters."""
        if self._engine is not None:
            self._engine.close()
        self._engine, self._engine_key = None, None
        self.__dict__.pop("_engine_param_list", None)

    def _apply(self, fn, *args, **kwargs):   # .to() / .float() / .half() / .cuda(): parameters are replaced
        out = super()._apply(fn, *args, **kwargs)
        self.__dict__.pop("_engine_param_list", None)
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_engine()
        return out

    def __getstate__(self):   # pickling / copy.deepcopy: the ctypes handle stays behind, the copy builds its own
        state = self.__dict__.copy()
        state["_engine"], state["_engine_key"] = None, None
        state.pop("_engine_param_list", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def engine(self, device: torch.device) -> _Engine:
        key = self._fingerprint(device)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            try:
                code = _COMPUTE_DTYPES[self.compute_dtype]
            except KeyError:
                raise ValueError(f"compute_dtype must be one of {sorted(_COMPUTE_DTYPES)}") from None
            if code == _lib.BF16 and not getattr(self, "_bf16_warned", False):
                import warnings
                warnings.warn("compute_dtype='bf16': single-pass bf16 tensor-core operands are OUTSIDE the parity "
                              "tolerance of this path (measured max |dlogit| 3.9e-2 vs the 1e-2 gate, argmax agreement "
                              "99.41 % vs 99.9 %); use the default 'fp16' (same tcgen05 rate) unless that is acceptable",
                              stacklevel=3)
                self._bf16_warned = True
            self._engine = _Engine(self, device, code, self.max_batch)
            self._engine_key = key
        return self._engine

    def preferred_batch(self, device: torch.device, max_batch: int | None = None) -> int:
        """Images per encode call (<= max_batch) whose GEMMs fill the device's SMs with whole waves of tiles; the
        batch size to use for extraction loops (the reference uses a fixed 16, methods/utils.py:142-173)."""
        tokens = (self.input_resolution // self.conv1.kernel_size[0]) ** 2 + 1
        idx = device.index if device.index is not None else torch.cuda.current_device()
        return int(_lib.load().aihab_preferred_batch(tokens, self.conv1.out_channels, int(max_batch or self.max_batch), idx))

    def _check_device(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("aihab_clip_b200 image tower runs on a B200 CUDA device only (no CPU fallback); "
                               "move the model and the images to 'cuda'")
        if self.conv1.weight.device != x.device:
            raise RuntimeError(f"Input is on {x.device} but the model is on {self.conv1.weight.device}")

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N,3,R,R] float (already preprocessed) -> [N, width] pre-projection features in x.dtype."""
        self._check_device(x)
        R = self.input_resolution
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, R, R):
            raise RuntimeError(f"expected images of shape [N, 3, {R}, {R}], got {tuple(x.shape)}")
        x = x.contiguous()
        eng = self.engine(x.device)
        out = torch.empty(x.shape[0], self.width, dtype=x.dtype, device=x.device)
        rc = eng._lib.aihab_vit_encode(eng.handle, C.c_void_p(x.data_ptr()), ops.dtype_code(x.dtype), x.shape[0],
                                       C.c_void_p(out.data_ptr()), ops.dtype_code(out.dtype),
                                       C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _lib.check(rc, "aihab_vit_encode")
        return out

    @torch.no_grad()
    def forward_u8(self, images_u8: torch.Tensor, out_dtype: torch.dtype | None = None) -> torch.Tensor:
        """uint8 [N,H,W,3] -> features, with the eval preprocessing (data/clip_transforms.py:50-56) fused in front."""
        self._check_device(images_u8)
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise RuntimeError("expected uint8 images of shape [N, H, W, 3]")
        images_u8 = images_u8.contiguous()
        eng = self.engine(images_u8.device)
        out_dtype = out_dtype or self.conv1.weight.dtype
        n, sh, sw, _ = images_u8.shape
        out = torch.empty(n, self.width, dtype=out_dtype, device=images_u8.device)
        rc = eng._lib.aihab_vit_encode_u8(eng.handle, C.c_void_p(images_u8.data_ptr()), n, sh, sw,
                                          C.c_void_p(out.data_ptr()), ops.dtype_code(out_dtype),
                                          C.c_void_p(torch.cuda.current_stream(images_u8.device).cuda_stream))
        _lib.check(rc, "aihab_vit_encode_u8")
        return out


class CLIP(nn.Module):
    """ref clip/model.py:238-353 (ViT image tower only; a tuple ``vision_layers`` selects the out-of-scope ResNet)."""

    def __init__(self, embed_dim: int, image_resolution: int, vision_layers, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int, transformer_width: int,
                 transformer_heads: int, transformer_layers: int):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError("ModifiedResNet image towers (RN50/RN101) are outside the B200 hot path; "
                                      "only ViT backbones are supported")
        self.context_length = context_length
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, causal=True)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.659260036932778)  # log(1 / 0.07)
        # engine knob (not part of the reference protocol): "torch" keeps the text tower in PyTorch at the model dtype
        # (the default: it runs once per class set), "b200" runs it on the tensor-core kernels (16-bit operands) for
        # large prompt ensembles
        self.text_engine = os.environ.get("AIHAB_CLIP_TEXT_ENGINE", "torch")
        self.text_max_batch = int(os.environ.get("AIHAB_CLIP_TEXT_MAX_BATCH", "512"))
        self._tengine = None
        self._tengine_key = None
        self._init_text()

    def _init_text(self):  # ref :294-321
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, n = self.transformer.width, self.transformer.layers
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=w ** -0.5)
            nn.init.normal_(blk.attn.out_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
            nn.init.normal_(blk.mlp.c_fc.weight, std=(2 * w) ** -0.5)
            nn.init.normal_(blk.mlp.c_proj.weight, std=(w ** -0.5) * ((2 * n) ** -0.5))
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.visual(image.type(self.dtype))

    def encode_image_u8(self, images_u8: torch.Tensor) -> torch.Tensor:
        """Extension: raw uint8 HWC batch -> features (GPU preprocessing fused in front of the tower)."""
        return self.visual.forward_u8(images_u8, self.dtype)

    def _text_engine(self, device: torch.device) -> _TextEngine:
        params = [self.token_embedding.weight, self.positional_embedding, self.ln_final.weight, self.ln_final.bias]
        params += list(self.transformer.parameters())
        key = (device, self.visual.compute_dtype, self.text_max_batch, tuple((p.data_ptr(), p._version) for p in params))
        if self._tengine is None or self._tengine_key != key:
            if self._tengine is not None:
                self._tengine.close()
            self._tengine = _TextEngine(self, device, _COMPUTE_DTYPES[self.visual.compute_dtype], self.text_max_batch)
            self._tengine_key = key
        return self._tengine

    def encode_text(self, text: torch.Tensor):
        if self.text_engine == "b200":  # ref :338-353 on libaihab_clip.so; same return contract
            if not text.is_cuda:
                raise RuntimeError("text_engine='b200' runs on a CUDA device only (no CPU fallback); move the tokens "
                                   "to 'cuda' or use text_engine='torch'")
            with torch.no_grad():
                x_before_proj = self._text_engine(text.device).encode(text, self.dtype)
            return x_before_proj, x_before_proj @ self.text_projection
        x = self.token_embedding(text).type(self.dtype)
        x = x + self.positional_embedding.type(self.dtype)
        x = self.transformer(x)
        x = self.ln_final(x).type(self.dtype)
        x_before_proj = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]  # EOT token has the highest id
        return x_before_proj, x_before_proj @ self.text_projection

    def forward(self, image, text):
        """The reference's CLIP.forward (ref :355-369) is broken in this fork (it calls .norm on the tuple returned
        by encode_text) and is never called; raise instead of guessing a behaviour."""
        raise NotImplementedError("CLIP.forward is unused in aihab-clip; call encode_image / encode_text")


def convert_weights(model: nn.Module):
    """ref clip/model.py:372-393 — cast Conv/Linear/attention weights and the projections to fp16; LayerNorm
    parameters, class/positional/token embeddings and logit_scale stay fp32."""
    def to_half(m):
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            m.weight.data = m.weight.data.half()
            if m.bias is not None:
                m.bias.data = m.bias.data.half()
        if isinstance(m, _Attn):
            m.in_proj_weight.data = m.in_proj_weight.data.half()
            m.in_proj_bias.data = m.in_proj_bias.data.half()
        for name in ("text_projection", "proj"):
            attr = getattr(m, name, None)
            if isinstance(attr, torch.Tensor):
                attr.data = attr.data.half()
    model.apply(to_half)


def build_model(state_dict: dict) -> CLIP:
    """ref clip/model.py:396-433 — infer the geometry from tensor shapes, build, cast, load, eval."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("ResNet CLIP checkpoints (no 'visual.proj') are outside the B200 hot path")
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    embed_dim = state_dict["text_projection"].shape[1]
    transformer_width = state_dict["ln_final.weight"].shape[0]
    transformer_layers = len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")})
    model = CLIP(embed_dim, vision_patch_size * grid_size, vision_layers, vision_width, vision_patch_size,
                 state_dict["positional_embedding"].shape[0], state_dict["token_embedding.weight"].shape[0],
                 transformer_width, transformer_width // 64, transformer_layers)
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    convert_weights(model)
    model.load_state_dict(sd)
    return model.eval()
