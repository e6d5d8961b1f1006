"""Prompt-ensembled class text embeddings — mirror of ``clip_classifier`` (reference utils.py:31-57)."""
from __future__ import annotations

import torch


def clip_classifier(classnames, template, clip_model, tokenize=None):
    """Per class: tokenize the T prompts, ``encode_text``, L2-normalise, mean over templates, renormalise.
    Returns ``(texts [C,77], clip_weights_before [T,C,W_txt], clip_weights [E,C])`` exactly like the reference."""
    if tokenize is None:
        from .clip import tokenize
    with torch.no_grad():
        device = next(clip_model.parameters()).device
        weights, before, first_tokens = [], [], []
        for classname in classnames:
            texts = tokenize([t.format(classname.replace("_", " ")) for t in template]).to(device)
            x_before, emb = clip_model.encode_text(texts)
            emb = emb / emb.norm(dim=-1, keepdim=True)
            mean = emb.mean(dim=0)
            weights.append(mean / mean.norm())
            before.append(x_before.squeeze(dim=1))
            first_tokens.append(texts[0])
        return (torch.stack(first_tokens, dim=0), torch.stack(before, dim=1).to(device),
                torch.stack(weights, dim=1).to(device))


def text_weights_from_tokens(clip_model, tokens: torch.Tensor, templates_per_class: int = 1) -> torch.Tensor:
    """Same ensemble from pre-tokenized prompts ``[C*T, 77]`` (class-major), for hosts without the BPE table."""
    with torch.no_grad():
        _, emb = clip_model.encode_text(tokens)
        emb = emb / emb.norm(dim=-1, keepdim=True)
        emb = emb.reshape(-1, templates_per_class, emb.shape[-1]).mean(dim=1)
        return (emb / emb.norm(dim=-1, keepdim=True)).t().contiguous()
