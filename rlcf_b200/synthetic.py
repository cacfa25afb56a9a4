"""Synthetic CLIP weights, prompts and augmented views of the right shapes (there are no OpenAI checkpoints or
datasets offline).  Used by bench.py, __graft_entry__.smoke() and the eval driver's --synthetic mode.

Architectures follow TPT/clip/model.py:399-439 (build_model infers them from the checkpoint); the weight scales
follow CLIP.initialize_parameters (model.py:299-326) so activations have realistic magnitudes.
"""
from __future__ import annotations

import math

import torch

# name: (embed_dim, resolution, vision_layers, vision_width, patch, ctx_len, vocab, text_width, text_heads, text_layers)
ARCHS = {
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": (768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": (768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "tiny-A": (128, 64, 2, 128, 16, 77, 512, 128, 2, 2),
    "tiny-B": (256, 64, 3, 256, 8, 77, 512, 128, 2, 2),
    "tiny-C": (256, 96, 3, 256, 16, 77, 512, 128, 2, 2),
    "tiny-P": (128, 64, 2, 128, 16, 77, 49408, 128, 2, 2),
    "tiny-Q": (256, 64, 3, 256, 8, 77, 49408, 128, 2, 2),
}


def make_state_dict(arch: str, seed: int, device="cpu", logit_scale: float = math.log(100.0)) -> dict:
    E, res, vl, vw, p, ctx, vocab, tw, _, tl = ARCHS[arch]
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g) * std).to(device)

    def ln(name, width):
        sd[name + ".weight"] = 1.0 + rn(width, std=0.05)
        sd[name + ".bias"] = rn(width, std=0.05)

    def blocks(prefix, width, layers):
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        for l in range(layers):
            rb = f"{prefix}transformer.resblocks.{l}."
            sd[rb + "attn.in_proj_weight"] = rn(3 * width, width, std=width ** -0.5)
            sd[rb + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[rb + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
            sd[rb + "attn.out_proj.bias"] = rn(width, std=0.02)
            ln(rb + "ln_1", width)
            sd[rb + "mlp.c_fc.weight"] = rn(4 * width, width, std=(2 * width) ** -0.5)
            sd[rb + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
            sd[rb + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
            sd[rb + "mlp.c_proj.bias"] = rn(width, std=0.02)
            ln(rb + "ln_2", width)

    L = (res // p) ** 2 + 1
    sd["visual.conv1.weight"] = rn(vw, 3, p, p, std=(3 * p * p) ** -0.5)
    sd["visual.class_embedding"] = rn(vw, std=vw ** -0.5)
    sd["visual.positional_embedding"] = rn(L, vw, std=vw ** -0.5)
    sd["visual.proj"] = rn(vw, E, std=vw ** -0.5)
    ln("visual.ln_pre", vw)
    blocks("visual.", vw, vl)
    ln("visual.ln_post", vw)
    sd["token_embedding.weight"] = rn(vocab, tw, std=0.02)
    sd["positional_embedding"] = rn(ctx, tw, std=0.01)
    blocks("", tw, tl)
    ln("ln_final", tw)
    sd["text_projection"] = rn(tw, E, std=tw ** -0.5)
    sd["logit_scale"] = torch.tensor(float(logit_scale), device=device)
    return sd


def make_tokens(n_cls: int, vocab: int, ctx: int = 77, seed: int = 7) -> torch.Tensor:
    """Prompt-shaped token ids [n_cls, ctx]: SOT, a few body tokens, EOT (largest id), zero padding (clip.py:197-233)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n_cls, ctx, dtype=torch.long)
    for c in range(n_cls):
        n = int(torch.randint(3, 9, (1,), generator=g))
        t[c, 0] = vocab - 2
        t[c, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=g)
        t[c, 1 + n] = vocab - 1
    return t


def make_views(n_img: int, n_views: int, res: int, seed: int, device="cpu", pin: bool = False) -> torch.Tensor:
    """[n_img*n_views, 3, res, res] fp32: per-image low-frequency pattern + per-view noise of growing strength
    (view 0 is the cleanest, standing in for the un-augmented view of TPT/data/datautils.py:113-128)."""
    g = torch.Generator(device=device).manual_seed(seed)
    base = torch.randn(n_img, 1, 3, 7, 7, generator=g, device=device)
    base = torch.nn.functional.interpolate(base.view(n_img, 3, 7, 7), size=(res, res), mode="bilinear",
                                           align_corners=False).view(n_img, 1, 3, res, res) * 1.5
    strength = (0.15 + 1.2 * torch.arange(n_views, device=device) / max(1, n_views - 1)).view(1, n_views, 1, 1, 1)
    out = base + strength * torch.randn(n_img, n_views, 3, res, res, generator=g, device=device)
    out = out.view(n_img * n_views, 3, res, res).contiguous()
    if pin and out.device.type == "cpu":
        out = out.pin_memory()
    return out
