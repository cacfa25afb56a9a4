"""CLIP model objects with the reference's attribute / state_dict layout (TPT/clip/model.py) whose forward passes
run on the sm_100a kernels of librlcf_b200.so.

The modules below hold fp32 parameters under exactly the reference's names
(`visual.transformer.resblocks.N.attn.in_proj_weight`, `visual.ln_post.weight`, `text_projection`, ...), so OpenAI
checkpoints and the reference's `state_dict()/load_state_dict()/named_parameters()` bookkeeping
(custom_clip.py:394-399,456-485) work unchanged.  There is no autograd graph: `forward` is inference through the
CUDA kernels; adaptation goes through `rlcf_b200.tpt_cls_rl.test_time_tuning`, which updates the LayerNorm
parameters in place.  Only ViT towers are supported (the configured models are all ViTs; ModifiedResNet is out of
scope).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .. import engine as E
from .._lib import RlcfError


class LayerNorm(nn.LayerNorm):
    """Parameter holder with nn.LayerNorm's names (model.py:157-163); computed by rlcf_layernorm_fwd."""


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) (model.py:166-168); fused into the c_fc GEMM epilogue."""


class _AttnParams(nn.Module):
    """nn.MultiheadAttention's parameter names (packed in_proj, out_proj Linear) without its forward."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, n_head
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        if d_model != 64 * n_head:
            raise RlcfError(f"head_dim must be 64 (width {d_model}, heads {n_head})")
        self.attn = _AttnParams(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])


class _KernelBacked:
    """Mixin: caches kernel-ready fp16 weights and keeps the LayerNorm parameters as views of one flat fp32 buffer
    (the layout the kernels index), re-packing whenever the parameters were re-allocated (.cuda(), load_state_dict)."""

    _prefix = ""

    def _invalidate(self):
        self._tower = None
        self._tower_key = None

    def _frozen_key(self):
        # frozen GEMM weights are re-prepared only when their storage or version changes
        return tuple((p.data_ptr(), p._version) for n, p in self.named_parameters() if "ln" not in n)

    def _build_tower(self, need_grad):  # pragma: no cover - implemented by subclasses
        raise NotImplementedError

    def tower(self, need_grad: bool = False) -> E.TowerWeights:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RlcfError("rlcf_b200 models run on CUDA only (no CPU fallback); move the model to a GPU")
        key = (self._frozen_key(), need_grad)
        t = getattr(self, "_tower", None)
        if t is None or self._tower_key[0] != key[0] or (need_grad and not self._tower_key[1]):
            t = self._build_tower(need_grad)
            self._tower, self._tower_key = t, key
        self._bind_ln(t)
        return t

    def _bind_ln(self, t: E.TowerWeights):
        """Make every LayerNorm parameter a view into t.ln_flat (so in-place updates of either are shared)."""
        named = dict(self.named_parameters())
        flat, d = t.ln_flat, t.d
        for key, off in t.ln_names(""):
            p = named[key]
            if p.data_ptr() != flat.data_ptr() + 4 * off:
                flat[off:off + d].copy_(p.data.to(flat.dtype))
                p.data = flat[off:off + d]

    def ln_flat(self) -> torch.Tensor:
        return self.tower().ln_flat


class VisionTransformer(_KernelBacked, nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._invalidate()

    def _build_tower(self, need_grad):
        return E.prepare_visual({k: v for k, v in self.state_dict().items()}, prefix="", need_grad=need_grad)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N,3,H,W] -> un-normalised image features [N, output_dim] (model.py:223-240)."""
        t = self.tower()
        x = x.float().contiguous()
        n = x.shape[0]
        run = getattr(self, "_runner", None)
        if run is None or run.w is not t or run.max_seq < n:
            run = self._runner = E.TowerRunner(t, n)
        xs = run.forward(n, t.ln_flat, images=x)
        feat = torch.empty(n, t.E, dtype=torch.float32, device=x.device)
        inv = torch.empty(n, dtype=torch.float32, device=x.device)
        run.head(xs, n, t.ln_flat, feat=feat, inv_norm=inv)
        return feat / inv[:, None]     # head_fwd returns f/|f| and 1/|f|

    @torch.no_grad()
    def logits(self, x: torch.Tensor, class_feat: torch.Tensor, logit_scale: float) -> torch.Tensor:
        """[N,3,H,W] -> cosine logits [N, C] against L2-normalised class features: tower forward + one rlcf_head_fwd
        launch (ln_post, projection, normalisation, logit_scale * f . class_feat) -- custom_clip.py:423-432."""
        t = self.tower()
        x = x.float().contiguous()
        n = x.shape[0]
        run = getattr(self, "_runner", None)
        if run is None or run.w is not t or run.max_seq < n:
            run = self._runner = E.TowerRunner(t, n)
        xs = run.forward(n, t.ln_flat, images=x)
        out = torch.empty(n, class_feat.shape[0], dtype=torch.float32, device=x.device)
        run.head(xs, n, t.ln_flat, class_feat=class_feat.float().contiguous(), logit_scale=logit_scale, logits=out)
        return out


class _TextTower(_KernelBacked, nn.Module):
    """Owns nothing: a view over CLIP's text-side parameters so they can be prepared as one tower."""

    def __init__(self, clip_model):
        super().__init__()
        object.__setattr__(self, "_clip", clip_model)
        self._invalidate()

    def named_parameters(self, *a, **k):
        c = self._clip
        yield "token_embedding.weight", c.token_embedding.weight
        yield "positional_embedding", c.positional_embedding
        for n, p in c.transformer.named_parameters():
            yield "transformer." + n, p
        for n, p in c.ln_final.named_parameters():
            yield "ln_final." + n, p
        yield "text_projection", c.text_projection

    def parameters(self, recurse=True):
        return (p for _, p in self.named_parameters())

    def _build_tower(self, need_grad):
        return E.prepare_text({n: p.data for n, p in self.named_parameters()}, need_grad=need_grad)


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers, vision_width: int, vision_patch_size: int,
                 context_length: int, vocab_size: int, transformer_width: int, transformer_heads: int,
                 transformer_layers: int):
        super().__init__()
        if isinstance(vision_layers, (tuple, list)):
            raise RlcfError("ModifiedResNet towers are out of scope; rlcf_b200 supports the ViT CLIP models")
        self.context_length = context_length
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()
        object.__setattr__(self, "_text", _TextTower(self))

    def initialize_parameters(self):
        """Same scales as model.py:299-326."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        for tr in (self.transformer, self.visual.transformer):
            proj_std = (tr.width ** -0.5) * ((2 * tr.layers) ** -0.5)
            for block in tr.resblocks:
                nn.init.normal_(block.attn.in_proj_weight, std=tr.width ** -0.5)
                nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(block.mlp.c_fc.weight, std=(2 * tr.width) ** -0.5)
                nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    @torch.no_grad()
    def encode_text(self, text):
        """Token ids [N, ctx] -> un-normalised text features [N, embed_dim] (model.py:342-356)."""
        t = self._text.tower()
        text = text.to(self.positional_embedding.device)
        return E.text_features(t, text, normalized=False)

    def forward(self, image, text):
        image_features = self.encode_image(image)
        text_features = self.encode_text(text)
        image_features = image_features / image_features.norm(dim=1, keepdim=True)
        text_features = text_features / text_features.norm(dim=1, keepdim=True)
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * image_features @ text_features.t()
        return logits_per_image, logits_per_image.t()


def build_model(state_dict: dict) -> CLIP:
    """Infers the architecture from a checkpoint's shapes and loads it (model.py:399-439)."""
    if "visual.proj" not in state_dict:
        raise RlcfError("ModifiedResNet checkpoints are out of scope; rlcf_b200 supports the ViT CLIP models")
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    image_resolution = vision_patch_size * grid_size
    embed_dim = state_dict["text_projection"].shape[1]
    context_length = state_dict["positional_embedding"].shape[0]
    vocab_size = state_dict["token_embedding.weight"].shape[0]
    transformer_width = state_dict["ln_final.weight"].shape[0]
    transformer_heads = transformer_width // 64
    transformer_layers = len(set(k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")))
    model = CLIP(embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length,
                 vocab_size, transformer_width, transformer_heads, transformer_layers)
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    model.load_state_dict(sd)
    return model.eval()
