"""Policy wrappers with the reference's surface (TPT/clip/custom_clip.py).

CLIPCLS_TTA (custom_clip.py:364-497): CLIP classification with test-time adaptation of the image encoder.  Class
text features are computed once with the policy's text tower; `forward(image) -> logits[N, C]` runs the image
tower on the CUDA kernels; `reset()`, `reset_classnames_and_state()`, `parameters()`, `train()/eval()`,
`momentum_update_model()` keep the reference semantics.  `test_time_tuning` adapts either mode on the GPU:
LayerNorm-only (`only_norm=True`, the `--tune_norm 1` recipe; engine.RlcfEngine) or the whole image encoder
(`only_norm=False`, scripts/rlcf-tune.sh; full_tune.FullTuneEngine: wgrad GEMMs + an 86 M-parameter AdamW per image).

Engines are cached per model.  A cache entry keeps strong references to every object its key identifies (towers,
class-feature tensors), so an `id()` / `data_ptr()` in the key can never be recycled by a new object while the entry
is alive.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .. import engine as E
from .._lib import RlcfError
from .clip import load, tokenize

DOWNLOAD_ROOT = None  # the reference hard-codes a checkpoint directory (custom_clip.py:19-30); here it is optional


def _reward_inputs(reward_model):
    """(tower(s), class features, weights) of a CLIPRewards / CLIPRewardsMultiple for the engines."""
    from ..clip_reward import engine_inputs
    rew, rcls, weights = engine_inputs(reward_model)
    if rcls is None:
        raise RlcfError("reward_model.set_class_features(...) must be called before adaptation")
    return rew, rcls, weights


def _ids(x):
    return None if x is None else (tuple(id(t) for t in x) if isinstance(x, (list, tuple)) else id(x))


def _ptrs(x):
    return None if x is None else (tuple(t.data_ptr() for t in x) if isinstance(x, (list, tuple)) else x.data_ptr())


class TextEncoder(nn.Module):
    """Text tower on ready-made prompt embeddings (custom_clip.py:53-73): [C, 77, d] -> un-normalised features [C, E]."""

    def __init__(self, clip_model):
        super().__init__()
        object.__setattr__(self, "_clip", clip_model)
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype

    @torch.no_grad()
    def forward(self, prompts, tokenized_prompts):
        t = self._clip._text.tower()
        n, L = tokenized_prompts.shape
        run = E.TowerRunner(t, n)
        x = run.forward(n, t.ln_flat, prompt=prompts.contiguous())
        eot = tokenized_prompts.argmax(dim=-1).to(device=x.device, dtype=torch.int32)
        rows = (torch.arange(n, device=x.device, dtype=torch.int32) * L + eot).contiguous()
        feat = torch.empty(n, t.E, dtype=torch.float32, device=x.device)
        inv = torch.empty(n, dtype=torch.float32, device=x.device)
        run.head(x, n, t.ln_flat, row_idx=rows, feat=feat, inv_norm=inv)
        return feat / inv[:, None]


class PromptLearner(nn.Module):
    """CoOp-style learnable context (custom_clip.py:76-289).  Prompts are assembled from the SOS embedding, the n_ctx
    learnable context vectors, the class-name tokens and the rest (".", EOS, padding) with the class token at the
    `end` ([SOS | ctx | class ...], every RLCF / TPT script), in the `middle` ([SOS | ctx[:h] | class | ctx[h:] | ...],
    also selected by a "[CLS]" word inside `ctx_init`) or at the `front` ([SOS | class | ctx | ...]); `learned_cls`
    replaces the class-name token by one learnable vector per class (custom_clip.py:209-221, position `end` only).
    `batch_size` (one context per view, custom_clip.py:120-122) is not supported.  `tokenized_prompts` / `ctx_tokens`
    bypass the BPE tokenizer (synthetic vocabularies); `cls_init` fixes the initial learned class vectors."""

    def __init__(self, clip_model, classnames, batch_size=None, n_ctx=16, ctx_init=None, ctx_position="end",
                 learned_cls=False, tokenized_prompts=None, ctx_tokens=None, cls_init=None):
        super().__init__()
        if batch_size is not None:
            raise NotImplementedError("PromptLearner(batch_size=...) (one context per view) is not supported")
        if ctx_position not in ("end", "middle", "front"):
            raise ValueError(f"ctx_position {ctx_position!r}")
        self.learned_cls, self.batch_size = learned_cls, batch_size
        self.dtype = clip_model.dtype
        self.device = clip_model.visual.conv1.weight.device
        self.ctx_dim = clip_model.ln_final.weight.shape[0]
        object.__setattr__(self, "_clip", clip_model)
        split_idx = None
        if ctx_init or ctx_tokens is not None:
            if ctx_init:
                ctx_init = ctx_init.replace("_", " ")
                if "[CLS]" in ctx_init:                                   # custom_clip.py:93-99
                    split_idx = ctx_init.split(" ").index("[CLS]")
                    ctx_init = ctx_init.replace("[CLS] ", "")
                    ctx_position = "middle"
            if ctx_tokens is None:
                n_ctx = len(ctx_init.split(" "))
                ctx_tokens = tokenize(ctx_init)[0, 1:1 + n_ctx]
            n_ctx = len(ctx_tokens)
            with torch.no_grad():
                ctx_vectors = clip_model.token_embedding(ctx_tokens.to(self.device)).type(self.dtype)
            prompt_prefix = ctx_init if ctx_init else " ".join(["X"] * n_ctx)
        else:
            ctx_vectors = torch.empty(n_ctx, self.ctx_dim, dtype=self.dtype, device=self.device)
            nn.init.normal_(ctx_vectors, std=0.02)
            prompt_prefix = " ".join(["X"] * n_ctx)
        if learned_cls and ctx_position != "end":
            raise AssertionError("learned_cls needs ctx_position='end' (custom_clip.py:227-228)")
        self.prompt_prefix, self.split_idx = prompt_prefix, split_idx
        self.ctx_init_state = ctx_vectors.detach().clone()
        self.ctx = nn.Parameter(ctx_vectors.detach().clone())
        self.ctx_init, self.class_token_position, self.n_ctx = ctx_init, ctx_position, n_ctx
        self._set_classes(classnames, tokenized_prompts, cls_init)

    def _set_classes(self, classnames, tokenized_prompts, cls_init=None):
        self.n_cls = len(classnames)
        self.classnames = [name.replace("_", " ") for name in classnames]
        if self.learned_cls:
            # one learnable vector per class in place of the class-name token "X" (custom_clip.py:130-140)
            if cls_init is None:
                cls_init = torch.empty(self.n_cls, 1, self.ctx_dim, dtype=self.dtype, device=self.device)
                nn.init.normal_(cls_init, std=0.02)
            cls_init = cls_init.to(device=self.device, dtype=self.dtype).reshape(self.n_cls, 1, self.ctx_dim)
            self.cls_init_state = cls_init.detach().clone()
            if hasattr(self, "cls") and self.cls.shape == cls_init.shape:
                with torch.no_grad():
                    self.cls.copy_(cls_init)
            else:
                self.cls = nn.Parameter(cls_init.detach().clone())
        if tokenized_prompts is None:
            names = ["X"] * self.n_cls if self.learned_cls else self.classnames
            prompts = [self.prompt_prefix + " " + name + "." for name in names]
            tokenized_prompts = torch.cat([tokenize(p) for p in prompts])
        self.tokenized_prompts = tokenized_prompts.to(self.device)
        # tokens: SOS, n_ctx placeholder words, the class name, ".", EOS -> name_len = eot - n_ctx - 2
        self.name_lens = [1 if self.learned_cls else int(t.argmax()) - 1 - self.n_ctx - 1 for t in self.tokenized_prompts]
        with torch.no_grad():
            embedding = self._clip.token_embedding(self.tokenized_prompts).type(self.dtype)
        self.token_prefix = embedding[:, :1, :]                  # SOS
        skip = 1 if self.learned_cls else 0
        self.token_suffix = embedding[:, 1 + self.n_ctx + skip:, :]     # class tokens, EOS, padding
        self._layout = None

    # ---- where every position of every class prompt comes from (custom_clip.py:229-289)
    def source_map(self):
        """(src_map [C, L], ctx_pos [C, n_ctx], cls_pos [C] | None) as int32 CPU tensors; see engine.PromptLayout."""
        C, L, n_ctx = self.n_cls, self.tokenized_prompts.shape[1], self.n_ctx
        src = torch.arange(L, dtype=torch.int32).repeat(C, 1)
        ctx_pos = torch.empty(C, n_ctx, dtype=torch.int32)
        cls_pos = torch.full((C,), 1 + n_ctx, dtype=torch.int32) if self.learned_cls else None
        where = self.class_token_position
        half = self.split_idx if self.split_idx is not None else n_ctx // 2
        for c in range(C):
            n = self.name_lens[c]
            if where == "end":
                order = [("ctx", v) for v in range(n_ctx)]
                if self.learned_cls:
                    order.append(("cls", c))
                else:
                    order += [("tok", 1 + n_ctx + k) for k in range(n)]
            elif where == "middle":
                order = ([("ctx", v) for v in range(half)] + [("tok", 1 + n_ctx + k) for k in range(n)]
                         + [("ctx", v) for v in range(half, n_ctx)])
            else:
                order = [("tok", 1 + n_ctx + k) for k in range(n)] + [("ctx", v) for v in range(n_ctx)]
            for t, (kind, v) in enumerate(order, start=1):
                if kind == "ctx":
                    src[c, t] = -1 - v
                    ctx_pos[c, v] = t
                elif kind == "cls":
                    src[c, t] = -1 - (n_ctx + v)
                else:
                    src[c, t] = v
        return src, ctx_pos, cls_pos

    def layout(self):
        """engine.PromptLayout on the device, or None for the default arrangement (served by the specialised kernels)."""
        if self.class_token_position == "end" and not self.learned_cls:
            return None
        if self._layout is None:
            src, ctx_pos, cls_pos = self.source_map()
            dev = self.device
            self._layout = E.PromptLayout(src.to(dev).contiguous(), ctx_pos.to(dev).contiguous(),
                                          None if cls_pos is None else cls_pos.to(dev).contiguous(), self.n_ctx)
        return self._layout

    def learnable_flat(self):
        """The trainable vectors as the engine's flat block: context vectors, then the learned class vectors."""
        parts = [self.ctx.detach().reshape(-1)]
        if self.learned_cls:
            parts.append(self.cls.detach().reshape(-1))
        return torch.cat(parts).float()

    @torch.no_grad()
    def load_flat(self, flat):
        n = self.ctx.numel()
        self.ctx.copy_(flat[:n].view_as(self.ctx))
        if self.learned_cls:
            self.cls.copy_(flat[n:n + self.cls.numel()].view_as(self.cls))

    @torch.no_grad()
    def reset(self):
        self.ctx.copy_(self.ctx_init_state)
        if self.learned_cls:
            self.cls.copy_(self.cls_init_state)

    def reset_classnames(self, classnames, arch, tokenized_prompts=None, cls_init=None):
        self._set_classes(classnames, tokenized_prompts, cls_init)

    def forward(self, init=None):
        """Prompt embeddings [C, L, d] (custom_clip.py:198-289), assembled from the source map."""
        ctx = self.ctx if init is None else init
        emb = self._clip.token_embedding(self.tokenized_prompts).type(self.dtype)
        if self.class_token_position == "end" and not self.learned_cls:
            return torch.cat([self.token_prefix, ctx.unsqueeze(0).expand(self.n_cls, -1, -1), self.token_suffix], dim=-2)
        src, _, _ = self.source_map()
        src = src.to(emb.device).long()
        vecs = ctx if not self.learned_cls else torch.cat([ctx, self.cls.reshape(self.n_cls, -1)], dim=0)
        out = torch.gather(emb, 1, src.clamp_min(0).unsqueeze(-1).expand(-1, -1, emb.shape[-1]))
        learn = src < 0
        out = torch.where(learn.unsqueeze(-1), vecs[(-1 - src).clamp_min(0)], out)
        return out


class ClipTestTimeTuning(nn.Module):
    """Prompt-tuning policy (custom_clip.py:292-344): frozen CLIP + PromptLearner; forward(image) -> logits [N, C]."""

    def __init__(self, device, classnames, batch_size, criterion="cosine", arch="ViT-L/14", n_ctx=16, ctx_init=None,
                 ctx_position="end", learned_cls=False, tokenized_prompts=None, ctx_tokens=None, cls_init=None):
        super().__init__()
        clip_model, _, _ = load(arch, device=device, download_root=DOWNLOAD_ROOT)
        object.__setattr__(self, "_clip", clip_model)
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = clip_model.logit_scale.data
        self.prompt_learner = PromptLearner(clip_model, classnames, batch_size, n_ctx, ctx_init, ctx_position,
                                            learned_cls, tokenized_prompts, ctx_tokens, cls_init)
        self.criterion = criterion
        self._engines = {}

    @property
    def dtype(self):
        return self.image_encoder.conv1.weight.dtype

    def reset(self):
        self.prompt_learner.reset()

    def reset_classnames(self, classnames, arch, tokenized_prompts=None):
        self.prompt_learner.reset_classnames(classnames, arch, tokenized_prompts)
        self._engines = {}

    @torch.no_grad()
    def get_text_features(self):
        t = self.text_encoder(self.prompt_learner(), self.prompt_learner.tokenized_prompts)
        return t / t.norm(dim=-1, keepdim=True)

    @torch.no_grad()
    def inference(self, image):
        image_features = self.image_encoder(image.type(self.dtype))
        text_features = self.get_text_features()
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)
        return self.logit_scale.exp() * image_features @ text_features.t()

    def forward(self, input):
        if isinstance(input, tuple) or input.dim() == 2:
            raise NotImplementedError("contrastive / directional prompt tuning are not part of the RLCF path")
        return self.inference(input)

    def engine(self, cfg: E.RlcfConfig, n_img: int, reward_model=None) -> E.PromptEngine:
        """PromptEngine (batched GPU adaptation of the context vectors) for this model, cached."""
        vis = self._clip.visual.tower()
        txt = self._clip._text.tower(need_grad=True)
        rew = rcls = None
        if reward_model is not None:
            rew, rcls, cfg.reward_weights = _reward_inputs(reward_model)
        pl = self.prompt_learner
        layout = pl.layout()
        key = (id(vis), id(txt), _ids(rew), n_img, tuple(sorted(vars(cfg).items())), pl.tokenized_prompts.data_ptr(),
               _ptrs(rcls), id(layout))
        hit = self._engines.get(key)
        if hit is None:
            self._engines.clear()
            vecs = pl.learnable_flat().view(-1, pl.ctx_dim)
            eng = E.PromptEngine(vis, txt, pl.tokenized_prompts, vecs, float(self.logit_scale.exp()), cfg,
                                 n_img, reward=rew, reward_class_feat=rcls, layout=layout)
            hit = self._engines[key] = (eng, (vis, txt, rew, rcls, pl.tokenized_prompts, layout))   # refs pin the keyed ids
        return hit[0]


def get_coop(clip_arch, test_set, device, n_ctx, ctx_init, learned_cls=False, classnames=None, tokenized_prompts=None):
    """Factory with the reference's signature (custom_clip.py:347-361).  The reference looks the class names up in its
    dataset tables (TPT/data/, not shipped here); pass `classnames` (and optionally `tokenized_prompts`) instead."""
    if classnames is None:
        raise RlcfError(f"class-name table for test set {test_set!r} is reference data that is not shipped: "
                        "pass classnames=[...]")
    return ClipTestTimeTuning(device, classnames, None, arch=clip_arch, n_ctx=n_ctx, ctx_init=ctx_init,
                              learned_cls=learned_cls, tokenized_prompts=tokenized_prompts)


class CLIPCLS_TTA(nn.Module):
    def __init__(self, device, classnames, arch="ViT-L/14", prompt_prefix=None, only_visual=True,
                 momentum_update=False, update_freq=256, update_w=1.0, momentum=0.9999, only_norm=False,
                 tokenized_prompts=None):
        """`tokenized_prompts` (optional, [C, 77] int64) bypasses the BPE tokenizer -- used with synthetic weights."""
        super().__init__()
        self.clip_model, _, _ = load(arch, device=device, download_root=DOWNLOAD_ROOT)
        self.device = device
        self.prompt_prefix = prompt_prefix
        self.classnames = [name.replace("_", " ") for name in classnames]
        self.n_cls = len(classnames)
        if tokenized_prompts is None:
            prompts = [self.prompt_prefix + " " + name + "." for name in self.classnames]
            tokenized_prompts = tokenize(prompts)
        self.tokenized_prompts = tokenized_prompts.to(self.device)
        self.class_features = self.get_class_features(self.tokenized_prompts)
        self.only_visual, self.only_norm = only_visual, only_norm
        self.freeze_parameters()
        self.momentum_update, self.update_freq, self.update_w, self.momentum = (momentum_update, update_freq,
                                                                                update_w, momentum)
        self.update_counter = 0
        with torch.no_grad():   # in-memory snapshots (custom_clip.py:394-399)
            self.clip_state_dict = copy.deepcopy(self.clip_model.visual.state_dict())
            self.initial_state_dict = copy.deepcopy(self.clip_model.visual.state_dict())
            if self.momentum_update:
                self.momentum_state_dict = copy.deepcopy(self.clip_model.visual.state_dict())
        self._engines = {}

    @torch.no_grad()
    def get_class_features(self, tokenized_prompts):
        class_features = self.clip_model.encode_text(tokenized_prompts)
        return class_features / class_features.norm(dim=-1, keepdim=True)

    @torch.no_grad()
    def freeze_parameters(self):
        if self.only_visual:
            for n, p in self.clip_model.named_parameters():
                if "visual" not in n:
                    p.requires_grad_(False)

    @torch.no_grad()
    def forward(self, image):
        """logits [N, C] = exp(logit_scale) * normalise(encode_image(image)) @ class_features^T  (custom_clip.py:423-432);
        ln_post, projection, normalisation and the cosine logits are one rlcf_head_fwd launch."""
        return self.clip_model.visual.logits(image, self.class_features, float(self.clip_model.logit_scale.exp()))

    @torch.no_grad()
    def reset_classnames_and_state(self, classnames, arch, tokenized_prompts=None):
        self.n_cls = len(classnames)
        self.classnames = [name.replace("_", " ") for name in classnames]
        if tokenized_prompts is None:
            prompts = [self.prompt_prefix + " " + name + "." for name in self.classnames]
            tokenized_prompts = torch.cat([tokenize(p) for p in prompts])
        self.tokenized_prompts = tokenized_prompts.to(self.device)
        if not self.only_visual:
            clip_m, _, _ = load(arch, device=self.device, download_root=DOWNLOAD_ROOT)
            class_features = clip_m.encode_text(self.tokenized_prompts)
            self.class_features = class_features / class_features.norm(dim=-1, keepdim=True)
        else:
            self.class_features = self.get_class_features(self.tokenized_prompts)
        self.clip_model.visual.load_state_dict(self.clip_state_dict)
        self.initial_state_dict = copy.deepcopy(self.clip_model.visual.state_dict())
        if self.momentum_update:
            self.momentum_state_dict = copy.deepcopy(self.clip_model.visual.state_dict())
        self._engines = {}

    @torch.no_grad()
    def reset(self):
        """Restore the pre-adaptation weights (custom_clip.py:456-458).  In LayerNorm-tuning mode only the LayerNorm
        entries can have changed, so only they are restored (160 KB for ViT-B/16 instead of the reference's 345 MB
        state-dict copy per test image)."""
        vis = self.clip_model.visual
        if not self.only_norm:
            vis.load_state_dict(self.initial_state_dict)
            return
        named = dict(vis.named_parameters())
        for k, v in self.initial_state_dict.items():
            if "ln" in k or "bn" in k:
                named[k].data.copy_(v)

    @torch.no_grad()
    def momentum_update_model(self):
        """EMA of the adapted weights folded back into the reset state every update_freq samples
        (custom_clip.py:460-475).  Makes samples order-dependent: single-process only."""
        if not self.momentum_update:
            return
        self.update_counter += 1
        state_dict = self.clip_model.visual.state_dict()
        for k, v in state_dict.items():
            self.momentum_state_dict[k] = self.momentum * self.momentum_state_dict[k] + (1.0 - self.momentum) * v
        if self.update_counter >= self.update_freq:
            self.update_counter = 0
            for k in state_dict:
                self.initial_state_dict[k] = ((1 - self.update_w) * self.clip_state_dict[k]
                                              + self.update_w * self.momentum_state_dict[k])

    def parameters(self, recurse: bool = True):
        if not self.only_norm:
            return self.clip_model.visual.parameters()
        return [p for n, p in self.clip_model.visual.named_parameters() if "ln" in n or "bn" in n]

    def train(self, mode: bool = True):
        super().train(mode)
        self.clip_model.transformer.eval()
        self.clip_model.ln_final.eval()
        return self

    def _full_engine(self, cfg, n_img, reward_model):
        from .. import full_tune as FT
        if reward_model is None or reward_model.class_features is None:
            raise RlcfError("full image-encoder tuning needs a reward model with class features set")
        vis = self.clip_model.visual
        rew, rcls, cfg.reward_weights = _reward_inputs(reward_model)
        # keyed on WHICH tensors hold the weights, not on their versions: reset() / the copy-back of adapted weights
        # bump every version without changing the reset state, and the engine checks the contents itself
        storage = tuple(p.data_ptr() for p in vis.parameters())
        key = ("full", storage, _ids(rew), n_img, tuple(sorted(vars(cfg).items())),
               self.class_features.data_ptr(), _ptrs(rcls))
        hit = self._engines.get(key)
        if hit is None:
            self._engines.clear()
            eng = FT.FullTuneEngine(vis.state_dict(), self.class_features, float(self.clip_model.logit_scale.exp()), cfg,
                                    n_img, rew, rcls, prefix="")
            hit = self._engines[key] = (eng, (vis, rew, rcls, self.class_features))
        else:
            hit[0].sync_initial_state(vis.state_dict(), prefix="")
        return hit[0]

    # ------------------------------------------------------------------ bridge to the batched CUDA engine
    def engine(self, cfg: E.RlcfConfig, n_img: int, reward_model=None) -> E.RlcfEngine:
        """Engine for this policy (+ reward model) and configuration, cached: RlcfEngine for LayerNorm tuning
        (only_norm=True), FullTuneEngine when the whole image encoder is trainable (only_norm=False)."""
        vis = self.clip_model.visual
        if not self.only_norm:
            return self._full_engine(cfg, n_img, reward_model)
        pol = vis.tower(need_grad=True)
        rew = rcls = None
        if reward_model is not None:
            rew, rcls, cfg.reward_weights = _reward_inputs(reward_model)
        key = (id(pol), _ids(rew), n_img, tuple(sorted(vars(cfg).items())), self.class_features.data_ptr(), _ptrs(rcls))
        hit = self._engines.get(key)
        if hit is None:
            self._engines.clear()
            eng = E.RlcfEngine(pol, self.class_features, float(self.clip_model.logit_scale.exp()), cfg, n_img,
                               reward=rew, reward_class_feat=rcls)
            hit = self._engines[key] = (eng, (pol, rew, rcls, self.class_features))   # refs pin the keyed ids
        return hit[0]
