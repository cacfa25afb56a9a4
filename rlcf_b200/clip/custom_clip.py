"""Policy wrappers with the reference's surface (TPT/clip/custom_clip.py).

CLIPCLS_TTA (custom_clip.py:364-497): CLIP classification with test-time adaptation of the image encoder.  Class
text features are computed once with the policy's text tower; `forward(image) -> logits[N, C]` runs the image
tower on the CUDA kernels; `reset()`, `reset_classnames_and_state()`, `parameters()`, `train()/eval()`,
`momentum_update_model()` keep the reference semantics.  The LayerNorm-only mode (`only_norm=True`, the
`--tune_norm 1` recipe) is the one `test_time_tuning` can adapt on the GPU; full-encoder tuning (wgrad GEMMs + an
86 M-parameter AdamW) is listed as the next step in DESIGN.md and raises NotImplementedError at tuning time.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .. import engine as E
from .._lib import RlcfError
from .clip import load, tokenize

DOWNLOAD_ROOT = None  # the reference hard-codes a checkpoint directory (custom_clip.py:19-30); here it is optional


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
        image_features = self.clip_model.encode_image(image)
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)
        logit_scale = self.clip_model.logit_scale.exp()
        return logit_scale * image_features @ self.class_features.t()

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

    # ------------------------------------------------------------------ bridge to the batched CUDA engine
    def engine(self, cfg: E.RlcfConfig, n_img: int, reward_model=None) -> E.RlcfEngine:
        """RlcfEngine for this policy (+ reward model) and configuration, cached."""
        if not self.only_norm:
            raise NotImplementedError("GPU adaptation currently covers LayerNorm tuning (--tune_norm 1); full "
                                      "image-encoder tuning is the next scope row (DESIGN.md, SURVEY.md 8(f2))")
        vis = self.clip_model.visual
        pol = vis.tower(need_grad=True)
        rew = rcls = None
        if reward_model is not None:
            rew = reward_model.clip_model.visual.tower()
            rcls = reward_model.class_features
            if rcls is None:
                raise RlcfError("reward_model.set_class_features(...) must be called before adaptation")
        key = (id(pol), id(rew), n_img, tuple(sorted(vars(cfg).items())), self.class_features.data_ptr(),
               None if rcls is None else rcls.data_ptr())
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()
            eng = E.RlcfEngine(pol, self.class_features, float(self.clip_model.logit_scale.exp()), cfg, n_img,
                               reward=rew, reward_class_feat=rcls)
            self._engines[key] = eng
        return eng
