"""Drives the UNMODIFIED reference (baseline/_ref/TPT, a verbatim copy made by baseline/make_ref.py) through its own
per-image loop -- tune_cls_rl.test_time_adapt_eval (TPT/tune_cls_rl.py:183-256): reset -> test_time_tuning
(tpt_cls_rl.py:47-79) -> adapted 1-view inference -> accuracy -- on synthetic config-2 inputs, on the CPU (all host
threads, fp32: torch.cuda.amp.autocast / GradScaler are no-ops without CUDA) or on one GPU (fp16 autocast +
GradScaler(1000) + nn.MultiheadAttention, exactly as the reference runs there).  BASELINE tooling: only bench.py's
baseline legs import this; rlcf_b200/ never does, and none of this repo's kernels is on this path.

Substitutions (none touches the arithmetic of the hot path):
  * `ftfy` is not installed -> a module whose fix_text is the identity (exact for ASCII class names);
  * the hard-coded DOWNLOAD_ROOT existence check at import (clip_reward.py:12-18, custom_clip.py:24-30);
  * checkpoints: no OpenAI archives offline -> `clip.load` returns `build_model(state_dict)` (the reference's own
    constructor, clip/model.py:399-439) on seeded synthetic weights of the named architecture;
  * the dataset: a list of (64 views, label) pairs in the format of the reference's DataLoader
    (list of [1,3,224,224] tensors + target [1], datautils.py:113-128);
  * CPU only: Tensor.cuda()/Module.cuda() are identity (test_time_adapt_eval moves every batch with .cuda(args.gpu)).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "TPT")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "tune_cls_rl.py"))


def import_reference():
    if not available():
        raise RuntimeError("baseline/_ref/TPT is missing: run `python baseline/make_ref.py` in the build container")
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    real_exists = os.path.exists
    os.path.exists = lambda p: True if p == "/YOUR/PATH" else real_exists(p)
    sys.path.insert(0, REF)
    try:
        import clip.custom_clip as custom_clip
        import clip.model as clip_model
        import clip_reward
        import params
        import tpt_cls_rl
        import tune_cls_rl
    finally:
        os.path.exists = real_exists
    return types.SimpleNamespace(custom_clip=custom_clip, clip_model=clip_model, clip_reward=clip_reward, params=params,
                                 tpt_cls_rl=tpt_cls_rl, tune_cls_rl=tune_cls_rl)


def reference_args(mods, wl: dict, out_dir: str):
    """The reference's own argparse (TPT/params.py) on the command line of scripts/rlcf-tune.sh with --tune_norm 1."""
    argv = ["tune_cls_rl.py", "SYNTHETIC", "--test_sets", "A", "-a", wl["policy"], "--batch_size", str(wl["n_views"]),
            "--selection_p", str(wl["selection_p"]), "--gpu", "0", "--tpt", "--ctx_init", "a_photo_of_a",
            "--tta_steps", str(wl["tta_steps"]), "--lr", str(wl["lr"]), "--weight_decay", "5e-4",
            "--output", out_dir, "--reward_arch", wl["reward"], "--reward_amplify", "0", "--reward_process", "1",
            "--process_batch", "0", "--momentum_update", "0", "--sample_k", str(wl["sample_k"]), "--tune_norm", "1",
            "--print-freq", "1000000"]
    old = sys.argv
    sys.argv = argv
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return mods.params.get_args()
    finally:
        sys.argv = old


class ReferenceRun:
    """Model, optimizer, reward model and synthetic loader wired as TPT/tune_cls_rl.py:main_worker does (lines 66-87,
    142-143)."""

    def __init__(self, device: str, wl: dict, make_state_dict, out_dir: str = "/tmp/rlcf_ref_out"):
        self.mods = m = import_reference()
        self.device = torch.device(device)
        self.cpu = self.device.type == "cpu"
        self.wl = wl
        sds = {wl["policy"]: make_state_dict(wl["policy"], 0), wl["reward"] + "#reward": make_state_dict(wl["reward"], 1)}

        def fake_load(key):
            def load(arch, device="cpu", jit=False, download_root=None):
                sd = sds[key]
                model = m.clip_model.build_model({k: v.clone() for k, v in sd.items()}).to(device).float()
                return model, sd["text_projection"].shape[1], None
            return load

        m.custom_clip.load = fake_load(wl["policy"])
        m.clip_reward.clip.load = fake_load(wl["reward"] + "#reward")
        self.args = args = reference_args(m, wl, out_dir)
        classnames = [f"class {i}" for i in range(wl["n_classes"])]
        self._patches = []
        if self.cpu:
            self._patch(torch.Tensor, "cuda", lambda t, *a, **k: t)
            self._patch(torch.nn.Module, "cuda", lambda mod, *a, **k: mod)
        dev = self.device
        import copy
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints a banner per model (bench.py must
            model = m.custom_clip.CLIPCLS_TTA(               # print ONE JSON line on stdout)
                dev, classnames, arch=args.arch, prompt_prefix=args.ctx_init, only_visual=True,
                momentum_update=args.momentum_update, update_freq=args.update_freq, update_w=args.update_w,
                momentum=args.tta_momentum, only_norm=args.tune_norm)
            self.model = model.cuda(args.gpu)
            self.optimizer = torch.optim.AdamW(self.model.parameters(), args.lr, weight_decay=args.weight_decay)
            self.optim_state = copy.deepcopy(self.optimizer.state_dict())
            self.reward_model = m.clip_reward.get_reward_model(dev, args)
            self.reward_model.set_class_features(tokenized_classes=self.model.tokenized_prompts)
        self.scaler = torch.cuda.amp.GradScaler(init_scale=1000)
        del sds

    def _patch(self, obj, name, fn):
        self._patches.append((obj, name, getattr(obj, name)))
        setattr(obj, name, fn)

    def close(self):
        for obj, name, old in self._patches:
            setattr(obj, name, old)
        self._patches = []

    def loader(self, views: torch.Tensor, n_images: int):
        """views [n*V,3,H,W] (host) -> the reference DataLoader's batches, cycling over the n distinct images."""
        V = self.wl["n_views"]
        n = views.shape[0] // V
        batches = []
        for i in range(n_images):
            img = views[(i % n) * V:(i % n + 1) * V]
            batches.append(([img[k:k + 1].clone() for k in range(V)], torch.tensor([i % self.wl["n_classes"]])))
        return batches

    def run(self, views: torch.Tensor, n_images: int) -> float:
        """Seconds for n_images through the reference's test_time_adapt_eval."""
        m, a = self.mods, self.args
        batches = self.loader(views, n_images)
        if not self.cpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            m.tune_cls_rl.test_time_adapt_eval(batches, self.model, self.optimizer, self.optim_state, self.scaler, a,
                                               device=self.device, reward_model=self.reward_model)
        if not self.cpu:
            torch.cuda.synchronize()
        return time.perf_counter() - t0
