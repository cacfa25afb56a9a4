"""The TTA inner loop with the reference's signatures (TPT/tpt_cls_rl.py:32-79), running on the CUDA engine.

test_time_tuning(model, inputs, optimizer, scaler, args, reward_model) adapts `model` (a CLIPCLS_TTA in
LayerNorm-tuning mode) on the views of ONE test image, or -- an extension the reference does not have -- on
`inputs` holding n_img * args.batch_size views of several independent images, each adapted from the same initial
state (pass n_img=...).  The caller's torch optimizer supplies the hyper-parameters (lr, betas, eps, weight decay);
the AdamW update itself is the fused rlcf_adamw_step kernel starting from an empty state, which is what the
reference's per-sample `optimizer.load_state_dict(optim_state)` (tune_cls_rl.py:213) amounts to.  The GradScaler
is not needed (fp32 gradients with a static scale on the fp16 dgrad operands) and is ignored.
"""
from __future__ import annotations

import torch

from . import engine as E
from . import ops
from ._lib import RlcfError


def select_confident_samples(logits, top):
    """(logits[idx], idx) of the int(N*top) lowest-entropy rows (tpt_cls_rl.py:32-35), rlcf_entropy_select kernel."""
    n, c = logits.shape
    s = int(n * top)
    lg = logits.float().contiguous()
    sel = torch.empty(1, max(s, 1), dtype=torch.int32, device=logits.device)
    if s > 0:
        ops.entropy_select(lg, 1, n, c, s, sel)
    idx = sel[0, :s].long()
    return logits[idx], idx


def avg_entropy(outputs):
    """Entropy of the view-averaged prediction (tpt_cls_rl.py:38-44), rlcf_avg_entropy_loss kernel (value only)."""
    s, c = outputs.shape
    lg = outputs.float().contiguous()
    scratch = torch.empty_like(lg)
    loss = torch.empty(1, dtype=torch.float32, device=outputs.device)
    ops.avg_entropy_loss(lg, None, 1, s, c, scratch, loss=loss)
    return loss[0]


def engine_config(args, optimizer, reward_model, loss="rlcf") -> E.RlcfConfig:
    g = optimizer.param_groups[0]
    return E.RlcfConfig(
        n_views=args.batch_size, selection_p=args.selection_p, tta_steps=args.tta_steps,
        sample_k=reward_model.sample_k if reward_model is not None else 1, lr=g["lr"],
        weight_decay=g["weight_decay"], betas=tuple(g["betas"]), eps=g["eps"],
        clipscore_weight=getattr(reward_model, "clipscore_weight", 2.5),
        reward_process=bool(getattr(reward_model, "reward_process", True)),
        process_batch=bool(getattr(reward_model, "process_batch", False)),
        reward_amplify=bool(getattr(reward_model, "amplify_rewards", False)), loss=loss,
        min_entropy_w=float(getattr(args, "min_entropy_w", 0.0)) if getattr(args, "min_entropy_reg", 0) else 0.0)


def test_time_tuning(model, inputs, optimizer, scaler, args, reward_model=None, n_img: int = 1):
    """Updates the LayerNorm parameters of `model` in place (n_img == 1) and returns the adapted slices [n_img, P]."""
    if not hasattr(model, "engine"):
        raise NotImplementedError("test_time_tuning needs a CLIPCLS_TTA or ClipTestTimeTuning model")
    if optimizer.state:
        raise RlcfError("optimizer state must be empty (call optimizer.load_state_dict(optim_state) first, as "
                        "tune_cls_rl.py:213 does): the fused AdamW restarts from step 0 for every image")
    n_views = inputs.shape[0] // n_img
    cfg = engine_config(args, optimizer, reward_model, loss="rlcf" if reward_model is not None else "tpt")
    cfg.n_views = n_views
    eng = model.engine(cfg, n_img, reward_model)
    if hasattr(model, "prompt_learner"):           # prompt tuning: the trainable slice is prompt_learner.ctx
        pl = model.prompt_learner
        eng.init_ctx.copy_(pl.learnable_flat())      # context vectors (+ learned class vectors) of the reset state
        eng.refresh_initial_text_features()
        params = eng.tune(inputs.float().contiguous())
        if n_img == 1:
            pl.load_flat(params[0])
    elif hasattr(eng, "init_rest"):                # full image-encoder tuning (engine built from the current weights)
        params = eng.tune(inputs.float().contiguous())
        if n_img == 1:
            named = dict(model.clip_model.visual.named_parameters())
            with torch.no_grad():
                for key, val in eng.export_params(0, prefix="").items():
                    named[key].copy_(val.view_as(named[key]))
    else:                                          # image-encoder (LayerNorm) tuning
        vis = model.clip_model.visual
        eng.init_params.copy_(vis.ln_flat())       # adapt from the model's current (reset) state
        params = eng.tune(inputs.float().contiguous())
        if n_img == 1:
            vis.ln_flat().copy_(params[0])         # model(image) now sees the adapted parameters
    if reward_model is not None:                   # mirror the reference's side effect (tpt_cls_rl.py:59)
        reward_model.image_features = eng.reward_feat
    return params
