"""CPU oracle for the RLCF test-time-adaptation hot path.  TEST INFRASTRUCTURE ONLY.

This is a plain PyTorch fp32 restatement of the reference algorithm (mzhaoshuai/RLCF, TPT/), written from the
reference's behaviour, with every function citing the file:line it follows.  It is the checker for the CUDA
path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
The product (rlcf_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF, run in the build container from /root/reference by oracle/make_golden.py on
identical deterministic weights and inputs; the resulting vectors are committed under tests/golden/ and
tests/test_oracle_golden.py holds the oracle to them (<= 1e-5 abs on logits/rewards, identical indices).

State-dict key layout is the one of TPT/clip/model.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# deterministic synthetic CLIP weights (no OpenAI checkpoints offline; SURVEY.md 8(c) "Weights")
# --------------------------------------------------------------------------------------------------------------
ARCHS = {
    # name: (embed_dim, resolution, vision_layers, vision_width, patch, ctx_len, vocab, text_width, text_heads, text_layers)
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": (768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": (768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # small towers for fast CPU tests / golden vectors (same code paths: head_dim 64, odd token counts)
    "tiny-A": (128, 64, 2, 128, 16, 77, 512, 128, 2, 2),     # 17 image tokens
    "tiny-B": (256, 64, 3, 256, 8, 77, 512, 128, 2, 2),      # 65 image tokens (reward model in tests)
    "tiny-C": (256, 96, 3, 256, 16, 77, 512, 128, 2, 2),     # 37 image tokens at 96 px: a reward model whose
                                                             # resolution differs from the views' (bicubic resize)
    # same towers with CLIP's full vocabulary, for cases tokenised by the real BPE tokenizer (prompt tuning)
    "tiny-P": (128, 64, 2, 128, 16, 77, 49408, 128, 2, 2),
    "tiny-Q": (256, 64, 3, 256, 8, 77, 49408, 128, 2, 2),
}


def make_clip_state_dict(arch: str, seed: int, logit_scale: float = math.log(100.0)) -> dict:
    """Random CLIP weights with the scales of CLIP.initialize_parameters (TPT/clip/model.py:299-326) and the
    VisionTransformer constructor (model.py:212-221), drawn from one seeded CPU generator in a fixed key order."""
    E, res, vl, vw, p, ctx, vocab, tw, th, tl = ARCHS[arch]
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    L = (res // p) ** 2 + 1
    sd["visual.conv1.weight"] = rn(vw, 3, p, p, std=(3 * p * p) ** -0.5)
    sd["visual.class_embedding"] = rn(vw, std=vw ** -0.5)
    sd["visual.positional_embedding"] = rn(L, vw, std=vw ** -0.5)
    sd["visual.proj"] = rn(vw, E, std=vw ** -0.5)

    def ln(prefix, width):
        # non-trivial affine parameters so that d(gamma), d(beta) and the per-image slices are exercised
        sd[prefix + ".weight"] = 1.0 + rn(width, std=0.05)
        sd[prefix + ".bias"] = rn(width, std=0.05)

    def blocks(prefix, width, layers):
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for l in range(layers):
            rb = f"{prefix}transformer.resblocks.{l}."
            sd[rb + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
            sd[rb + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[rb + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
            sd[rb + "attn.out_proj.bias"] = rn(width, std=0.02)
            ln(rb + "ln_1", width)
            sd[rb + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
            sd[rb + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
            sd[rb + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
            sd[rb + "mlp.c_proj.bias"] = rn(width, std=0.02)
            ln(rb + "ln_2", width)

    ln("visual.ln_pre", vw)
    blocks("visual.", vw, vl)
    ln("visual.ln_post", vw)
    sd["token_embedding.weight"] = rn(vocab, tw, std=0.02)
    sd["positional_embedding"] = rn(ctx, tw, std=0.01)
    blocks("", tw, tl)
    ln("ln_final", tw)
    sd["text_projection"] = rn(tw, E, std=tw ** -0.5)
    sd["logit_scale"] = torch.tensor(float(logit_scale))
    return sd


def make_views(n_img: int, n_views: int, res: int, seed: int) -> torch.Tensor:
    """Synthetic augmented views [n_img*n_views, 3, res, res]: a per-image low-frequency base pattern plus
    per-view noise of varying strength, so that view entropies and class margins are well separated
    (SURVEY.md section 7 'Discrete decisions amplify noise')."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(n_img, n_views, 3, res, res)
    for i in range(n_img):
        base = F.interpolate(torch.randn(1, 3, 7, 7, generator=g), size=(res, res), mode="bilinear",
                             align_corners=False)[0] * 1.5
        for v in range(n_views):
            strength = 0.15 + 1.2 * (v / max(1, n_views - 1))
            out[i, v] = base + strength * torch.randn(3, res, res, generator=g)
    return out.view(n_img * n_views, 3, res, res)


def make_tokens(n_cls: int, vocab: int, ctx: int = 77, seed: int = 7) -> torch.Tensor:
    """Synthetic tokenised prompts [n_cls, ctx]: SOT, 3..8 body tokens, EOT (= largest id, vocab-1, so that
    argmax finds it as in TPT/clip/model.py:354), zero padding -- the shape clip.tokenize produces (clip.py:197-233)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n_cls, ctx, dtype=torch.long)
    for c in range(n_cls):
        n = int(torch.randint(3, 9, (1,), generator=g))
        t[c, 0] = vocab - 2
        t[c, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=g)
        t[c, 1 + n] = vocab - 1
    return t


# --------------------------------------------------------------------------------------------------------------
# towers
# --------------------------------------------------------------------------------------------------------------
def _block(x, sd, rb, heads, mask):
    """ResidualAttentionBlock.forward (TPT/clip/model.py:189-192) on x [N, L, d] (batch-first restatement of the
    reference's seq-first nn.MultiheadAttention call, model.py:185-187)."""
    N, L, d = x.shape
    hd = d // heads
    h = F.layer_norm(x, (d,), sd[rb + "ln_1.weight"], sd[rb + "ln_1.bias"], 1e-5)          # model.py:157-163
    qkv = h @ sd[rb + "attn.in_proj_weight"].t() + sd[rb + "attn.in_proj_bias"]
    q, k, v = qkv.view(N, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(N, L, d)
    x = x + a @ sd[rb + "attn.out_proj.weight"].t() + sd[rb + "attn.out_proj.bias"]
    h = F.layer_norm(x, (d,), sd[rb + "ln_2.weight"], sd[rb + "ln_2.bias"], 1e-5)
    u = h @ sd[rb + "mlp.c_fc.weight"].t() + sd[rb + "mlp.c_fc.bias"]
    u = u * torch.sigmoid(1.702 * u)                                                        # QuickGELU, model.py:166-168
    return x + u @ sd[rb + "mlp.c_proj.weight"].t() + sd[rb + "mlp.c_proj.bias"]


def _n_layers(sd, prefix):
    return len([k for k in sd if k.startswith(prefix + "transformer.resblocks.") and k.endswith("attn.in_proj_weight")])


def encode_image(sd: dict, images: torch.Tensor) -> torch.Tensor:
    """VisionTransformer.forward (TPT/clip/model.py:223-240): un-normalised image features [N, E]."""
    w = sd["visual.conv1.weight"]
    d, p = w.shape[0], w.shape[-1]
    x = F.conv2d(images, w, stride=p)                                                       # model.py:224
    x = x.reshape(x.shape[0], d, -1).permute(0, 2, 1)
    x = torch.cat([sd["visual.class_embedding"].expand(x.shape[0], 1, d), x], dim=1)        # model.py:227
    x = x + sd["visual.positional_embedding"]
    x = F.layer_norm(x, (d,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], 1e-5)
    for l in range(_n_layers(sd, "visual.")):
        x = _block(x, sd, f"visual.transformer.resblocks.{l}.", d // 64, None)
    x = F.layer_norm(x[:, 0, :], (d,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], 1e-5)
    return x @ sd["visual.proj"]                                                            # model.py:237-238


def encode_text(sd: dict, tokens: torch.Tensor) -> torch.Tensor:
    """CLIP.encode_text (TPT/clip/model.py:342-356): un-normalised text features [N, E]."""
    return text_from_embeddings(sd, sd["token_embedding.weight"][tokens], tokens)


def text_from_embeddings(sd: dict, prompts: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
    """TextEncoder.forward (TPT/clip/custom_clip.py:62-73) == the tail of CLIP.encode_text (model.py:345-356)."""
    x = prompts + sd["positional_embedding"]
    L, d = x.shape[1], x.shape[2]
    mask = torch.full((L, L), float("-inf")).triu_(1)                                       # model.py:328-334
    for l in range(_n_layers(sd, "")):
        x = _block(x, sd, f"transformer.resblocks.{l}.", d // 64, mask)
    x = F.layer_norm(x, (d,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    x = x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ sd["text_projection"]          # model.py:354
    return x


def class_features(sd: dict, tokens: torch.Tensor) -> torch.Tensor:
    """CLIPCLS_TTA.get_class_features (TPT/clip/custom_clip.py:404-408) / CLIPRewards.extract_text_features
    (TPT/clip_reward.py:139-150): L2-normalised text features."""
    with torch.no_grad():
        f = encode_text(sd, tokens)
        return f / f.norm(dim=-1, keepdim=True)


def policy_logits(sd: dict, class_feat: torch.Tensor, images: torch.Tensor) -> torch.Tensor:
    """CLIPCLS_TTA.forward (TPT/clip/custom_clip.py:423-432)."""
    f = encode_image(sd, images)
    f = f / f.norm(dim=-1, keepdim=True)
    return sd["logit_scale"].exp() * f @ class_feat.t()


# --------------------------------------------------------------------------------------------------------------
# TTA loop
# --------------------------------------------------------------------------------------------------------------
def select_confident_samples(logits, top):
    """TPT/tpt_cls_rl.py:32-35."""
    batch_entropy = -(logits.softmax(1) * logits.log_softmax(1)).sum(1)
    idx = torch.argsort(batch_entropy, descending=False)[:int(batch_entropy.size()[0] * top)]
    return logits[idx], idx, batch_entropy


def avg_entropy(outputs):
    """TPT/tpt_cls_rl.py:38-44."""
    logits = outputs - outputs.logsumexp(dim=-1, keepdim=True)
    avg_logits = logits.logsumexp(dim=0) - np.log(logits.shape[0])
    avg_logits = torch.clamp(avg_logits, min=torch.finfo(avg_logits.dtype).min)
    return -(avg_logits * torch.exp(avg_logits)).sum(dim=-1)


def reward_image_features(sd_reward: dict, images: torch.Tensor) -> torch.Tensor:
    """CLIPRewards.extract_image_features (TPT/clip_reward.py:130-137), including the bicubic resize to the reward
    model's own input resolution (133-134; e.g. ViT-L/14@336px scoring 224-pixel views)."""
    with torch.no_grad():
        conv = sd_reward["visual.conv1.weight"]
        res = conv.shape[-1] * int(round(math.sqrt(sd_reward["visual.positional_embedding"].shape[0] - 1)))
        if images.shape[-1] != res:
            images = F.interpolate(images, size=res, mode="bicubic", align_corners=True)
        f = encode_image(sd_reward, images).float()
        return f / f.norm(dim=1, keepdim=True)


def clip_score(reward_cls, reward_img, class_index, sample_k, weight=2.5):
    """CLIPRewards.CLIPScore with pairwise=False (TPT/clip_reward.py:111-128)."""
    text_features = reward_cls[class_index]
    image_features = torch.repeat_interleave(reward_img, sample_k, dim=0)
    similarity = weight * torch.sum(text_features * image_features, dim=-1)
    return torch.maximum(similarity, torch.zeros_like(similarity)).squeeze()


CONFIDENCES = {"ViT-L/14@336px": 10, "ViT-L/14": 5, "RN50x64": 3, "ViT-B/16": 1}      # TPT/clip_reward.py:22-27


def ensemble_weights(confidences) -> list:
    """CLIPRewardsMultiple.__init__ (TPT/clip_reward.py:207): normalised confidences rounded to 2 decimals."""
    return [round(x / sum(confidences), 2) for x in confidences]


def clip_score_multi(reward_cls_list, reward_img_list, class_index, sample_k, weights, weighted=True, weight=2.5):
    """CLIPRewardsMultiple.CLIPScore with pairwise=False (TPT/clip_reward.py:226-250)."""
    scores = torch.stack([clip_score(c, f, class_index, sample_k, weight) for c, f in zip(reward_cls_list, reward_img_list)],
                         dim=0)
    if weighted:
        w = torch.tensor(weights, dtype=scores.dtype).unsqueeze(1)
        return torch.sum(w * scores, dim=0)
    return torch.mean(scores, dim=0)


def rewards_post_process(clip_score_, reward_process=True, amplify=False):
    """CLIPRewards.rewards_post_process (TPT/clip_reward.py:152-165)."""
    if clip_score_.shape[-1] > 1 and reward_process:
        mean = torch.mean(clip_score_, dim=-1, keepdim=True)
        std = torch.std(clip_score_, dim=-1, keepdim=True) + 1e-5 if amplify else 1.0
        clip_score_ = (clip_score_ - mean) / std
    return clip_score_.flatten()


@dataclass
class OracleConfig:
    n_views: int = 64
    selection_p: float = 0.1
    tta_steps: int = 1
    sample_k: int = 3
    lr: float = 5e-3
    weight_decay: float = 5e-4
    reward_process: bool = True
    process_batch: bool = False
    reward_amplify: bool = False
    loss: str = "rlcf"        # "rlcf" | "tpt"
    reward_weights: tuple = ()    # ensemble of reward models (sd_reward / reward_cls are then lists): per-model weights
    weighted_scores: bool = True  # CLIPRewardsMultiple(weighted_scores=...): weighted sum, else plain mean
    min_entropy_w: float = 0.0    # --min_entropy_reg 1: loss += min_entropy_w * avg_entropy(output) (tpt_cls_rl.py:73-74)


def ln_param_names(sd: dict) -> list:
    """CLIPCLS_TTA.parameters with only_norm=True (TPT/clip/custom_clip.py:477-485): visual parameters whose name
    contains 'ln' (there is no BatchNorm in a ViT), in named_parameters order."""
    order = ["visual.ln_pre.weight", "visual.ln_pre.bias"]
    for l in range(_n_layers(sd, "visual.")):
        rb = f"visual.transformer.resblocks.{l}."
        order += [rb + "ln_1.weight", rb + "ln_1.bias", rb + "ln_2.weight", rb + "ln_2.bias"]
    order += ["visual.ln_post.weight", "visual.ln_post.bias"]
    return order


def adapt_one_image(sd_policy: dict, class_feat: torch.Tensor, views: torch.Tensor, cfg: OracleConfig,
                    sd_reward: dict | None = None, reward_cls: torch.Tensor | None = None, amp: bool = False,
                    scaler=None, tune: str = "ln") -> dict:
    """One iteration of the per-image loop of TPT/tune_cls_rl.py:192-222 in LayerNorm-tuning mode
    (--tune_norm 1): reset -> test_time_tuning (TPT/tpt_cls_rl.py:47-79) -> adapted prediction on views[0].
    fp32 on CPU: torch.cuda.amp.autocast and GradScaler are no-ops without CUDA, so the scaler lines
    (tpt_cls_rl.py:77-79) reduce to loss.backward(); optimizer.step().  amp=True (+ a GradScaler) reproduces the
    reference's GPU execution mode -- fp16 autocast around the forward (tpt_cls_rl.py:52) and scaled backward -- and is
    used only by bench.py's optional PyTorch-on-GPU baseline leg."""
    import contextlib
    autocast = (lambda: torch.autocast(device_type=views.device.type, dtype=torch.float16)) if amp else contextlib.nullcontext
    # tune="ln": --tune_norm 1 (LayerNorm parameters); tune="full": every visual parameter (custom_clip.py:477-479)
    names = ln_param_names(sd_policy) if tune == "ln" else [k for k in sd_policy if k.startswith("visual.")]
    sd = {k: v.clone() for k, v in sd_policy.items()}                    # model.reset(), tune_cls_rl.py:210
    params = [sd[n].requires_grad_(True) for n in names]
    opt = torch.optim.AdamW(params, cfg.lr, weight_decay=cfg.weight_decay)   # fresh state, tune_cls_rl.py:80,213
    out = {"losses": [], "grads": []}
    selected_idx = None
    for _ in range(cfg.tta_steps):
      with autocast():
        if selected_idx is not None:
            output = policy_logits(sd, class_feat, views[selected_idx])                      # tpt_cls_rl.py:55
        else:
            logits_all = policy_logits(sd, class_feat, views)                                # tpt_cls_rl.py:57
            output, selected_idx, ent = select_confident_samples(logits_all, cfg.selection_p)
            out["logits_all"], out["entropy"], out["selected_idx"] = logits_all.detach(), ent.detach(), selected_idx
            if cfg.loss == "rlcf":
                multi = isinstance(sd_reward, (list, tuple))
                reward_img = ([reward_image_features(r, views[selected_idx]) for r in sd_reward] if multi
                              else reward_image_features(sd_reward, views[selected_idx]))   # tpt_cls_rl.py:59
                out["reward_img"] = reward_img
        bs = output.shape[0]
        if cfg.loss == "rlcf":
            _, index = torch.topk(output, cfg.sample_k, dim=-1)                              # tpt_cls_rl.py:63
            flat = index.flatten()
            score = (clip_score_multi(reward_cls, reward_img, flat, cfg.sample_k, cfg.reward_weights,
                                      cfg.weighted_scores) if isinstance(reward_img, list)
                     else clip_score(reward_cls, reward_img, flat, cfg.sample_k))           # tpt_cls_rl.py:66
            rewards = rewards_post_process(score if cfg.process_batch else score.reshape(bs, -1),
                                           cfg.reward_process, cfg.reward_amplify)           # tpt_cls_rl.py:67
            rep = torch.repeat_interleave(output, cfg.sample_k, dim=0)
            all_loss = F.cross_entropy(rep, flat, reduction="none")                          # tpt_cls_rl.py:70
            loss = torch.mean(rewards * all_loss)                                            # tpt_cls_rl.py:71
            if cfg.min_entropy_w:
                loss = loss + cfg.min_entropy_w * avg_entropy(output)                        # tpt_cls_rl.py:73-74
            out.setdefault("topk_idx", []).append(index)
            out.setdefault("scores", []).append(score.reshape(bs, -1))
            out.setdefault("rewards", []).append(rewards.reshape(bs, -1))
        else:
            loss = avg_entropy(output)                                                       # tpt_cls.py:49-78
      opt.zero_grad()
      if scaler is not None:                                                                 # tpt_cls_rl.py:77-79
          scaler.scale(loss).backward()
          scaler.step(opt)
          scaler.update()
      else:
          loss.backward()
          out["grads"].append(torch.cat([p.grad.flatten() for p in params]).clone())
          opt.step()
      out["losses"].append(float(loss.detach()))
    with torch.no_grad(), autocast():
        out["logits_final"] = policy_logits(sd, class_feat, views[:1])                       # tune_cls_rl.py:218-222
    out["params"] = torch.cat([p.detach().flatten() for p in params])
    out["param_names"] = names
    out["param_dict"] = {n: p.detach() for n, p in zip(names, params)}
    out["grad_dicts"] = [dict(zip(names, [t.view_as(p) for t, p in zip(torch.split(g, [p.numel() for p in params]),
                                                                      params)])) for g in out["grads"]]
    return out


def prompt_source_map(tokens: torch.Tensor, n_ctx: int, position: str = "end", split_idx=None,
                      learned_cls: bool = False) -> torch.Tensor:
    """Where every position of every class prompt comes from in PromptLearner.forward (TPT/clip/custom_clip.py:198-289):
    src[c, t] >= 0 = position in class c's own tokenised prompt (frozen embedding), src[c, t] < 0 = learnable vector
    -1 - src (context vectors 0..n_ctx-1; with learned class tokens, custom_clip.py:209-221, vector n_ctx + c).
    The tokenised prompt is [SOS, n_ctx placeholder words, class-name tokens, '.', EOS, padding]."""
    C, L = tokens.shape
    src = torch.arange(L).repeat(C, 1)
    half = split_idx if split_idx is not None else n_ctx // 2
    for c in range(C):
        n = 1 if learned_cls else int(tokens[c].argmax()) - n_ctx - 2
        name = [1 + n_ctx + k for k in range(n)]
        ctx = [-1 - v for v in range(n_ctx)]
        if position == "end":
            order = ctx + ([-1 - (n_ctx + c)] if learned_cls else name)
        elif position == "middle":
            order = ctx[:half] + name + ctx[half:]
        elif position == "front":
            order = name + ctx
        else:
            raise ValueError(position)
        src[c, 1:1 + len(order)] = torch.tensor(order)
    return src


def prompt_text_features(sd: dict, tokens: torch.Tensor, ctx: torch.Tensor, src_map: torch.Tensor | None = None,
                         cls: torch.Tensor | None = None) -> torch.Tensor:
    """PromptLearner.forward (TPT/clip/custom_clip.py:198-289; src_map=None: class token at the end, 229-246) followed
    by ClipTestTimeTuning.get_text_features (315-323): L2-normalised text features [C, E], differentiable in ctx (and
    in the learned class vectors `cls` [C, d])."""
    emb = sd["token_embedding.weight"][tokens]                         # frozen prefix (SOS) and suffix (class, EOS)
    n_ctx = ctx.shape[0]
    if src_map is None:
        prompts = torch.cat([emb[:, :1, :], ctx.unsqueeze(0).expand(tokens.shape[0], -1, -1), emb[:, 1 + n_ctx:, :]], dim=1)
    else:
        vecs = ctx if cls is None else torch.cat([ctx, cls.reshape(tokens.shape[0], -1)], dim=0)
        frozen = torch.gather(emb, 1, src_map.clamp_min(0).unsqueeze(-1).expand(-1, -1, emb.shape[-1]))
        prompts = torch.where((src_map < 0).unsqueeze(-1), vecs[(-1 - src_map).clamp_min(0)], frozen)
    t = text_from_embeddings(sd, prompts, tokens)
    return t / t.norm(dim=-1, keepdim=True)


def adapt_one_image_prompt(sd_policy: dict, tokens: torch.Tensor, ctx_init: torch.Tensor, views: torch.Tensor,
                           cfg: OracleConfig, sd_reward: dict | None = None,
                           reward_cls: torch.Tensor | None = None, src_map: torch.Tensor | None = None,
                           cls_init: torch.Tensor | None = None) -> dict:
    """One iteration of the per-image loop of TPT/tpt_cls_rl.py:219-279 (prompt tuning): reset the context vectors,
    test_time_tuning (47-79) with ClipTestTimeTuning.inference (custom_clip.py:325-335: image tower under no_grad,
    text tower differentiable), AdamW on ctx only (tpt_cls_rl.py:103-120), adapted prediction on views[0]."""
    ctx = ctx_init.clone().requires_grad_(True)                                              # prompt_learner.reset()
    cls = None if cls_init is None else cls_init.clone().requires_grad_(True)                # learned class tokens
    trainable = [ctx] if cls is None else [ctx, cls]                                         # prompt_learner.parameters()
    opt = torch.optim.AdamW(trainable, cfg.lr, weight_decay=cfg.weight_decay)
    scale = sd_policy["logit_scale"].exp()

    def model(images):
        with torch.no_grad():
            f = encode_image(sd_policy, images)
        f = f / f.norm(dim=-1, keepdim=True)
        return scale * f @ prompt_text_features(sd_policy, tokens, ctx, src_map, cls).t()

    out = {"losses": [], "grads": []}
    selected_idx = None
    for _ in range(cfg.tta_steps):
        if selected_idx is not None:
            output = model(views[selected_idx])
        else:
            logits_all = model(views)
            output, selected_idx, ent = select_confident_samples(logits_all, cfg.selection_p)
            out["logits_all"], out["entropy"], out["selected_idx"] = logits_all.detach(), ent.detach(), selected_idx
            if cfg.loss == "rlcf":
                reward_img = reward_image_features(sd_reward, views[selected_idx])
        bs = output.shape[0]
        if cfg.loss == "rlcf":
            _, index = torch.topk(output, cfg.sample_k, dim=-1)
            flat = index.flatten()
            score = clip_score(reward_cls, reward_img, flat, cfg.sample_k)
            rewards = rewards_post_process(score if cfg.process_batch else score.reshape(bs, -1),
                                           cfg.reward_process, cfg.reward_amplify)
            all_loss = F.cross_entropy(torch.repeat_interleave(output, cfg.sample_k, dim=0), flat, reduction="none")
            loss = torch.mean(rewards * all_loss)
            out.setdefault("topk_idx", []).append(index)
            out.setdefault("rewards", []).append(rewards.reshape(bs, -1))
        else:
            loss = avg_entropy(output)
        opt.zero_grad()
        loss.backward()
        out["grads"].append(torch.cat([p.grad.flatten() for p in trainable]).clone())
        opt.step()
        out["losses"].append(float(loss.detach()))
    with torch.no_grad():
        out["logits_final"] = model(views[:1])
    out["params"] = torch.cat([p.detach().flatten() for p in trainable]).clone()
    return out


def flat_ln_params(sd: dict) -> torch.Tensor:
    return torch.cat([sd[n].detach().flatten() for n in ln_param_names(sd)])


def accuracy(output, target, topk=(1,)):
    """TPT/utils/tools.py:84-98."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


# --------------------------------------------------------------------------------------------------------------
# retrieval TTA (retrieval/clip_ret_policy.py, retrieval/custom_models.py, retrieval/clip_reward.py)
# The retrieval CLIP (retrieval/lavis/models/clip_models/model.py:262-376,538-569) is the same ViT / text transformer
# under the same state-dict keys as TPT/clip/model.py, so the towers above are reused.
# --------------------------------------------------------------------------------------------------------------
def openai_load_rounding(sd: dict) -> dict:
    """load_openai_model -> build_model_from_openai_state_dict (lavis/models/clip_models/model.py:763-791,869-871):
    the model is converted to fp16 BEFORE load_state_dict, so Conv/Linear/MultiheadAttention weights and biases,
    `visual.proj` and `text_projection` take fp16-representable values; CLIPRet_TTA / CLIPRewards then call .float()
    (custom_models.py:41, clip_reward.py:124).  A no-op for OpenAI archives (already fp16), visible for synthetic
    fp32 weights."""
    out = {}
    for k, v in sd.items():
        rounded = (k.endswith(("conv1.weight", "in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias",
                               "c_fc.weight", "c_fc.bias", "c_proj.weight", "c_proj.bias"))
                   or k in ("visual.proj", "text_projection"))
        out[k] = v.half().float() if rounded else v.clone()
    return out


def retrieval_features(sd: dict, images=None, tokens=None) -> torch.Tensor:
    """CLIPRet_TTA.get_image_features / get_text_features (retrieval/custom_models.py:78-90) and the reward model's
    extract_*_features (retrieval/clip_reward.py:170-189): L2-normalised features."""
    f = encode_image(sd, images) if images is not None else encode_text(sd, tokens)
    return F.normalize(f, dim=-1)


def retrieval_clip_score(reward_text, reward_image, sample_k, text_index=None, images_index=None, weight=2.5):
    """CLIPRewards.CLIPScore with pairwise=False (retrieval/clip_reward.py:143-168)."""
    t = reward_text[text_index] if text_index is not None else torch.repeat_interleave(reward_text, sample_k, dim=0)
    i = reward_image[images_index] if images_index is not None else torch.repeat_interleave(reward_image, sample_k, dim=0)
    similarity = weight * torch.sum(t * i, dim=-1)
    return torch.maximum(similarity, torch.zeros_like(similarity)).squeeze()


@dataclass
class RetrievalConfig:
    tta_steps: int = 8            # scripts/tta_coco_ret.sh:33
    sample_k: int = 20            # 20 for image->text, 12 for text->image (tta_coco_ret.sh:19-20)
    lr: float = 1e-6
    weight_decay: float = 5e-4
    eps: float = 1e-6             # clip_ret_policy.py:235
    reward_process: bool = True
    process_batch: bool = False
    reward_amplify: bool = False
    momentum_update: bool = False
    update_freq: int = 256
    update_w: float = 1.0
    momentum: float = 0.9999


def retrieval_trainable(sd: dict, task: str) -> list:
    """CLIPRet_TTA.parameters (retrieval/custom_models.py:144-152): image->text tunes every `visual.*` parameter,
    text->image every other parameter -- including token_embedding and logit_scale.  Sorted by name so that a
    concatenation of them does not depend on the state dict's insertion order."""
    vis = task == "image2text"
    return sorted(k for k in sd if k.startswith("visual.") == vis)


_AUTOCAST_KEYS = ("conv1.weight", "in_proj_weight", "out_proj.weight", "c_fc.weight", "c_proj.weight")


def autocast_weights(sd: dict) -> dict:
    """What torch.cuda.amp.autocast does to the parameters on the reference's GPU path (clip_ret_policy.py:87,120): every
    Conv / Linear / MultiheadAttention weight is cast to fp16 for the matmul at each forward, the fp32 master stays the
    optimizer's.  The cast is differentiable (identity gradient), so the masters still receive gradient."""
    return {k: (v.half().float() if k.endswith(_AUTOCAST_KEYS) else v) for k, v in sd.items()}


def retrieval_tune_query(sd_init: dict, cfg: RetrievalConfig, task: str, query: torch.Tensor, gallery: torch.Tensor,
                         reward_query: torch.Tensor, reward_gallery: torch.Tensor, fp16_weights: bool = False) -> dict:
    """tune_image / tune_text (retrieval/clip_ret_policy.py:76-137) followed by the evaluation forward of
    test_time_tune (161-168 / 178-184), for ONE query on the weights `sd_init`.

    task "image2text": query = image [1,3,H,W]; gallery = policy text features [Nt,E]; reward_query = reward-model
    feature of the image [1,Er]; reward_gallery = reward text features [Nt,Er].
    task "text2image": query = token ids [1,77]; gallery = policy image features [Ni,E]; reward_query = reward text
    feature [1,Er]; reward_gallery = reward image features [Ni,Er]."""
    i2t = task == "image2text"
    sd = {k: v.clone() for k, v in sd_init.items()}                      # reset_initial(), custom_models.py:122-124
    names = retrieval_trainable(sd, task)
    params = [sd[n].requires_grad_(True) for n in names]
    opt = torch.optim.AdamW(params, lr=cfg.lr, eps=cfg.eps, weight_decay=cfg.weight_decay)   # clip_ret_policy.py:235
    K = cfg.sample_k

    def model():                                                         # CLIPRet_TTA.forward, custom_models.py:66-76
        w = autocast_weights(sd) if fp16_weights else sd                 # fp16_weights: the reference's GPU numerics
        f = retrieval_features(w, images=query) if i2t else retrieval_features(w, tokens=query)
        return sd["logit_scale"].exp() * f @ gallery.t()

    out = {"losses": [], "topk_idx": [], "scores": [], "rewards": [], "grads": []}
    for _ in range(cfg.tta_steps):
        opt.zero_grad()
        logits = model()
        _, index = torch.topk(logits, K, dim=-1)                         # clip_ret_policy.py:90 / 123
        flat = index.flatten()
        if i2t:
            score = retrieval_clip_score(reward_gallery, reward_query, K, text_index=flat)
        else:
            score = retrieval_clip_score(reward_query, reward_gallery, K, images_index=flat)
        rewards = rewards_post_process(score if cfg.process_batch else score.reshape(logits.shape[0], -1),
                                       cfg.reward_process, cfg.reward_amplify)
        rep = torch.repeat_interleave(logits, K, dim=0)
        loss = torch.mean(rewards * F.cross_entropy(rep, flat, reduction="none"))   # clip_ret_policy.py:97-98
        loss.backward()
        out["grads"].append({n: (torch.zeros_like(p) if p.grad is None else p.grad.clone()) for n, p in zip(names, params)})
        opt.step()
        out["losses"].append(float(loss.detach()))
        out["topk_idx"].append(index.detach()[0])
        out["scores"].append(score.detach().reshape(-1))
        out["rewards"].append(rewards.detach().reshape(-1))
    with torch.no_grad():
        out["score_row"] = model()[0]                                    # clip_ret_policy.py:166-167 / 183-184
    out["state"] = {k: v.detach() for k, v in sd.items()}
    out["param_names"] = names
    return out


class RetrievalMomentum:
    """CLIPRet_TTA's cross-query state (retrieval/custom_models.py:55-60,114-142): `initial` is what every query
    starts from; with momentum_update an EMA of the adapted weights replaces it every `update_freq` queries."""

    def __init__(self, sd: dict, cfg: RetrievalConfig):
        self.cfg = cfg
        self.clip = {k: v.clone() for k, v in sd.items()}
        self.initial = {k: v.clone() for k, v in sd.items()}
        self.ema = {k: v.clone() for k, v in sd.items()} if cfg.momentum_update else None
        self.counter = 0

    def update(self, adapted: dict):
        """momentum_update_model (custom_models.py:126-142), called after each query with the adapted weights."""
        c = self.cfg
        if not c.momentum_update:
            return
        self.counter += 1
        for k, v in adapted.items():
            self.ema[k] = c.momentum * self.ema[k] + (1.0 - c.momentum) * v
        if self.counter >= c.update_freq:
            self.counter = 0
            for k in adapted:
                self.initial[k] = (1 - c.update_w) * self.clip[k] + c.update_w * self.ema[k]


def retrieval_report_metrics(scores_i2t: np.ndarray, scores_t2i: np.ndarray, txt2img, img2txt) -> dict:
    """RetrievalTask._report_metrics (retrieval/lavis/tasks/retrieval.py:52-107) without the log-file write:
    recall@1/5/10 both ways.  Ranks follow np.argsort(score)[::-1] exactly (ties broken as numpy does)."""
    ranks = np.zeros(scores_i2t.shape[0])
    for index, score in enumerate(scores_i2t):
        inds = np.argsort(score)[::-1]
        ranks[index] = min(int(np.where(inds == i)[0][0]) for i in img2txt[index])
    tr = [100.0 * float(np.sum(ranks < k)) / len(ranks) for k in (1, 5, 10)]
    ranks = np.zeros(scores_t2i.shape[0])
    for index, score in enumerate(scores_t2i):
        inds = np.argsort(score)[::-1]
        ranks[index] = np.where(inds == txt2img[index])[0][0]
    ir = [100.0 * float(np.sum(ranks < k)) / len(ranks) for k in (1, 5, 10)]
    tr_mean, ir_mean = sum(tr) / 3, sum(ir) / 3
    return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": tr_mean, "img_r1": ir[0], "img_r5": ir[1],
            "img_r10": ir[2], "img_r_mean": ir_mean, "r_mean": (tr_mean + ir_mean) / 2, "agg_metrics": sum(tr) / 3}
