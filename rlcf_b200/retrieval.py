"""Retrieval test-time adaptation (retrieval/clip_ret_policy.py:76-195, retrieval/custom_models.py:28-160), config 4.

Every query (one image for image->text, one caption for text->image) tunes ALL parameters of its encoder for
`tta_steps` AdamW steps against a fixed gallery of candidate features, is scored once with the adapted weights and
the weights are restored.  Queries are independent when --momentum_update 0 (the published recipe,
scripts/tta_coco_ret.sh:39-44), so `n_query` of them run in one launch sequence:

  * step 1 of every query starts from the same initial weights -> one batched pass over shared weights;
  * from step 2 on each query owns its weights (fp32 masters + Adam moments + fp16 GEMM copies, 1.7 GB per query for
    ViT-B/16) and every Linear is ONE grouped tcgen05 launch with one weight group per query (M = 197 rows per
    group): the pass is weight-bandwidth bound, which is what HBM3e is for;
  * the 5k-25k wide score row never leaves the device: top-K sampling, CLIPScore, rewards, loss and dlogits are one
    kernel (rlcf_retrieval_loss), d(feature) = dlogits @ gallery a two-stage deterministic reduction.

With --momentum_update 1 the initial weights of query i+1 depend on the adapted weights of queries <= i
(custom_models.py:126-142): run with n_query = 1 and call momentum_update() after each query.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import engine as E
from . import full_tune as FT
from . import ops
from ._lib import RlcfError


@dataclass
class RetrievalConfig:
    """Hyper-parameters; names follow retrieval/params.py and scripts/tta_coco_ret.sh."""
    tta_steps: int = 8
    sample_k: int = 20            # 20 for image->text, 12 for text->image
    lr: float = 1e-6
    weight_decay: float = 5e-4
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-6             # clip_ret_policy.py:235
    clipscore_weight: float = 2.5
    reward_process: bool = True
    process_batch: bool = False   # one query per loss: batch statistics == per-query statistics
    reward_amplify: bool = False
    loss_scale: float = 1024.0    # static scale on the fp16 dgrad operands (the reference's GradScaler(1000))
    momentum_update: bool = False
    update_freq: int = 256
    update_w: float = 1.0
    momentum: float = 0.9999


def _bytes_per_query(lay, steps: int) -> float:
    g, n = lay.p_gemm, lay.total
    return float(steps * (34 * g + 36 * (n - g)) + 2 * g)


def _n_chunks(C: int) -> int:
    return max(1, min(128, C // 64))


class ImageQueryEngine:
    """image->text retrieval TTA (tune_image + the evaluation forward of test_time_tune, clip_ret_policy.py:76-103,
    161-168) for `n_query` independent query images per call.

    sd_policy: CLIP state dict (only `visual.*` is read); gallery_feat: policy text features of the gallery
    [C, E], L2-normalised (CLIPRet_TTA.set_text_features); reward: the reward model's visual tower;
    reward_gallery: its text features of the gallery [C, Er] (CLIPRewards.set_many_text_features)."""

    def __init__(self, sd_policy: dict, gallery_feat: torch.Tensor, logit_scale: float, cfg: RetrievalConfig,
                 n_query: int, reward: E.TowerWeights, reward_gallery: torch.Tensor, prefix: str = "visual."):
        self.cfg, self.n_query = cfg, n_query
        base = E.prepare_visual(sd_policy, prefix=prefix, need_grad=True)
        dev = base.ln_flat.device
        self.lay = lay = FT.FullLayout(base)
        self.gallery = gallery_feat.float().contiguous()
        self.reward, self.reward_gallery = reward, reward_gallery.float().contiguous()
        self.logit_scale = float(logit_scale)
        Q, C, K, P = n_query, self.gallery.shape[0], cfg.sample_k, base.P
        if K > C:
            raise RlcfError(f"sample_k {K} exceeds the gallery size {C}")
        if self.reward_gallery.shape[0] != C:
            raise RlcfError("policy and reward galleries differ in size")
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        # initial weights (what every query starts from): fp32 masters + fp16 GEMM copies, as one "sample" of the layout
        self.clip_ln, self.clip_rest = base.ln_flat.clone(), FT.pack_rest(sd_policy, lay, prefix)
        self.init_ln, self.init_rest = self.clip_ln.clone(), self.clip_rest.clone()
        self.init_w16 = torch.empty(1, lay.p_gemm, **f16)
        self.init_w16t = torch.empty(1, lay.p_gemm, **f16)
        self.w0 = FT.tower_view(base, lay, self.init_rest, self.init_w16[0], self.init_w16t[0], self.init_ln)
        self._cast_initial()
        if cfg.momentum_update:
            self.ema_ln, self.ema_rest = self.clip_ln.clone(), self.clip_rest.clone()
            self.update_counter = 0
        # per-query state
        self.ln = torch.empty(Q, P, **f32); self.ln_m = torch.empty(Q, P, **f32); self.ln_v = torch.empty(Q, P, **f32)
        self.n_slots = E.N_SLOTS
        self.partials = torch.empty(Q, self.n_slots, P, **f32)
        self.ln_grad = torch.empty(Q, P, **f32)
        self.rest = torch.empty(Q, lay.total, **f32)
        self.rest_m = torch.empty(Q, lay.total, **f32)
        self.rest_v = torch.empty(Q, lay.total, **f32)
        self.grads = torch.zeros(Q, lay.total, **f32)
        self.w16 = torch.empty(Q, lay.p_gemm, **f16)
        self.w16t = torch.empty(Q, lay.p_gemm, **f16) if cfg.tta_steps > 1 else None
        self.gw = FT.tower_view_grouped(base, lay, self.rest, self.w16, self.w16t, self.ln)
        self.run = E.TowerRunner(base, Q)
        self.run.reserve_backward(Q)
        self.store = E.ActStore(base, Q, dev, full=True)
        self.rrun = E.TowerRunner(reward, Q)
        self.hook = FT.WgradHook(lay, base, Q, 1, dev)
        self.base = base
        self.n_chunks = _n_chunks(C)
        self.ones = torch.ones(self.n_chunks, **f32)
        self.df_partial = torch.empty(Q, self.n_chunks, base.E, **f32)
        self.feat = torch.empty(Q, base.E, **f32)
        self.inv_norm = torch.empty(Q, **f32)
        self.logits = torch.empty(Q, C, **f32)
        self.dlogits = torch.empty(Q, C, **f32)
        self.topk_idx = torch.empty(cfg.tta_steps, Q, K, **i32)
        self.scores = torch.empty(cfg.tta_steps, Q, K, **f32)
        self.rewards = torch.empty(cfg.tta_steps, Q, K, **f32)
        self.loss = torch.empty(cfg.tta_steps, Q, **f32)
        self.reward_feat = torch.empty(Q, reward.E, **f32)
        self.score_rows = torch.empty(Q, C, **f32)
        self.seq_idx = torch.arange(Q, device=dev, dtype=torch.int32)
        self._graph = None
        self._static_images = None
        self.fused_adamw = FT.FUSED_ADAMW

    # ------------------------------------------------------------------ initial weights / momentum
    def _cast_initial(self):
        FT.cast_weights(self.lay, self.init_rest[None], self.init_w16, self.init_w16t)

    def momentum_update(self, q: int = 0):
        """CLIPRet_TTA.momentum_update_model (custom_models.py:126-142) with the adapted weights of query q, then
        reset_initial: EMA of the adapted weights; every update_freq queries the initial weights become
        (1-update_w)*clip + update_w*EMA.  Plain tensor arithmetic on the flat parameter vectors (once per query)."""
        c = self.cfg
        if not c.momentum_update:
            return
        self.update_counter += 1
        self.ema_ln.mul_(c.momentum).add_(self.ln[q], alpha=1.0 - c.momentum)
        self.ema_rest.mul_(c.momentum).add_(self.rest[q], alpha=1.0 - c.momentum)
        if self.update_counter >= c.update_freq:
            self.update_counter = 0
            torch.add(self.clip_ln * (1 - c.update_w), self.ema_ln, alpha=c.update_w, out=self.init_ln)
            torch.add(self.clip_rest * (1 - c.update_w), self.ema_rest, alpha=c.update_w, out=self.init_rest)
            self._cast_initial()

    # ------------------------------------------------------------------ one TTA step
    def _step(self, step: int, images: torch.Tensor, w: E.TowerWeights):
        cfg, Q, lay, base, hk = self.cfg, self.n_query, self.lay, self.base, self.hook
        P, C, K = base.P, self.gallery.shape[0], cfg.sample_k
        xs = self.run.forward(Q, self.ln, pstride=P, seqs_per_set=1, images=images, store=self.store, w=w)
        self.run.head(xs, Q, self.ln, pstride=P, seqs_per_set=1, feat=self.feat, inv_norm=self.inv_norm, w=w)
        # logits_per_image = logit_scale * image_features @ text_features.t()      (custom_models.py:66-76)
        ops.pair_logits(self.feat, self.gallery, 0, Q, 1, C, base.E, self.logit_scale, self.logits)
        ops.retrieval_loss(self.logits, self.reward_feat, self.reward_gallery, K, self.dlogits,
                           clipscore_weight=cfg.clipscore_weight, reward_process=cfg.reward_process,
                           amplify=cfg.reward_amplify, loss_scale=cfg.loss_scale, topk_idx=self.topk_idx[step - 1],
                           scores=self.scores[step - 1], rewards=self.rewards[step - 1], loss=self.loss[step - 1])
        ops.dfeat_partial(self.dlogits, self.gallery, self.df_partial)
        self.partials.zero_()
        self.run.dres[:Q * base.L].zero_()
        off = base.ln_off("ln_post")
        lnv = self.ln.view(-1)
        # the head backward sums the gallery chunks: "K" = n_chunks unit weights against this query's partial sums
        ops.head_bwd_ex(self.ones, (0, 0, 1), xs, lnv[off:], w.proj, self.df_partial, self.n_chunks * base.E,
                        self.logit_scale, self.feat, self.inv_norm, Q, 1, base.d, base.E, self.n_chunks, self.run.dres,
                        row_stride=base.L, param_stride=P, partials=self.partials, n_slots=self.n_slots, p_total=P,
                        p_off=off, beta=lnv[off + base.d:], y_out=hk.y, df_out=hk.df, proj_stride=w.proj_stride)
        fz = None
        if self.fused_adamw:   # the wgrad GEMMs apply AdamW to the GEMM weights in their epilogue
            first = step == 1
            fz = FT.FusedAdamw(self.rest, self.rest_m, self.rest_v, self.w16, self.init_rest if first else self.rest,
                               0 if first else lay.total, first, cfg, step)
        hk.bind(self.grads, Q, self.run.patches, fused=fz)
        hk.proj()
        self.run.backward(self.store, Q, 1, self.ln, P, self.partials, self.n_slots, w=w, hook=hk)
        kw = dict(beta1=cfg.betas[0], beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
                  loss_scale=cfg.loss_scale)
        ops.adamw_step(self.ln, self.ln_m, self.ln_v, self.partials, Q, self.n_slots, P, cfg.lr, step,
                       grad_out=self.ln_grad, **kw)
        w16t = self.w16t if step < cfg.tta_steps else None
        if step == 1:   # masters come from the shared initial copy, moments start at zero: no per-query restore
            FT.adamw_weights(lay, self.rest, self.rest_m, self.rest_v, self.grads, self.w16, w16t, cfg, step,
                             self.init_rest, 0, True, fused=self.fused_adamw)
        else:
            FT.adamw_weights(lay, self.rest, self.rest_m, self.rest_v, self.grads, self.w16, w16t, cfg, step,
                             self.rest, lay.total, False, fused=self.fused_adamw)

    def tune(self, images: torch.Tensor):
        """tune_image for n_query images [Q,3,H,W] at once.  Leaves the adapted parameters in self.ln / self.rest."""
        cfg, Q, P = self.cfg, self.n_query, self.base.P
        if images.shape[0] != Q:
            raise RlcfError(f"expected {Q} query images, got {images.shape[0]}")
        # reward_model.set_image_features(images=image)                          (clip_ret_policy.py:83)
        xr = self.rrun.forward(Q, self.reward.ln_flat, images=images)
        self.rrun.head(xr, Q, self.reward.ln_flat, feat=self.reward_feat)
        ops.reset_params(self.init_ln, self.ln, self.ln_m, self.ln_v, Q, P)
        if cfg.tta_steps == 0:
            self.rest.copy_(self.init_rest.expand_as(self.rest))
            FT.cast_weights(self.lay, self.rest, self.w16, None)
        for step in range(1, cfg.tta_steps + 1):
            self._step(step, images, self.w0 if step == 1 else self.gw)
        return self.ln, self.rest

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        """Score row of every query with its adapted weights (clip_ret_policy.py:161-167)."""
        Q, P, base = self.n_query, self.base.P, self.base
        xf = self.run.forward(Q, self.ln, pstride=P, seqs_per_set=1, images=images, w=self.gw)
        self.run.head(xf, Q, self.ln, pstride=P, seqs_per_set=1, feat=self.feat, w=self.gw)
        ops.pair_logits(self.feat, self.gallery, 0, Q, 1, self.gallery.shape[0], base.E, self.logit_scale,
                        self.score_rows)
        return self.score_rows

    def adapt(self, images: torch.Tensor) -> torch.Tensor:
        self.tune(images)
        return self.predict(images)

    @property
    def logits_final(self):   # name shared with the classification engines (graph / host-pipeline helpers)
        return self.score_rows

    capture = E.RlcfEngine.capture
    adapt_graph = E.RlcfEngine.adapt_graph
    adapt_host = E.RlcfEngine.adapt_host
    host_pipeline = E.RlcfEngine.host_pipeline

    def export_params(self, q: int, prefix: str = "visual.") -> dict:
        """Adapted parameters of query q under the reference's state-dict names."""
        lay, base = self.lay, self.base
        out = {}
        for key, off, shape in lay.entries(prefix):
            n = 1
            for s_ in shape:
                n *= s_
            t = self.rest[q, off:off + n].view(shape)
            if key.endswith("conv1.weight"):
                t = t[:, :3 * base.patch * base.patch].reshape(lay.d, 3, base.patch, base.patch)
            out[key] = t
        for key, off in base.ln_names(prefix):
            out[key] = self.ln[q, off:off + base.d]
        return out

    def algorithmic_flops_per_query(self) -> float:
        """2*MACs: reward forward + tta_steps x (forward + dgrad + wgrad) + evaluation forward + score rows."""
        w, cfg, C = self.base, self.cfg, self.gallery.shape[0]
        f = E.RlcfEngine.tower_fwd_flops(w)
        wgrad = w.n_layers * 24 * w.L * w.d * w.d + 2 * (w.L - 1) * w.d * 3 * w.patch * w.patch + 2 * w.d * w.E
        step = f + E.RlcfEngine.tower_dgrad_flops(w) + wgrad + 4 * C * w.E
        return float(E.RlcfEngine.tower_fwd_flops(self.reward) + cfg.tta_steps * step + f + 2 * C * w.E)

    def bytes_per_query(self) -> float:
        """Algorithmic HBM bytes of one query (the path is weight-bandwidth bound, SURVEY.md 8(d) tail-kernel rule).
        Per step and GEMM weight: 2 B + 2 B read by the forward and dgrad GEMMs, 12 B read + 12 B written for the fp32
        master and the two Adam moments, 2 B for the refreshed fp16 copy, 2 B + 2 B to re-lay the transposed dgrad copy
        = 34 B (the gradient itself stays in TMEM: fused epilogue); every other parameter: gradient written and read
        (8 B) + AdamW (28 B).  Plus one fp16 weight read for the scoring forward."""
        return _bytes_per_query(self.lay, self.cfg.tta_steps)


# ====================================================================================================================
# text -> image
# ====================================================================================================================
class TextLayout:
    """Offsets of one caption query's trainable non-LayerNorm parameters inside a flat fp32 vector:
    [ GEMM block: per layer in_proj_w, out_proj_w, c_fc_w, c_proj_w ]   (cast to fp16 in one go)
    [ positional_embedding | text_projection | per layer biases | the caption's L token-embedding rows | logit_scale ]
    Attribute names match full_tune.FullLayout so that WgradHook / cast_weights serve both towers."""

    def __init__(self, w: E.TowerWeights):
        d, L, Ed, nl = w.d, w.L, w.E, w.n_layers
        self.d, self.L, self.E, self.nl, self.k_pad = d, L, Ed, nl, 0
        off = 0
        self.wq, self.wo, self.wf, self.wp = [], [], [], []
        for _ in range(nl):
            self.wq.append(off); off += 3 * d * d
            self.wo.append(off); off += d * d
            self.wf.append(off); off += 4 * d * d
            self.wp.append(off); off += 4 * d * d
        self.p_gemm = off
        self.pos = off; off += L * d
        self.proj = off; off += d * Ed
        self.bq, self.bo, self.bf, self.bp = [], [], [], []
        for _ in range(nl):
            self.bq.append(off); off += 3 * d
            self.bo.append(off); off += d
            self.bf.append(off); off += 4 * d
            self.bp.append(off); off += d
        self.shared = off                    # everything before this is a parameter of the model, not of the caption
        self.tok = off; off += L * d
        self.ls = off; off += 8              # logit_scale + padding: per-query vectors stay 32-byte aligned (GEMM group strides)
        self.total = off

    def entries(self):
        d, nl = self.d, self.nl
        out = [("positional_embedding", self.pos, (self.L, d)), ("text_projection", self.proj, (d, self.E))]
        for l in range(nl):
            rb = f"transformer.resblocks.{l}."
            out += [(rb + "attn.in_proj_weight", self.wq[l], (3 * d, d)), (rb + "attn.out_proj.weight", self.wo[l], (d, d)),
                    (rb + "mlp.c_fc.weight", self.wf[l], (4 * d, d)), (rb + "mlp.c_proj.weight", self.wp[l], (d, 4 * d)),
                    (rb + "attn.in_proj_bias", self.bq[l], (3 * d,)), (rb + "attn.out_proj.bias", self.bo[l], (d,)),
                    (rb + "mlp.c_fc.bias", self.bf[l], (4 * d,)), (rb + "mlp.c_proj.bias", self.bp[l], (d,))]
        return out


def _text_view_grouped(base: E.TowerWeights, lay: TextLayout, rest, w16, w16t, ln) -> E.TowerWeights:
    d, G = lay.d, rest.shape[0]
    t = E.TowerWeights(kind="text", d=d, heads=base.heads, n_layers=lay.nl, L=lay.L, E=lay.E, has_ln_pre=False)

    def m(buf, off, r, c):
        return None if buf is None else buf.as_strided((G, r, c), (buf.stride(0), c, 1), buf.storage_offset() + off)

    def b(off, n):
        return rest.as_strided((G, n), (rest.stride(0), 1), rest.storage_offset() + off)

    t.pos = rest[0, lay.pos:lay.pos + lay.L * d].view(lay.L, d)
    t.proj = rest[0, lay.proj:lay.proj + d * lay.E].view(d, lay.E)
    t.embed_stride = t.proj_stride = rest.stride(0)
    t.ln_flat = ln[0]   # P = one sample's LayerNorm slice; callers pass the [G, P] tensor and its stride explicitly
    for l in range(lay.nl):
        t.layers.append(E.LayerWeights(
            wqkv=m(w16, lay.wq[l], 3 * d, d), bqkv=b(lay.bq[l], 3 * d), wo=m(w16, lay.wo[l], d, d), bo=b(lay.bo[l], d),
            wfc=m(w16, lay.wf[l], 4 * d, d), bfc=b(lay.bf[l], 4 * d), wproj=m(w16, lay.wp[l], d, 4 * d),
            bproj=b(lay.bp[l], d), wqkv_t=m(w16t, lay.wq[l], d, 3 * d), wo_t=m(w16t, lay.wo[l], d, d),
            wfc_t=m(w16t, lay.wf[l], d, 4 * d), wproj_t=m(w16t, lay.wp[l], 4 * d, d)))
    return t


class TextQueryEngine:
    """text->image retrieval TTA (tune_text + the evaluation forward, clip_ret_policy.py:106-137,178-184) for
    `n_query` independent captions per call.  Every non-visual parameter is tuned (custom_models.py:144-152): the
    text transformer, ln_final, text_projection, positional_embedding, logit_scale and token_embedding -- of which
    only the caption's own rows can receive gradient or influence its score row, so each query carries a private
    copy of those L rows (positions holding the same token id are tied through a summed gradient).

    sd_policy: CLIP state dict (non-`visual.` keys are read); gallery_feat: policy image features [C,E];
    reward: the reward model's TEXT tower; reward_gallery: its image features of the gallery [C,Er]."""

    def __init__(self, sd_policy: dict, gallery_feat: torch.Tensor, cfg: RetrievalConfig, n_query: int,
                 reward: E.TowerWeights, reward_gallery: torch.Tensor):
        self.cfg, self.n_query = cfg, n_query
        base = E.prepare_text(sd_policy, need_grad=True)
        dev = base.ln_flat.device
        self.base = base
        self.lay = lay = TextLayout(base)
        self.gallery = gallery_feat.float().contiguous()
        self.reward, self.reward_gallery = reward, reward_gallery.float().contiguous()
        Q, C, K, P, L, d = n_query, self.gallery.shape[0], cfg.sample_k, base.P, base.L, base.d
        if K > C:
            raise RlcfError(f"sample_k {K} exceeds the gallery size {C}")
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.clip_ln = base.ln_flat.clone()
        self.clip_rest = torch.zeros(lay.total, **f32)
        for key, off, shape in lay.entries():
            src = sd_policy[key].detach().float().reshape(-1)
            self.clip_rest[off:off + src.numel()].copy_(src)
        self.clip_rest[lay.ls] = sd_policy["logit_scale"].detach().float().reshape(())
        self.clip_tab = base.tok_emb.clone()
        self.init_ln, self.init_rest, self.init_tab = self.clip_ln.clone(), self.clip_rest.clone(), self.clip_tab.clone()
        if cfg.momentum_update:
            self.ema_ln, self.ema_rest, self.ema_tab = self.clip_ln.clone(), self.clip_rest.clone(), self.clip_tab.clone()
            self.update_counter = 0
        self.ln = torch.empty(Q, P, **f32); self.ln_m = torch.empty(Q, P, **f32); self.ln_v = torch.empty(Q, P, **f32)
        self.n_slots = E.N_SLOTS
        self.partials = torch.empty(Q, self.n_slots, P, **f32)
        self.ln_grad = torch.empty(Q, P, **f32)
        self.rest = torch.empty(Q, lay.total, **f32)
        self.rest_m = torch.empty(Q, lay.total, **f32)
        self.rest_v = torch.empty(Q, lay.total, **f32)
        self.grads = torch.zeros(Q, lay.total, **f32)
        self.w16 = torch.empty(Q, lay.p_gemm, **f16)
        self.w16t = torch.empty(Q, lay.p_gemm, **f16)
        self.gw = _text_view_grouped(base, lay, self.rest, self.w16, self.w16t, self.ln)
        self.rows_in = E.EmbedRows(self.rest.view(-1)[lay.tok:], self.rest.view(-1)[lay.pos:], lay.total)
        self.run = E.TowerRunner(base, Q)
        self.run.reserve_backward(Q)
        self.store = E.ActStore(base, Q, dev, full=True)
        self.rrun = E.TowerRunner(reward, Q)
        self.hook = FT.WgradHook(lay, base, Q, 1, dev)
        self.n_chunks = _n_chunks(C)
        self.ones = torch.ones(self.n_chunks, **f32)
        self.df_partial = torch.empty(Q, self.n_chunks, base.E, **f32)
        self.feat = torch.empty(Q, base.E, **f32)
        self.inv_norm = torch.empty(Q, **f32)
        self.cos = torch.empty(Q, C, **f32)
        self.logits = torch.empty(Q, C, **f32)
        self.dlogits = torch.empty(Q, C, **f32)
        self.topk_idx = torch.empty(cfg.tta_steps, Q, K, **i32)
        self.scores = torch.empty(cfg.tta_steps, Q, K, **f32)
        self.rewards = torch.empty(cfg.tta_steps, Q, K, **f32)
        self.loss = torch.empty(cfg.tta_steps, Q, **f32)
        self.reward_feat = torch.empty(Q, reward.E, **f32)
        self.score_rows = torch.empty(Q, C, **f32)
        self.tokens = torch.zeros(Q, L, dtype=torch.int64, device=dev)
        self.eot_rows = torch.empty(Q, **i32)
        self.reot_rows = torch.empty(Q, **i32)
        self._graph = None
        self._static_images = None
        self.fused_adamw = FT.FUSED_ADAMW

    # ------------------------------------------------------------------
    def momentum_update(self, q: int = 0):
        """CLIPRet_TTA.momentum_update_model + reset_initial (custom_models.py:122-142) after query q.  Rows of
        token_embedding the caption does not use only see AdamW's decoupled weight decay (zero gradient, zero
        moments): p *= (1 - lr*wd) per step."""
        c, lay = self.cfg, self.lay
        if not c.momentum_update:
            return
        self.update_counter += 1
        a = 1.0 - c.momentum
        self.ema_ln.mul_(c.momentum).add_(self.ln[q], alpha=a)
        self.ema_rest[:lay.shared].mul_(c.momentum).add_(self.rest[q, :lay.shared], alpha=a)
        self.ema_rest[lay.ls].mul_(c.momentum).add_(self.rest[q, lay.ls], alpha=a)
        tab = self.init_tab.clone()
        for _ in range(c.tta_steps):
            tab.mul_(1.0 - c.lr * c.weight_decay)
        tab[self.tokens[q]] = self.rest[q, lay.tok:lay.tok + lay.L * lay.d].view(lay.L, lay.d)
        self.ema_tab.mul_(c.momentum).add_(tab, alpha=a)
        if self.update_counter >= c.update_freq:
            self.update_counter = 0
            w = c.update_w
            torch.add(self.clip_ln * (1 - w), self.ema_ln, alpha=w, out=self.init_ln)
            torch.add(self.clip_rest * (1 - w), self.ema_rest, alpha=w, out=self.init_rest)
            torch.add(self.clip_tab * (1 - w), self.ema_tab, alpha=w, out=self.init_tab)

    def _logits(self, out: torch.Tensor):
        Q, C, base, lay = self.n_query, self.gallery.shape[0], self.base, self.lay
        ops.pair_logits(self.feat, self.gallery, 0, Q, 1, C, base.E, 1.0, self.cos)
        ops.scale_rows_exp(self.cos, self.rest.view(-1)[lay.ls:], lay.total, out)

    def _step(self, step: int):
        cfg, Q, lay, base, hk, w = self.cfg, self.n_query, self.lay, self.base, self.hook, self.gw
        P, K = base.P, cfg.sample_k
        xs = self.run.forward(Q, self.ln, pstride=P, seqs_per_set=1, prompt=self.rows_in, store=self.store, w=w)
        self.run.head(xs, Q, self.ln, pstride=P, seqs_per_set=1, row_idx=self.eot_rows, feat=self.feat,
                      inv_norm=self.inv_norm, w=w)
        self._logits(self.logits)                                               # logits_per_text
        ops.retrieval_loss(self.logits, self.reward_feat, self.reward_gallery, K, self.dlogits,
                           clipscore_weight=cfg.clipscore_weight, reward_process=cfg.reward_process,
                           amplify=cfg.reward_amplify, loss_scale=cfg.loss_scale, topk_idx=self.topk_idx[step - 1],
                           scores=self.scores[step - 1], rewards=self.rewards[step - 1], loss=self.loss[step - 1])
        flat_g = self.grads.view(-1)
        ops.rowdot(self.dlogits, self.logits, flat_g[lay.ls:], out_stride=lay.total)      # d logit_scale
        ops.scale_rows_exp(self.dlogits, self.rest.view(-1)[lay.ls:], lay.total, self.dlogits)   # d cos
        ops.dfeat_partial(self.dlogits, self.gallery, self.df_partial)
        self.partials.zero_()
        self.run.dres[:Q * base.L].zero_()
        off = base.ln_off("ln_final")
        lnv = self.ln.view(-1)
        ops.head_bwd_ex(self.ones, (0, 0, 1), xs, lnv[off:], w.proj, self.df_partial, self.n_chunks * base.E, 1.0,
                        self.feat, self.inv_norm, Q, 1, base.d, base.E, self.n_chunks, self.run.dres,
                        row_idx=self.eot_rows, param_stride=P, partials=self.partials, n_slots=self.n_slots,
                        p_total=P, p_off=off, beta=lnv[off + base.d:], y_out=hk.y, df_out=hk.df,
                        proj_stride=w.proj_stride)
        fz = FT.FusedAdamw(self.rest, self.rest_m, self.rest_v, self.w16, self.rest, lay.total, step == 1, cfg,
                           step) if self.fused_adamw else None
        hk.bind(self.grads, Q, None, fused=fz)
        hk.proj()
        self.run.backward(self.store, Q, 1, self.ln, P, self.partials, self.n_slots, w=w, hook=hk)
        ops.tied_rows_grad(self.run.dres, self.tokens, Q, base.L, base.d, flat_g[lay.tok:], flat_g[lay.pos:], lay.total)
        kw = dict(beta1=cfg.betas[0], beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
                  loss_scale=cfg.loss_scale)
        ops.adamw_step(self.ln, self.ln_m, self.ln_v, self.partials, Q, self.n_slots, P, cfg.lr, step,
                       grad_out=self.ln_grad, **kw)
        FT.adamw_weights(lay, self.rest, self.rest_m, self.rest_v, self.grads, self.w16,
                         self.w16t if step < cfg.tta_steps else None, cfg, step, self.rest, lay.total, step == 1,
                         fused=self.fused_adamw)

    def tune(self, tokens: torch.Tensor):
        """tune_text for n_query tokenised captions [Q, 77] (int64) at once."""
        cfg, Q, lay, base = self.cfg, self.n_query, self.lay, self.base
        if tuple(tokens.shape) != (Q, base.L):
            raise RlcfError(f"expected tokens [{Q}, {base.L}], got {tuple(tokens.shape)}")
        self.tokens.copy_(tokens)
        eot = self.tokens.argmax(dim=-1).to(torch.int32)
        torch.add(torch.arange(Q, device=eot.device, dtype=torch.int32) * base.L, eot, out=self.eot_rows)
        # reward_model.set_text_features(captions=text)                          (clip_ret_policy.py:117)
        rw = self.reward
        xr = self.rrun.forward(Q, rw.ln_flat, tokens=self.tokens)
        torch.add(torch.arange(Q, device=eot.device, dtype=torch.int32) * rw.L, eot, out=self.reot_rows)
        self.rrun.head(xr, Q, rw.ln_flat, row_idx=self.reot_rows, feat=self.reward_feat)
        # reset_initial + the caption's own embedding rows
        ops.reset_params(self.init_ln, self.ln, self.ln_m, self.ln_v, Q, base.P)
        self.rest.copy_(self.init_rest.expand_as(self.rest))
        self.rest[:, lay.tok:lay.tok + base.L * base.d].copy_(self.init_tab[self.tokens].view(Q, -1))
        FT.cast_weights(lay, self.rest, self.w16, self.w16t)
        for step in range(1, cfg.tta_steps + 1):
            self._step(step)
        return self.ln, self.rest

    def predict(self, tokens: torch.Tensor | None = None) -> torch.Tensor:
        Q, P = self.n_query, self.base.P
        xf = self.run.forward(Q, self.ln, pstride=P, seqs_per_set=1, prompt=self.rows_in, w=self.gw)
        self.run.head(xf, Q, self.ln, pstride=P, seqs_per_set=1, row_idx=self.eot_rows, feat=self.feat, w=self.gw)
        self._logits(self.score_rows)
        return self.score_rows

    def adapt(self, tokens: torch.Tensor) -> torch.Tensor:
        self.tune(tokens)
        return self.predict()

    @property
    def logits_final(self):   # name shared with the classification engines (graph helpers)
        return self.score_rows

    capture = E.RlcfEngine.capture            # the whole adaptation of a token batch as one CUDA graph
    adapt_graph = E.RlcfEngine.adapt_graph

    def export_params(self, q: int) -> dict:
        """Adapted parameters of query q under the reference's state-dict names; `token_embedding.rows` holds the
        caption's own L rows (row t belongs to token id tokens[q, t])."""
        lay, base = self.lay, self.base
        out = {}
        for key, off, shape in lay.entries():
            n = 1
            for s_ in shape:
                n *= s_
            out[key] = self.rest[q, off:off + n].view(shape)
        for key, off in base.ln_names(""):
            out[key] = self.ln[q, off:off + base.d]
        out["logit_scale"] = self.rest[q, lay.ls]
        out["token_embedding.rows"] = self.rest[q, lay.tok:lay.tok + lay.L * lay.d].view(lay.L, lay.d)
        return out

    def bytes_per_query(self) -> float:
        return _bytes_per_query(self.lay, self.cfg.tta_steps)
