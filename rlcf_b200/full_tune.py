"""Full image-encoder tuning (TPT/tune_cls_rl.py with --tune_norm 0, the default recipe of scripts/rlcf-tune.sh):
every parameter of the visual tower is trainable (custom_clip.py:477-479), so each test image ends up with its own
86 M weights after the first AdamW step.

Step 1 still runs batched over all images on the shared initial weights (forward of all views, training forward +
dgrad of the selected views); what is new is
  * weight gradients: per image, dW[out,in] = dY^T X as a tcgen05 GEMM over transposed fp16 copies of dY and X,
    bias gradients as column sums, conv1 / class / positional / projection gradients from the ln_pre input gradient,
  * a per-image fp32 master copy of all non-LayerNorm parameters with Adam moments (the first step reads the shared
    initial copy, so the reference's 345 MB per-image state-dict restore never happens),
  * per-image fp16 weight copies, used by the remaining TTA steps and the adapted prediction, which therefore run one
    image at a time (small-M GEMMs, weight-read bound).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import engine as E
from . import ops
from ._lib import EPI_F32, RlcfError


class FullLayout:
    """Offsets of the non-LayerNorm visual parameters inside one flat fp32 vector.
    [ GEMM block: conv1 [d,k_pad] | per layer in_proj_w, out_proj_w, c_fc_w, c_proj_w ]  (cast to fp16 in one go)
    [ class_embedding | positional_embedding | proj | per layer in_proj_b, out_proj_b, c_fc_b, c_proj_b ]"""

    def __init__(self, w: E.TowerWeights):
        d, L, Ed, nl, kp = w.d, w.L, w.E, w.n_layers, w.k_pad
        self.d, self.L, self.E, self.nl, self.k_pad = d, L, Ed, nl, kp
        off = 0
        self.conv = off; off += d * kp
        self.wq, self.wo, self.wf, self.wp = [], [], [], []
        for _ in range(nl):
            self.wq.append(off); off += 3 * d * d
            self.wo.append(off); off += d * d
            self.wf.append(off); off += 4 * d * d
            self.wp.append(off); off += 4 * d * d
        self.p_gemm = off
        self.cls = off; off += d
        self.pos = off; off += L * d
        self.proj = off; off += d * Ed
        self.bq, self.bo, self.bf, self.bp = [], [], [], []
        for _ in range(nl):
            self.bq.append(off); off += 3 * d
            self.bo.append(off); off += d
            self.bf.append(off); off += 4 * d
            self.bp.append(off); off += d
        self.total = off

    def entries(self, prefix="visual."):
        """(state_dict key, offset, shape) of every entry, conv1 in its flattened zero-padded [d, k_pad] form."""
        d, nl = self.d, self.nl
        out = [(prefix + "conv1.weight", self.conv, (d, self.k_pad)), (prefix + "class_embedding", self.cls, (d,)),
               (prefix + "positional_embedding", self.pos, (self.L, d)), (prefix + "proj", self.proj, (d, self.E))]
        for l in range(nl):
            rb = f"{prefix}transformer.resblocks.{l}."
            out += [(rb + "attn.in_proj_weight", self.wq[l], (3 * d, d)), (rb + "attn.out_proj.weight", self.wo[l], (d, d)),
                    (rb + "mlp.c_fc.weight", self.wf[l], (4 * d, d)), (rb + "mlp.c_proj.weight", self.wp[l], (d, 4 * d)),
                    (rb + "attn.in_proj_bias", self.bq[l], (3 * d,)), (rb + "attn.out_proj.bias", self.bo[l], (d,)),
                    (rb + "mlp.c_fc.bias", self.bf[l], (4 * d,)), (rb + "mlp.c_proj.bias", self.bp[l], (d,))]
        return out


def pack_rest(sd: dict, lay: FullLayout, prefix="visual.") -> torch.Tensor:
    dev = sd[prefix + "proj"].device
    flat = torch.zeros(lay.total, dtype=torch.float32, device=dev)
    for key, off, shape in lay.entries(prefix):
        src = sd[key].detach().float()
        if key.endswith("conv1.weight"):
            k_real = src[0].numel()
            flat[off:off + shape[0] * shape[1]].view(shape)[:, :k_real].copy_(src.reshape(shape[0], k_real))
        else:
            flat[off:off + src.numel()].copy_(src.reshape(-1))
    return flat


def tower_view(base: E.TowerWeights, lay: FullLayout, rest: torch.Tensor, w16: torch.Tensor, w16t, ln_flat):
    """TowerWeights whose tensors are views into one image's parameter vectors (fp32 `rest`, fp16 `w16`/`w16t`)."""
    d = lay.d
    t = E.TowerWeights(kind="visual", d=d, heads=base.heads, n_layers=lay.nl, L=lay.L, E=lay.E, patch=base.patch,
                       resolution=base.resolution, k_pad=lay.k_pad)
    t.conv_w = w16[lay.conv:lay.conv + d * lay.k_pad].view(d, lay.k_pad)
    t.cls = rest[lay.cls:lay.cls + d]
    t.pos = rest[lay.pos:lay.pos + lay.L * d].view(lay.L, d)
    t.proj = rest[lay.proj:lay.proj + d * lay.E].view(d, lay.E)
    t.ln_flat = ln_flat
    for l in range(lay.nl):
        def m(buf, off, r, c):
            return None if buf is None else buf[off:off + r * c].view(r, c)
        t.layers.append(E.LayerWeights(
            wqkv=m(w16, lay.wq[l], 3 * d, d), bqkv=rest[lay.bq[l]:lay.bq[l] + 3 * d],
            wo=m(w16, lay.wo[l], d, d), bo=rest[lay.bo[l]:lay.bo[l] + d],
            wfc=m(w16, lay.wf[l], 4 * d, d), bfc=rest[lay.bf[l]:lay.bf[l] + 4 * d],
            wproj=m(w16, lay.wp[l], d, 4 * d), bproj=rest[lay.bp[l]:lay.bp[l] + d],
            wqkv_t=m(w16t, lay.wq[l], d, 3 * d), wo_t=m(w16t, lay.wo[l], d, d),
            wfc_t=m(w16t, lay.wf[l], d, 4 * d), wproj_t=m(w16t, lay.wp[l], 4 * d, d)))
    return t


def tower_view_grouped(base: E.TowerWeights, lay: FullLayout, rest: torch.Tensor, w16: torch.Tensor, w16t, ln):
    """TowerWeights over ALL samples' parameter vectors at once (rest [G,total] fp32, w16 / w16t [G,p_gemm] fp16,
    ln [G,P]): GEMM weights become [G,out,in] views for the grouped tcgen05 launch, biases [G,n] views, and the
    embedding / projection tensors carry the per-sample stride."""
    d, G = lay.d, rest.shape[0]
    t = E.TowerWeights(kind="visual", d=d, heads=base.heads, n_layers=lay.nl, L=lay.L, E=lay.E, patch=base.patch,
                       resolution=base.resolution, k_pad=lay.k_pad)

    def m(buf, off, r, c):
        return None if buf is None else buf.as_strided((G, r, c), (buf.stride(0), c, 1), buf.storage_offset() + off)

    def b(off, n):
        return rest.as_strided((G, n), (rest.stride(0), 1), rest.storage_offset() + off)

    t.conv_w = m(w16, lay.conv, d, lay.k_pad)
    t.cls = rest[0, lay.cls:lay.cls + d]
    t.pos = rest[0, lay.pos:lay.pos + lay.L * d].view(lay.L, d)
    t.proj = rest[0, lay.proj:lay.proj + d * lay.E].view(d, lay.E)
    t.embed_stride = t.proj_stride = rest.stride(0)
    t.ln_flat = ln[0]   # P = one sample's LayerNorm slice; callers pass the [G, P] tensor and its stride explicitly
    for l in range(lay.nl):
        t.layers.append(E.LayerWeights(
            wqkv=m(w16, lay.wq[l], 3 * d, d), bqkv=b(lay.bq[l], 3 * d),
            wo=m(w16, lay.wo[l], d, d), bo=b(lay.bo[l], d),
            wfc=m(w16, lay.wf[l], 4 * d, d), bfc=b(lay.bf[l], 4 * d),
            wproj=m(w16, lay.wp[l], d, 4 * d), bproj=b(lay.bp[l], d),
            wqkv_t=m(w16t, lay.wq[l], d, 3 * d), wo_t=m(w16t, lay.wo[l], d, d),
            wfc_t=m(w16t, lay.wf[l], d, 4 * d), wproj_t=m(w16t, lay.wp[l], 4 * d, d)))
    return t


def cast_weights(lay: FullLayout, rest: torch.Tensor, w16: torch.Tensor, w16t=None):
    """fp16 copies of every sample's GEMM weights (and their transposes, the dgrad B operands) from the fp32 masters."""
    G, d = rest.shape[0], lay.d
    ops.call("rlcf_cast_f16", ops.ptr(rest), G, lay.p_gemm, rest.stride(0), ops.ptr(w16), w16.stride(0), ops.stream())
    if w16t is None:
        return
    for l in range(lay.nl):
        for off, r, c in ((lay.wq[l], 3 * d, d), (lay.wo[l], d, d), (lay.wf[l], 4 * d, d), (lay.wp[l], d, 4 * d)):
            ops.transpose_cast_f16_sets(rest.view(-1)[off:], r, c, G, rest.stride(0), w16t.view(-1)[off:], w16t.stride(0))


def transpose_weights(lay, w16: torch.Tensor, w16t: torch.Tensor):
    """w16t[g] = per-matrix transposes of the fp16 weights w16[g] (dgrad B operands), for all samples at once."""
    G, d = w16.shape[0], lay.d
    for l in range(lay.nl):
        for off, r, c in ((lay.wq[l], 3 * d, d), (lay.wo[l], d, d), (lay.wf[l], 4 * d, d), (lay.wp[l], d, 4 * d)):
            ops.transpose_f16_sets(w16.view(-1)[off:], r, c, w16t.view(-1)[off:], G, w16.stride(0))


# Default for the engines' `fused_adamw`.  True: the weight-gradient GEMMs apply AdamW in their epilogue
# (rlcf_gemm_wgrad_adamw), so the gradient of a GEMM weight never reaches HBM -- 26 B instead of 34 B of traffic per
# weight and step, bit-identical results.  Measured on a B200 (round 1) the epilogue-driven streaming reaches 3.4 TB/s
# against 5.5 TB/s for the dedicated streaming pass (only the 8 epilogue warps of a CTA have loads in flight), so the
# unfused sequence wgrad GEMM -> rlcf_adamw_full is still faster end to end (retrieval image->text 130 vs 110 queries/s)
# and stays the default until the epilogue prefetches its tiles with TMA.
FUSED_ADAMW = False


@dataclass
class FusedAdamw:
    """What the fused wgrad + AdamW epilogue needs besides the GEMM operands (one optimizer step of all samples)."""
    rest: torch.Tensor          # fp32 masters [G, total] (written)
    m: torch.Tensor
    v: torch.Tensor
    w16: torch.Tensor           # fp16 GEMM copies [G, p_gemm] (written)
    p_in: torch.Tensor          # where the parameters are read: the shared initial vector (stride 0) or `rest`
    p_in_gs: int
    fresh: bool
    cfg: object
    step: int


def adamw_weights(lay, rest, rest_m, rest_v, grads, w16, w16t, cfg, step, params_in, params_in_stride, fresh,
                  fused: bool = False):
    """AdamW over every sample's non-LayerNorm parameters + refreshed fp16 GEMM copies in the same pass; the transposed
    copies follow only when another backward will use them (w16t not None).  fused=True: the GEMM weights were already
    updated by the wgrad epilogues (FusedAdamw), only the rest of the vector (embeddings, projection, biases) is left."""
    kw = dict(beta1=cfg.betas[0], beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
              loss_scale=cfg.loss_scale)
    if fused:
        g, tot = lay.p_gemm, rest.shape[1]
        ops.adamw_full(rest.view(-1)[g:], rest_m.view(-1)[g:], rest_v.view(-1)[g:], grads.view(-1)[g:], rest.shape[0],
                       tot - g, cfg.lr, step, params_in.view(-1)[g:], params_in_stride, fresh, set_stride=tot, **kw)
    else:
        ops.adamw_full(rest, rest_m, rest_v, grads, rest.shape[0], rest.shape[1], cfg.lr, step, params_in,
                       params_in_stride, fresh, w16=w16, n16=lay.p_gemm, **kw)
    if w16t is not None:
        transpose_weights(lay, w16, w16t)


class WgradHook:
    """Weight / bias / embedding gradients of `n_sets` images whose rows are laid out set after set.
    grads: fp32 [n_sets, lay.total] (written, not accumulated).  dY carries the loss scale; so do the gradients."""

    def __init__(self, lay: FullLayout, w: E.TowerWeights, n_sets_max: int, seqs_per_set: int, device):
        self.lay, self.w, self.S = lay, w, seqs_per_set
        d, L = lay.d, lay.L
        self.rows = seqs_per_set * L
        self.rows_pad = E._round_up(self.rows, 8)
        self.prow = seqs_per_set * (L - 1)
        self.prow_pad = E._round_up(self.prow, 8)
        f16 = dict(dtype=torch.float16, device=device)
        self.t_dy = torch.empty(4 * d, n_sets_max * self.rows_pad, **f16)
        self.t_x = torch.empty(max(4 * d, lay.k_pad), n_sets_max * self.rows_pad, **f16)
        self.dx_pre = torch.empty(n_sets_max * self.rows, d, dtype=torch.float32, device=device)
        self.y = torch.empty(n_sets_max * seqs_per_set, d, dtype=torch.float32, device=device)
        self.df = torch.empty(n_sets_max * seqs_per_set, lay.E, dtype=torch.float32, device=device)
        self.grads = None
        self.n_sets = 0
        self.patches = None
        self.fused = None

    def bind(self, grads: torch.Tensor, n_sets: int, patches: torch.Tensor, fused: "FusedAdamw | None" = None):
        self.grads, self.n_sets, self.patches, self.fused = grads, n_sets, patches, fused

    def _wgrad(self, dY, X, n_out, n_in, off, rows, rows_pad, skip=0, dy_stride_rows=None, bias_off=None):
        """grads[g, off:...] = dY_g^T X_g for every set g (one grouped GEMM over transposed fp16 operands); with
        bias_off the column sums of dY (the bias gradient) come out of the same transposing pass."""
        ns = self.n_sets
        ld = ns * rows_pad
        g = self.grads
        if dY.dtype == torch.float16:
            ops.transpose_blocks_colsum(dY, ns, rows, rows_pad, n_out, self.t_dy, ld, skip_first=skip,
                                        in_set_stride_rows=dy_stride_rows,
                                        colsum=None if bias_off is None else g.view(-1)[bias_off:],
                                        colsum_stride=g.stride(0))
        else:
            ops.transpose_blocks(dY, ns, rows, rows_pad, n_out, self.t_dy, ld, skip_first=skip,
                                 in_set_stride_rows=dy_stride_rows)
        ops.transpose_blocks_colsum(X, ns, rows, rows_pad, n_in, self.t_x, ld)
        # one grouped launch: group g = columns [g*rows_pad, (g+1)*rows_pad) of the transposed operands
        ty = self.t_dy.as_strided((ns, n_out, rows_pad), (rows_pad, ld, 1))
        tx = self.t_x.as_strided((ns, n_in, rows_pad), (rows_pad, ld, 1))
        fz = self.fused
        if fz is None:
            out = g.as_strided((ns, n_out, n_in), (g.stride(0), n_in, 1), g.storage_offset() + off)
            ops.gemm_grouped(ty, tx, out, epilogue=EPI_F32)
            return

        def view(t):   # this weight's [n_out, n_in] tile of every sample's flat vector
            return t.as_strided((ns, n_out, n_in), (t.stride(0), n_in, 1), t.storage_offset() + off)
        c = fz.cfg
        ops.gemm_wgrad_adamw(ty, tx, view(fz.rest), view(fz.m), view(fz.v), fz.p_in.view(-1)[off:], fz.p_in_gs, fz.fresh,
                             view(fz.w16), c.lr, fz.step, beta1=c.betas[0], beta2=c.betas[1], eps=c.eps,
                             weight_decay=c.weight_decay, loss_scale=c.loss_scale)

    def linear(self, l: int, name: str, dY: torch.Tensor, X: torch.Tensor):
        lay, d = self.lay, self.lay.d
        n_out, n_in, woff, boff = {"in_proj": (3 * d, d, lay.wq[l], lay.bq[l]), "out_proj": (d, d, lay.wo[l], lay.bo[l]),
                                   "c_fc": (4 * d, d, lay.wf[l], lay.bf[l]),
                                   "c_proj": (d, 4 * d, lay.wp[l], lay.bp[l])}[name]
        self._wgrad(dY, X, n_out, n_in, woff, self.rows, self.rows_pad, bias_off=boff)

    def embed(self, dx_pre: torch.Tensor):
        lay, d, L = self.lay, self.lay.d, self.lay.L
        ns = self.n_sets
        # positional embedding (row 0 of it is also the class-embedding gradient: x_pre[v,0] = cls + pos[0])
        ops.seq_sum(dx_pre, ns, self.S, L, d, self.grads[:, lay.pos:], self.grads.stride(0))
        self.grads[:ns, lay.cls:lay.cls + d].copy_(self.grads[:ns, lay.pos:lay.pos + d])
        # conv1: d patch_out = dx_pre without the class-token rows; X = the im2col patches of the same views
        self._wgrad(dx_pre, self.patches, d, lay.k_pad, lay.conv, self.prow, self.prow_pad, skip=L,
                    dy_stride_rows=self.rows)

    def proj(self):
        lay = self.lay
        ops.outer_sum(self.y, self.df, self.n_sets, self.S, lay.d, lay.E, self.grads[:, lay.proj:], self.grads.stride(0))


class FullTuneEngine:
    """Batched RLCF adaptation of the whole image encoder for `n_img` independent test images."""

    def __init__(self, sd_visual: dict, class_feat: torch.Tensor, logit_scale: float, cfg: E.RlcfConfig, n_img: int,
                 reward: E.TowerWeights, reward_class_feat: torch.Tensor, prefix: str = "visual."):
        self.cfg, self.n_img = cfg, n_img
        self.base = E.prepare_visual(sd_visual, prefix=prefix, need_grad=True)
        pol = self.base
        dev = pol.ln_flat.device
        self.lay = lay = FullLayout(pol)
        self.class_feat = class_feat.float().contiguous()
        self.logit_scale = float(logit_scale)
        self.reward = reward
        B, V, S, C = n_img, cfg.n_views, cfg.n_selected, self.class_feat.shape[0]
        if S < 1:
            raise RlcfError(f"int(n_views * selection_p) = {S}: no view would be selected (tpt_cls_rl.py:34)")
        if cfg.loss != "rlcf":
            raise NotImplementedError("full tuning implements the RLCF loss")
        P = pol.P
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.run = E.TowerRunner(pol, B * V)
        self.run.reserve_backward(B * S)
        self.store = E.ActStore(pol, B * S, dev, full=True)
        self.scorer = E.RewardScorer(reward, reward_class_feat, B * S, cfg.reward_weights)
        self.hook = WgradHook(lay, pol, B, S, dev)
        # parameters: LayerNorm slice as in RlcfEngine, everything else in `rest`
        self.init_ln = pol.ln_flat.clone()
        self.init_rest = pack_rest(sd_visual, lay, prefix)
        self.ln = torch.empty(B, P, **f32); self.ln_m = torch.empty(B, P, **f32); self.ln_v = torch.empty(B, P, **f32)
        self.n_slots = max(E.N_SLOTS, S)
        self.partials = torch.empty(B, self.n_slots, P, **f32)
        self.ln_grad = torch.empty(B, P, **f32)
        self.rest = torch.empty(B, lay.total, **f32)
        self.rest_m = torch.empty(B, lay.total, **f32)
        self.rest_v = torch.empty(B, lay.total, **f32)
        self.grads = torch.zeros(B, lay.total, **f32)
        self.w16 = torch.empty(B, lay.p_gemm, dtype=torch.float16, device=dev)
        self.w16t = torch.empty(B, lay.p_gemm, dtype=torch.float16, device=dev) if cfg.tta_steps > 1 else None
        # all images' weights as one set of [B, out, in] views: steps >= 2 and the adapted prediction are grouped launches
        self.gw = tower_view_grouped(pol, lay, self.rest, self.w16, self.w16t, self.ln)
        self.logits_all = torch.empty(B * V, C, **f32)
        self.entropy = torch.empty(B, V, **f32)
        self.sel = torch.empty(B, S, **i32)
        self.sel_global = torch.empty(B * S, **i32)
        self.first_view = (torch.arange(B, device=dev, dtype=torch.int32) * V).contiguous()
        self.logits_sel = torch.empty(B * S, C, **f32)
        self.feat_sel = torch.empty(B * S, pol.E, **f32)
        self.inv_norm_sel = torch.empty(B * S, **f32)
        self.dlogits = torch.empty(B * S, C, **f32)
        self.topk_idx = torch.empty(B * S, cfg.sample_k, **i32)
        self.scores = torch.empty(B * S, cfg.sample_k, **f32)
        self.rewards = torch.empty(B * S, cfg.sample_k, **f32)
        self.loss = torch.empty(cfg.tta_steps, B, **f32)
        self.logits_final = torch.empty(B, C, **f32)
        self._graph = None
        self._static_images = None
        self.fused_adamw = FUSED_ADAMW
        E.setup_view_store(self, pol, self.run, B, V, S, dev)     # step 1 adopts the all-views pass's activations

    def sync_initial_state(self, sd_visual: dict, prefix: str = "visual.") -> bool:
        """Makes the engine start from `sd_visual` (the model's current reset state).  A no-op -- one packed copy and
        a comparison, ~0.3 ms for ViT-B/16 -- when the weights equal the snapshot the engine already holds, which is
        the case for every image of a dataset unless --momentum_update folds the EMA back
        (custom_clip.py:460-475); otherwise the fp16 GEMM copies and the snapshot are refreshed IN PLACE (buffers,
        grouped views and the captured graph stay valid).  Returns True when a refresh happened."""
        rest = pack_rest(sd_visual, self.lay, prefix)
        ln = torch.cat([sd_visual[k].detach().float().reshape(-1) for k, _ in self.base.ln_names(prefix)])
        if torch.equal(rest, self.init_rest) and torch.equal(ln, self.init_ln):
            return False
        E.copy_tower_(self.base, E.prepare_visual(sd_visual, prefix=prefix, need_grad=True))
        self.init_rest.copy_(rest)
        self.init_ln.copy_(ln)
        return True

    def _fused(self, step):
        if not self.fused_adamw:
            return None
        first = step == 1
        return FusedAdamw(self.rest, self.rest_m, self.rest_v, self.w16, self.init_rest if first else self.rest,
                          0 if first else self.lay.total, first, self.cfg, step)

    # ------------------------------------------------------------------
    def _loss_and_head_bwd(self, step, xs, runner, ln, pstride, n_sets, w, rows0=0):
        """reward loss + head backward for sets [rows0, rows0 + n_sets) whose tower outputs are in xs."""
        cfg = self.cfg
        S, K, C = cfg.n_selected, cfg.sample_k, self.class_feat.shape[0]
        pol = self.base
        sl = slice(rows0 * S, (rows0 + n_sets) * S)
        runner.head(xs, n_sets * S, ln, pstride=pstride, seqs_per_set=S, class_feat=self.class_feat,
                    logit_scale=self.logit_scale, feat=self.feat_sel[sl], inv_norm=self.inv_norm_sel[sl],
                    logits=self.logits_sel[sl], w=w)
        self.scorer.loss(self.logits_sel[sl], n_sets, S, K, C, self.dlogits[sl], cfg, topk_idx=self.topk_idx[sl],
                         scores=self.scores[sl], rewards=self.rewards[sl], loss=self.loss[step - 1][rows0:rows0 + n_sets])
        if cfg.min_entropy_w:                                                          # tpt_cls_rl.py:73-74
            ops.avg_entropy_reg(self.logits_sel[sl], None, n_sets, S, C, self.dlogits[sl], cfg.min_entropy_w,
                                loss=self.loss[step - 1][rows0:rows0 + n_sets], loss_scale=cfg.loss_scale)
        off = pol.ln_off("ln_post")
        lnv = ln.view(-1)
        runner.dres[:n_sets * S * pol.L].zero_()
        hk = self.hook
        ops.head_bwd_ex(self.dlogits[sl], (S * C, C, 1), xs, lnv[off:], w.proj, self.class_feat, 0, self.logit_scale,
                        self.feat_sel[sl], self.inv_norm_sel[sl], n_sets, S, pol.d, pol.E, C, runner.dres,
                        row_stride=pol.L, param_stride=pstride, partials=self.partials[rows0:rows0 + n_sets],
                        n_slots=self.n_slots, p_total=pol.P, p_off=off, beta=lnv[off + pol.d:], y_out=hk.y,
                        df_out=hk.df, proj_stride=w.proj_stride)

    def _adamw(self, step):
        cfg, B, P, lay = self.cfg, self.n_img, self.base.P, self.lay
        kw = dict(beta1=cfg.betas[0], beta2=cfg.betas[1], eps=cfg.eps, weight_decay=cfg.weight_decay,
                  loss_scale=cfg.loss_scale)
        ops.adamw_step(self.ln, self.ln_m, self.ln_v, self.partials, B, self.n_slots, P, cfg.lr, step,
                       grad_out=self.ln_grad, **kw)
        # step 1: parameters come from the shared initial copy, moments start from zero -- no per-image reset.
        # The same pass writes the fp16 copies of the updated GEMM weights; transposes only if another backward follows.
        w16t = self.w16t if step < cfg.tta_steps else None
        if step == 1:
            adamw_weights(lay, self.rest, self.rest_m, self.rest_v, self.grads, self.w16, w16t, cfg, step,
                          self.init_rest, 0, True, fused=self.fused_adamw)
        else:
            adamw_weights(lay, self.rest, self.rest_m, self.rest_v, self.grads, self.w16, w16t, cfg, step,
                          self.rest, lay.total, False, fused=self.fused_adamw)

    def tune(self, images: torch.Tensor):
        cfg, B = self.cfg, self.n_img
        V, S, C = cfg.n_views, cfg.n_selected, self.class_feat.shape[0]
        pol, P, hk = self.base, self.base.P, self.hook
        if images.shape[0] != B * V:
            raise RlcfError(f"expected {B * V} views, got {images.shape[0]}")
        ops.reset_params(self.init_ln, self.ln, self.ln_m, self.ln_v, B, P)
        E.all_views_pass(self, self.run, self.init_ln, images, B, V, S, C)
        self.scorer.features(images, self.sel_global, B * S)
        # ---- step 1: all images on the shared initial weights
        if self.views is not None:      # the selected views' activations were adopted from the all-views pass
            xs = self.run.complete(self.store, B * S, self.ln, pstride=P, seqs_per_set=S, images=images,
                                   view_idx=self.sel_global)
        else:
            xs = self.run.forward(B * S, self.ln, pstride=P, seqs_per_set=S, images=images, view_idx=self.sel_global,
                                  store=self.store)
        self.partials.zero_()
        self._loss_and_head_bwd(1, xs, self.run, self.ln, P, B, pol)
        hk.bind(self.grads, B, self.run.patches, fused=self._fused(1))
        hk.proj()
        self.run.backward(self.store, B, S, self.ln, P, self.partials, self.n_slots, hook=hk)
        self._adamw(1)
        # ---- steps >= 2: every image on its own weights (grouped GEMMs, one group per image)
        for step in range(2, cfg.tta_steps + 1):
            self.partials.zero_()
            xs = self.run.forward(B * S, self.ln, pstride=P, seqs_per_set=S, images=images, view_idx=self.sel_global,
                                  store=self.store, w=self.gw)
            self._loss_and_head_bwd(step, xs, self.run, self.ln, P, B, self.gw)
            hk.bind(self.grads, B, self.run.patches, fused=self._fused(step))
            hk.proj()
            self.run.backward(self.store, B, S, self.ln, P, self.partials, self.n_slots, w=self.gw, hook=hk)
            self._adamw(step)
        return self.ln, self.rest

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        B, P = self.n_img, self.base.P
        xf = self.run.forward(B, self.ln, pstride=P, seqs_per_set=1, images=images, view_idx=self.first_view, w=self.gw)
        self.run.head(xf, B, self.ln, pstride=P, seqs_per_set=1, class_feat=self.class_feat,
                      logit_scale=self.logit_scale, logits=self.logits_final, w=self.gw)
        return self.logits_final

    def adapt(self, images: torch.Tensor) -> torch.Tensor:
        self.tune(images)
        return self.predict(images)

    capture = E.RlcfEngine.capture
    adapt_graph = E.RlcfEngine.adapt_graph
    adapt_host = E.RlcfEngine.adapt_host
    host_pipeline = E.RlcfEngine.host_pipeline
    reward_feat = E.RlcfEngine.reward_feat

    def algorithmic_flops_per_image(self) -> float:
        """SURVEY.md 8(d), full tuning: backward = dgrad + wgrad over the selected views."""
        cfg, w = self.cfg, self.base
        V, S = cfg.n_views, cfg.n_selected
        f = E.RlcfEngine.tower_fwd_flops(w)
        wgrad = w.n_layers * 24 * w.L * w.d * w.d + 2 * (w.L - 1) * w.d * 3 * w.patch * w.patch + 2 * w.d * w.E
        bwd = E.RlcfEngine.tower_dgrad_flops(w) + wgrad
        fi = E.RlcfEngine.tower_fwd_flops(w, cls_only_last=True)   # the V-view forward with the shared initial weights
        total = V * fi + S * bwd + f + (cfg.tta_steps - 1) * S * (f + bwd) + S * self.scorer.fwd_flops()
        return float(total)

    def reference_flops_per_image(self) -> float:
        """The same count with every block of every forward on every token, as the reference executes it."""
        cfg, w = self.cfg, self.base
        V, S = cfg.n_views, cfg.n_selected
        f = E.RlcfEngine.tower_fwd_flops(w)
        wgrad = w.n_layers * 24 * w.L * w.d * w.d + 2 * (w.L - 1) * w.d * 3 * w.patch * w.patch + 2 * w.d * w.E
        bwd = E.RlcfEngine.tower_dgrad_flops(w) + wgrad
        total = V * f + S * bwd + f + (cfg.tta_steps - 1) * S * (f + bwd)
        total += S * sum(E.RlcfEngine.tower_fwd_flops(t) for t in self.scorer.towers)
        return float(total)

    def export_params(self, b: int, prefix: str = "visual.") -> dict:
        """Adapted parameters of image b under the reference's state-dict names (conv1 in its [d,3,p,p] shape)."""
        lay, pol = self.lay, self.base
        out = {}
        for key, off, shape in lay.entries(prefix):
            n = 1
            for s_ in shape:
                n *= s_
            t = self.rest[b, off:off + n].view(shape)
            if key.endswith("conv1.weight"):
                t = t[:, :3 * pol.patch * pol.patch].reshape(lay.d, 3, pol.patch, pol.patch)
            out[key] = t
        for key, off in pol.ln_names(prefix):
            out[key] = self.ln[b, off:off + pol.d]
        return out
