"""Tensor-level wrappers over the C ABI (one Python function per exported kernel).

These only validate devices/dtypes/contiguity and forward raw pointers; all arithmetic happens in
librlcf_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_F16, EPI_F32, EPI_GELU_BWD_F16, EPI_GELU_F16, EPI_RESID_F32, call, ptr, stream  # noqa: F401


def _chk(t, dtype, name, strided=False):
    """strided=True: the tensor is only a base pointer for a kernel that takes its own stride argument."""
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.RlcfError(f"{name} must be a CUDA tensor (rlcf_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.RlcfError(f"{name} must be {dtype}, got {t.dtype}")
    if not strided and not t.is_contiguous():
        raise _lib.RlcfError(f"{name} must be contiguous")


# When set to a list, every gemm() launch is bracketed by CUDA events on the launching stream and
# (M, N, K, start_event, end_event) is appended: bench.py's live per-launch timing of the dominant kernel.
GEMM_TIMER = None


def gemm(a, b, out, epilogue=EPI_F16, bias=None, resid=None, aux_in=None, aux_out=None, alpha=1.0, M=None):
    """out[M,N] = epilogue(alpha * a[M,K] @ b[N,K]^T).  a, b fp16 row-major (row-strided views allowed: the leading
    dimension is passed to the kernel); M may restrict the rows used."""
    for t, nm in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.float16 or t.dim() != 2 or t.stride(1) != 1:
            raise _lib.RlcfError(f"gemm: {nm} must be a 2-D CUDA fp16 tensor with unit inner stride")
    _chk(bias, torch.float32, "bias"); _chk(resid, torch.float32, "resid")
    _chk(aux_in, torch.float16, "aux_in"); _chk(aux_out, torch.float16, "aux_out")
    m = a.shape[0] if M is None else M
    k = a.shape[1]
    n = b.shape[0]
    if b.shape[1] != k:
        raise _lib.RlcfError(f"gemm: K mismatch {a.shape} vs {b.shape}")
    want = torch.float32 if epilogue in (EPI_RESID_F32, EPI_F32) else torch.float16
    _chk(out, want, "out")
    if out.shape[-1] != n or out.shape[0] < m:
        raise _lib.RlcfError(f"gemm: out shape {tuple(out.shape)} does not fit [{m},{n}]")
    if GEMM_TIMER is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    call("rlcf_gemm_f16", ptr(a), a.stride(0), ptr(b), b.stride(0), m, n, k, epilogue, float(alpha), ptr(bias),
         ptr(resid), ptr(aux_in), ptr(aux_out), ptr(out), out.stride(0), stream())
    if GEMM_TIMER is not None:
        e1.record()
        GEMM_TIMER.append((m, n, k, e0, e1))
    return out


def gemm_grouped(a, b, out, epilogue=EPI_F16, bias=None, resid=None, aux_in=None, aux_out=None, alpha=1.0):
    """G independent GEMMs in one launch: out[g] = epilogue(alpha * a[g] @ b[g]^T).  a [G,M,K], b [G,N,K] fp16 and
    out [G,M,N] are 3-D views with unit inner stride (any row / group strides that are multiples of 8 elements);
    bias [G,N] fp32; resid / aux share out's layout."""
    for t, nm in ((a, "a"), (b, "b"), (out, "out")):
        if not t.is_cuda or t.dim() != 3 or t.stride(2) != 1:
            raise _lib.RlcfError(f"gemm_grouped: {nm} must be a 3-D CUDA tensor with unit inner stride")
    if a.dtype != torch.float16 or b.dtype != torch.float16:
        raise _lib.RlcfError("gemm_grouped: a and b must be fp16")
    G, m, k = a.shape
    n = b.shape[1]
    if b.shape[0] != G or b.shape[2] != k or tuple(out.shape) != (G, m, n):
        raise _lib.RlcfError(f"gemm_grouped: shapes {tuple(a.shape)} x {tuple(b.shape)} -> {tuple(out.shape)}")
    want = torch.float32 if epilogue in (EPI_RESID_F32, EPI_F32) else torch.float16
    if out.dtype != want:
        raise _lib.RlcfError(f"gemm_grouped: out must be {want}")
    bias_gs = 0
    if bias is not None:
        if bias.dtype != torch.float32 or bias.dim() != 2 or bias.shape != (G, n) or bias.stride(1) != 1:
            raise _lib.RlcfError("gemm_grouped: bias must be fp32 [G, N]")
        bias_gs = bias.stride(0)
    for t, dt, nm in ((resid, torch.float32, "resid"), (aux_in, torch.float16, "aux_in"), (aux_out, torch.float16, "aux_out")):
        if t is not None and (t.dtype != dt or t.dim() != 3 or t.stride() != out.stride()):
            raise _lib.RlcfError(f"gemm_grouped: {nm} must be {dt} with out's layout")
    call("rlcf_gemm_f16_grouped", ptr(a), a.stride(1), a.stride(0), ptr(b), b.stride(1), b.stride(0), G, m, n, k,
         epilogue, float(alpha), ptr(bias), bias_gs, ptr(resid), ptr(aux_in), ptr(aux_out), ptr(out), out.stride(1),
         out.stride(0), stream())
    return out


def im2col(images, view_idx, n_views, patch, k_pad, out):
    _chk(images, torch.float32, "images"); _chk(view_idx, torch.int32, "view_idx"); _chk(out, torch.float16, "out")
    _, c, h, w = images.shape
    call("rlcf_im2col_f16", ptr(images), ptr(view_idx), n_views, c, h, w, patch, k_pad, ptr(out), stream())
    return out


def embed_lnpre(patch_out, cls, pos, gamma, beta, param_stride, rows_per_set, n_views, L, d, x, x_pre=None, eps=1e-5,
                embed_stride=0):
    """embed_stride != 0: set g reads its own class / positional embedding at cls + g*embed_stride, pos + g*embed_stride."""
    _chk(patch_out, torch.float32, "patch_out"); _chk(x, torch.float32, "x"); _chk(x_pre, torch.float32, "x_pre")
    if embed_stride:
        call("rlcf_embed_lnpre_sets", ptr(patch_out), ptr(cls), ptr(pos), embed_stride, ptr(gamma), ptr(beta),
             param_stride, rows_per_set, n_views, L, d, eps, ptr(x_pre), ptr(x), stream())
    else:
        call("rlcf_embed_lnpre", ptr(patch_out), ptr(cls), ptr(pos), ptr(gamma), ptr(beta), param_stride, rows_per_set,
             n_views, L, d, eps, ptr(x_pre), ptr(x), stream())
    return x


def embed_text(tokens, tok_emb, pos, x):
    _chk(tokens, torch.int64, "tokens"); _chk(tok_emb, torch.float32, "tok_emb"); _chk(x, torch.float32, "x")
    n, L = tokens.shape
    call("rlcf_embed_text", ptr(tokens), ptr(tok_emb), ptr(pos), n, L, tok_emb.shape[1], ptr(x), stream())
    return x


def layernorm_fwd(x, gamma, beta, M, d, out16=None, out32=None, ldx=None, param_stride=0, rows_per_set=None,
                  eps=1e-5):
    _chk(x, torch.float32, "x"); _chk(out16, torch.float16, "out16"); _chk(out32, torch.float32, "out32")
    call("rlcf_layernorm_fwd", ptr(x), d if ldx is None else ldx, ptr(gamma), ptr(beta), param_stride,
         M if rows_per_set is None else rows_per_set, M, d, eps, ptr(out16), ptr(out32), stream())


def layernorm_bwd(dy, x, gamma, rows_per_set, n_sets, d, partials, n_slots, p_total, p_off, dx=None, accumulate=True,
                  param_stride=0, lddy=None, ldx=None, lddx=None, eps=1e-5, dx16=None):
    _chk(x, torch.float32, "x"); _chk(partials, torch.float32, "partials"); _chk(dx, torch.float32, "dx")
    _chk(dx16, torch.float16, "dx16")
    if partials is not None and partials.numel() < n_sets * n_slots * p_total:
        raise _lib.RlcfError(f"layernorm_bwd: partials holds {partials.numel()} floats, {n_sets}x{n_slots}x{p_total} needed")
    is32 = dy.dtype == torch.float32
    call("rlcf_layernorm_bwd", ptr(dy), int(is32), d if lddy is None else lddy, ptr(x), d if ldx is None else ldx,
         ptr(gamma), param_stride, rows_per_set, n_sets, d, eps, ptr(dx), d if lddx is None else lddx,
         int(accumulate), ptr(dx16), ptr(partials), n_slots, p_total, p_off, stream())


def attention_fwd(qkv, n_seq, L, heads, out, causal=False, lse=None):
    _chk(qkv, torch.float16, "qkv"); _chk(out, torch.float16, "out"); _chk(lse, torch.float32, "lse")
    call("rlcf_attention_fwd", ptr(qkv), n_seq, L, heads, int(causal), ptr(out), ptr(lse), stream())
    return out


def attention_row_fwd(qkv, n_seq, L, heads, out, q_row=0, x=None, x_row=None, q_rows=None):
    """Attention output of row q_row of every sequence ([n_seq, d]); optionally gathers that row of x into x_row.
    q_rows [n_seq, d]: that row's queries, projected by the caller; qkv is then k | v only ([n_seq*L, 2d])."""
    _chk(qkv, torch.float16, "qkv"); _chk(out, torch.float16, "out"); _chk(q_rows, torch.float16, "q_rows")
    _chk(x, torch.float32, "x"); _chk(x_row, torch.float32, "x_row")
    call("rlcf_attention_row_fwd", ptr(qkv), ptr(q_rows), n_seq, L, heads, q_row, ptr(out), ptr(x), ptr(x_row), stream())
    return out


def attention_bwd(qkv, out, dout, lse, n_seq, L, heads, dqkv, causal=False):
    _chk(qkv, torch.float16, "qkv"); _chk(out, torch.float16, "out"); _chk(dout, torch.float16, "dout")
    _chk(lse, torch.float32, "lse"); _chk(dqkv, torch.float16, "dqkv")
    call("rlcf_attention_bwd", ptr(qkv), ptr(out), ptr(dout), ptr(lse), n_seq, L, heads, int(causal), ptr(dqkv),
         stream())
    return dqkv


def head_fwd(x, gamma, beta, proj, n, d, E, feat=None, inv_norm=None, logits=None, class_feat=None, logit_scale=1.0,
             row_idx=None, row_stride=1, param_stride=0, seqs_per_set=None, eps=1e-5, proj_stride=0):
    _chk(x, torch.float32, "x"); _chk(proj, torch.float32, "proj", strided=bool(proj_stride))
    _chk(class_feat, torch.float32, "class_feat")
    _chk(row_idx, torch.int32, "row_idx"); _chk(feat, torch.float32, "feat"); _chk(logits, torch.float32, "logits")
    C = 0 if class_feat is None else class_feat.shape[0]
    sps = n if seqs_per_set is None else seqs_per_set
    if proj_stride:
        call("rlcf_head_fwd_sets", ptr(x), ptr(row_idx), row_stride, ptr(gamma), ptr(beta), param_stride, sps,
             ptr(proj), proj_stride, ptr(class_feat), float(logit_scale), n, d, E, C, eps, ptr(feat), ptr(inv_norm),
             ptr(logits), stream())
    else:
        call("rlcf_head_fwd", ptr(x), ptr(row_idx), row_stride, ptr(gamma), ptr(beta), param_stride, sps, ptr(proj),
             ptr(class_feat), float(logit_scale), n, d, E, C, eps, ptr(feat), ptr(inv_norm), ptr(logits), stream())


def entropy_select(logits, n_img, V, C, S, sel, sel_global=None, entropy=None):
    _chk(logits, torch.float32, "logits"); _chk(sel, torch.int32, "sel"); _chk(sel_global, torch.int32, "sel_global")
    call("rlcf_entropy_select", ptr(logits), n_img, V, C, S, ptr(sel), ptr(sel_global), ptr(entropy), stream())


def reward_loss(logits, row_idx, reward_img, reward_cls, n_img, S, K, C, dlogits, clipscore_weight=2.5,
                reward_process=True, process_batch=False, amplify=False, loss_scale=1.0, topk_idx=None, scores=None,
                rewards=None, loss=None):
    _chk(logits, torch.float32, "logits"); _chk(row_idx, torch.int32, "row_idx")
    _chk(reward_img, torch.float32, "reward_img"); _chk(reward_cls, torch.float32, "reward_cls")
    _chk(dlogits, torch.float32, "dlogits"); _chk(topk_idx, torch.int32, "topk_idx")
    call("rlcf_reward_loss", ptr(logits), ptr(row_idx), ptr(reward_img), ptr(reward_cls), n_img, S, K, C,
         reward_cls.shape[1], float(clipscore_weight), int(bool(reward_process)), int(bool(process_batch)),
         int(bool(amplify)), float(loss_scale), ptr(dlogits), ptr(topk_idx), ptr(scores), ptr(rewards), ptr(loss),
         stream())


def reward_loss_multi(logits, row_idx, reward_imgs, reward_clss, weights, n_img, S, K, C, dlogits, clipscore_weight=2.5,
                      reward_process=True, process_batch=False, amplify=False, loss_scale=1.0, topk_idx=None,
                      scores=None, rewards=None, loss=None):
    """reward_loss with an ensemble of reward models: lists of image features [n, E_i], class features [C, E_i] and
    per-model weights (CLIPRewardsMultiple, clip_reward.py:180-307)."""
    n = len(reward_imgs)
    if not 1 <= n <= 4 or len(reward_clss) != n or len(weights) != n:
        raise _lib.RlcfError("reward_loss_multi: 1..4 reward models with one weight each")
    _chk(logits, torch.float32, "logits"); _chk(row_idx, torch.int32, "row_idx"); _chk(dlogits, torch.float32, "dlogits")
    _chk(topk_idx, torch.int32, "topk_idx")
    for a, b in zip(reward_imgs, reward_clss):
        _chk(a, torch.float32, "reward_img"); _chk(b, torch.float32, "reward_cls")
        if a.shape[1] != b.shape[1] or b.shape[0] != C:
            raise _lib.RlcfError("reward_loss_multi: feature shapes do not match")
    pad = [None] * (4 - n)
    imgs, clss = list(reward_imgs) + pad, list(reward_clss) + pad
    ers = [t.shape[1] for t in reward_imgs] + [0] * (4 - n)
    wts = [float(x) for x in weights] + [0.0] * (4 - n)
    call("rlcf_reward_loss_multi", ptr(logits), ptr(row_idx), n, *[ptr(t) for t in imgs], *[ptr(t) for t in clss], *ers,
         *wts, n_img, S, K, C, float(clipscore_weight), int(bool(reward_process)), int(bool(process_batch)),
         int(bool(amplify)), float(loss_scale), ptr(dlogits), ptr(topk_idx), ptr(scores), ptr(rewards), ptr(loss), stream())


def avg_entropy_loss(logits, row_idx, n_img, S, C, dlogits, loss=None, loss_scale=1.0):
    _chk(logits, torch.float32, "logits"); _chk(dlogits, torch.float32, "dlogits")
    call("rlcf_avg_entropy_loss", ptr(logits), ptr(row_idx), n_img, S, C, float(loss_scale), ptr(dlogits), ptr(loss),
         stream())


def avg_entropy_reg(logits, row_idx, n_img, S, C, dlogits, weight, loss=None, loss_scale=1.0):
    """dlogits += weight * loss_scale * d avg_entropy/d logits; loss += weight * avg_entropy  (tpt_cls_rl.py:73-74)."""
    _chk(logits, torch.float32, "logits"); _chk(dlogits, torch.float32, "dlogits")
    call("rlcf_avg_entropy_reg", ptr(logits), ptr(row_idx), n_img, S, C, float(loss_scale), float(weight), ptr(dlogits),
         ptr(loss), stream())


def head_bwd(dlogits, x, gamma, proj, class_feat, logit_scale, feat, inv_norm, n_img, S, d, E, C, dres, partials,
             n_slots, p_total, p_off, row_idx=None, row_stride=1, param_stride=0, eps=1e-5):
    _chk(dlogits, torch.float32, "dlogits"); _chk(x, torch.float32, "x"); _chk(dres, torch.float32, "dres")
    call("rlcf_head_bwd", ptr(dlogits), ptr(x), ptr(row_idx), row_stride, ptr(gamma), param_stride, ptr(proj),
         ptr(class_feat), float(logit_scale), ptr(feat), ptr(inv_norm), n_img, S, d, E, C, eps, ptr(dres),
         ptr(partials), n_slots, p_total, p_off, stream())


def head_bwd_ex(dlogits, dl_strides, x, gamma, proj, other_feat, other_set_stride, logit_scale, feat, inv_norm, n_sets,
                seqs_per_set, d, E, K, dres, row_idx=None, row_stride=1, param_stride=0, partials=None, n_slots=1,
                p_total=0, p_off=0, eps=1e-5, beta=None, y_out=None, df_out=None, proj_stride=0):
    _chk(dlogits, torch.float32, "dlogits"); _chk(x, torch.float32, "x"); _chk(dres, torch.float32, "dres")
    _chk(other_feat, torch.float32, "other_feat"); _chk(row_idx, torch.int32, "row_idx")
    if partials is not None and partials.numel() < n_sets * n_slots * p_total:
        raise _lib.RlcfError(f"head_bwd: partials holds {partials.numel()} floats, {n_sets}x{n_slots}x{p_total} needed")
    if proj_stride:
        call("rlcf_head_bwd_sets", ptr(dlogits), dl_strides[0], dl_strides[1], dl_strides[2], ptr(x), ptr(row_idx),
             row_stride, ptr(gamma), param_stride, ptr(proj), proj_stride, ptr(other_feat), other_set_stride,
             float(logit_scale), ptr(feat), ptr(inv_norm), n_sets, seqs_per_set, d, E, K, eps, ptr(dres), ptr(partials),
             n_slots, p_total, p_off, ptr(beta), ptr(y_out), ptr(df_out), stream())
        return
    call("rlcf_head_bwd_ex", ptr(dlogits), dl_strides[0], dl_strides[1], dl_strides[2], ptr(x), ptr(row_idx), row_stride,
         ptr(gamma), param_stride, ptr(proj), ptr(other_feat), other_set_stride, float(logit_scale), ptr(feat),
         ptr(inv_norm), n_sets, seqs_per_set, d, E, K, eps, ptr(dres), ptr(partials), n_slots, p_total, p_off,
         ptr(beta), ptr(y_out), ptr(df_out), stream())


def adamw_step_from(params, m, v, grads, n_sets, n_slots, p_total, lr, step, params_in, params_in_stride, fresh,
                    beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, loss_scale=1.0):
    for t, nm in ((params, "params"), (m, "m"), (v, "v"), (grads, "grads"), (params_in, "params_in")):
        _chk(t, torch.float32, nm)
    call("rlcf_adamw_step_from", ptr(params), ptr(m), ptr(v), ptr(grads), n_sets, n_slots, p_total, float(lr),
         float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(loss_scale), ptr(params_in),
         params_in_stride, int(bool(fresh)), stream())


def transpose_blocks(src, n_sets, rows_per_set, rows_pad, cols, out, ld_out, skip_first=0, in_set_stride_rows=None):
    """out[c, g*rows_pad + r] = src[row(g, r), c] as fp16 (src fp16 or fp32), zero padded rows."""
    _chk(out, torch.float16, "out")
    if src.dtype not in (torch.float16, torch.float32) or not src.is_cuda:
        raise _lib.RlcfError("transpose_blocks: src must be a CUDA fp16/fp32 tensor")
    stride = rows_per_set if in_set_stride_rows is None else in_set_stride_rows
    call("rlcf_transpose_blocks_f16", ptr(src), int(src.dtype == torch.float32), n_sets, rows_per_set, rows_pad, cols,
         skip_first, stride, ptr(out), ld_out, stream())
    return out


def transpose_blocks_colsum(src, n_sets, rows_per_set, rows_pad, cols, out, ld_out, colsum=None, colsum_stride=0,
                            skip_first=0, in_set_stride_rows=None):
    """transpose_blocks for fp16 input; colsum (fp32 base pointer, set stride colsum_stride) also receives the per-set
    column sums."""
    _chk(src, torch.float16, "src"); _chk(out, torch.float16, "out"); _chk(colsum, torch.float32, "colsum", strided=True)
    stride = rows_per_set if in_set_stride_rows is None else in_set_stride_rows
    call("rlcf_transpose_blocks_colsum", ptr(src), n_sets, rows_per_set, rows_pad, cols, skip_first, stride, ptr(out),
         ld_out, ptr(colsum), colsum_stride, stream())
    return out


def colsum_f16(src, n_sets, rows_per_set, cols, out, out_stride):
    _chk(src, torch.float16, "src"); _chk(out, torch.float32, "out", strided=True)
    call("rlcf_colsum_f16", ptr(src), n_sets, rows_per_set, cols, ptr(out), out_stride, stream())


def seq_sum(dx, n_sets, S, L, d, out, out_stride):
    _chk(dx, torch.float32, "dx"); _chk(out, torch.float32, "out", strided=True)
    call("rlcf_seq_sum", ptr(dx), n_sets, S, L, d, ptr(out), out_stride, stream())


def outer_sum(y, df, n_sets, S, d, E, out, out_stride):
    _chk(y, torch.float32, "y"); _chk(df, torch.float32, "df"); _chk(out, torch.float32, "out", strided=True)
    call("rlcf_outer_sum", ptr(y), ptr(df), n_sets, S, d, E, ptr(out), out_stride, stream())


def embed_prompts(tokens, tok_emb, pos, ctx, ctx_stride, n_ctx, n_sets, x):
    _chk(tokens, torch.int64, "tokens"); _chk(tok_emb, torch.float32, "tok_emb"); _chk(ctx, torch.float32, "ctx")
    _chk(x, torch.float32, "x")
    n_cls, L = tokens.shape
    call("rlcf_embed_prompts", ptr(tokens), ptr(tok_emb), ptr(pos), ptr(ctx), ctx_stride, n_ctx, n_sets, n_cls, L,
         tok_emb.shape[1], ptr(x), stream())
    return x


def embed_prompts_map(tokens, tok_emb, pos, vec, vec_stride, src_map, n_sets, x):
    """x[g, c, t] = (src_map[c, t] >= 0 ? tok_emb[tokens[c, src_map[c, t]]] : vec[g][-1 - src_map[c, t]]) + pos[t]."""
    _chk(tokens, torch.int64, "tokens"); _chk(tok_emb, torch.float32, "tok_emb"); _chk(vec, torch.float32, "vec")
    _chk(src_map, torch.int32, "src_map"); _chk(x, torch.float32, "x")
    n_cls, L = tokens.shape
    if tuple(src_map.shape) != (n_cls, L):
        raise _lib.RlcfError("embed_prompts_map: src_map must be [n_cls, L]")
    call("rlcf_embed_prompts_map", ptr(tokens), ptr(tok_emb), ptr(pos), ptr(vec), vec_stride, ptr(src_map), n_sets, n_cls,
         L, tok_emb.shape[1], ptr(x), stream())
    return x


def vec_grad_map(dx, ctx_pos, cls_pos, n_sets, n_cls, L, n_ctx, d, dvec):
    _chk(dx, torch.float32, "dx"); _chk(ctx_pos, torch.int32, "ctx_pos"); _chk(cls_pos, torch.int32, "cls_pos")
    _chk(dvec, torch.float32, "dvec")
    call("rlcf_vec_grad_map", ptr(dx), ptr(ctx_pos), ptr(cls_pos), n_sets, n_cls, L, n_ctx, d, ptr(dvec), stream())
    return dvec


def pair_logits(img_feat, txt_feat, txt_set_stride, n_sets, S, C, E, logit_scale, logits):
    _chk(img_feat, torch.float32, "img_feat"); _chk(txt_feat, torch.float32, "txt_feat")
    _chk(logits, torch.float32, "logits")
    call("rlcf_pair_logits", ptr(img_feat), ptr(txt_feat), txt_set_stride, n_sets, S, C, E, float(logit_scale),
         ptr(logits), stream())
    return logits


def ctx_grad(dx, n_sets, n_cls, L, n_ctx, d, dctx):
    _chk(dx, torch.float32, "dx"); _chk(dctx, torch.float32, "dctx")
    call("rlcf_ctx_grad", ptr(dx), n_sets, n_cls, L, n_ctx, d, ptr(dctx), stream())
    return dctx


def adamw_step(params, m, v, partials, n_sets, n_slots, p_total, lr, step, beta1=0.9, beta2=0.999, eps=1e-8,
               weight_decay=1e-2, loss_scale=1.0, grad_out=None):
    for t, nm in ((params, "params"), (m, "m"), (v, "v"), (partials, "partials"), (grad_out, "grad_out")):
        _chk(t, torch.float32, nm)
    if partials.numel() < n_sets * n_slots * p_total or params.numel() < n_sets * p_total:
        raise _lib.RlcfError(f"adamw_step: buffers too small for {n_sets} sets x {n_slots} slots x {p_total} parameters")
    call("rlcf_adamw_step", ptr(params), ptr(m), ptr(v), ptr(partials), n_sets, n_slots, p_total, float(lr),
         float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(loss_scale), ptr(grad_out),
         stream())


def reset_params(init, params, m, v, n_sets, p_total):
    _chk(init, torch.float32, "init"); _chk(params, torch.float32, "params")
    call("rlcf_reset_params", ptr(init), ptr(params), ptr(m), ptr(v), n_sets, p_total, stream())


def gather_seqs(src, idx, dst, n, rows_per_seq, dst_seq0=0):
    """dst[l, (dst_seq0 + j) * rows_per_seq + r] = src[l, idx[j] * rows_per_seq + r] for j < n: src / dst are [layers, rows,
    width] (or [rows, width]) tensors of the same dtype and width; idx int32 sequence numbers."""
    _chk(idx, torch.int32, "idx")
    if src.dtype != dst.dtype or src.shape[-1] != dst.shape[-1] or not src.is_contiguous() or not dst.is_contiguous():
        raise _lib.RlcfError("gather_seqs: src and dst must be contiguous tensors of one dtype and width")
    layers = src.shape[0] if src.dim() == 3 else 1
    if dst.dim() != src.dim() or (src.dim() == 3 and dst.shape[0] != layers):
        raise _lib.RlcfError("gather_seqs: src and dst must have the same number of layers")
    row_bytes = src.shape[-1] * src.element_size()
    seq_bytes = rows_per_seq * row_bytes
    if (dst_seq0 + n) * rows_per_seq > dst.shape[-2]:
        raise _lib.RlcfError("gather_seqs: destination too small")
    call("rlcf_gather_seqs", ptr(src), ptr(idx), dst.data_ptr() + dst_seq0 * seq_bytes, seq_bytes,
         src.shape[-2] * row_bytes, dst.shape[-2] * row_bytes, layers, n, stream())
    return dst


def cast_f16(src, k_pad=None, out=None, rows=None):
    """fp32 [rows, cols] -> fp16 [rows, k_pad] (zero padded)."""
    _chk(src, torch.float32, "src"); _chk(out, torch.float16, "out")
    cols = src.shape[1]
    rows = src.shape[0] if rows is None else rows
    ld = cols if k_pad is None else k_pad
    if out is None:
        out = torch.empty(rows, ld, dtype=torch.float16, device=src.device)
    call("rlcf_cast_f16", ptr(src), rows, cols, cols, ptr(out), ld, stream())
    return out


def transpose_cast_f16(src):
    """fp32 [rows, cols] -> fp16 [cols, rows]."""
    _chk(src, torch.float32, "src")
    rows, cols = src.shape
    out = torch.empty(cols, rows, dtype=torch.float16, device=src.device)
    call("rlcf_transpose_cast_f16", ptr(src), rows, cols, ptr(out), stream())
    return out


def transpose_cast_f16_sets(src, rows, cols, n_sets, in_stride, out, out_stride):
    """n_sets transposing casts in one launch: out[g] ([cols, rows] fp16 at g*out_stride) = src[g] ([rows, cols] fp32 at
    g*in_stride)^T.  src / out are base pointers (strided views)."""
    _chk(src, torch.float32, "src", strided=True); _chk(out, torch.float16, "out", strided=True)
    call("rlcf_transpose_cast_f16_sets", ptr(src), rows, cols, n_sets, in_stride, ptr(out), out_stride, stream())
    return out


def retrieval_loss(logits, reward_query, reward_gallery, K, dlogits, clipscore_weight=2.5, reward_process=True,
                   amplify=False, loss_scale=1.0, topk_idx=None, scores=None, rewards=None, loss=None):
    """One retrieval query per row of logits [Q, C]: top-K sampling, CLIPScore, rewards, loss and dlogits
    (retrieval/clip_ret_policy.py:88-98 / 121-131)."""
    _chk(logits, torch.float32, "logits"); _chk(reward_query, torch.float32, "reward_query")
    _chk(reward_gallery, torch.float32, "reward_gallery"); _chk(dlogits, torch.float32, "dlogits")
    _chk(topk_idx, torch.int32, "topk_idx"); _chk(scores, torch.float32, "scores")
    _chk(rewards, torch.float32, "rewards"); _chk(loss, torch.float32, "loss")
    Q, C = logits.shape
    if reward_gallery.shape[0] != C or reward_query.shape[0] != Q or reward_query.shape[1] != reward_gallery.shape[1]:
        raise _lib.RlcfError("retrieval_loss: reward feature shapes do not match the score rows")
    call("rlcf_retrieval_loss", ptr(logits), C, ptr(reward_query), ptr(reward_gallery), Q, K, C,
         reward_gallery.shape[1], float(clipscore_weight), int(bool(reward_process)), int(bool(amplify)),
         float(loss_scale), ptr(dlogits), ptr(topk_idx), ptr(scores), ptr(rewards), ptr(loss), stream())


def dfeat_partial(dlogits, gallery, partial):
    """partial [Q, n_chunks, E] = per-chunk pieces of dlogits [Q, C] @ gallery [C, E]."""
    _chk(dlogits, torch.float32, "dlogits"); _chk(gallery, torch.float32, "gallery"); _chk(partial, torch.float32, "partial")
    Q, C = dlogits.shape
    call("rlcf_dfeat_partial", ptr(dlogits), ptr(gallery), Q, C, gallery.shape[1], partial.shape[1], ptr(partial),
         stream())
    return partial


def rowdot(a, b, out, out_stride=1, scale=1.0):
    _chk(a, torch.float32, "a"); _chk(b, torch.float32, "b"); _chk(out, torch.float32, "out", strided=True)
    call("rlcf_rowdot", ptr(a), ptr(b), a.shape[0], a.shape[1], float(scale), ptr(out), out_stride, stream())
    return out


def gemm_f32(a, w, out, M=None, epilogue=0, bias=None, resid=None):
    """out[:M] = epi(a[:M] @ w^T + bias) in fp32 on the CUDA cores (epilogue 0 none, 1 QuickGELU, 2 + resid)."""
    for t, nm in ((a, "a"), (w, "w"), (out, "out")):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise _lib.RlcfError(f"gemm_f32: {nm} must be a 2-D CUDA fp32 tensor with unit inner stride")
    _chk(bias, torch.float32, "bias"); _chk(resid, torch.float32, "resid")
    m = a.shape[0] if M is None else M
    n, k = w.shape
    if a.shape[1] != k or out.shape[1] != n or out.shape[0] < m:
        raise _lib.RlcfError(f"gemm_f32: shapes {tuple(a.shape)} x {tuple(w.shape)} -> {tuple(out.shape)}")
    if resid is not None and (resid.stride(0) != out.stride(0) or resid.shape[1] != n):
        raise _lib.RlcfError("gemm_f32: resid must share out's layout")
    call("rlcf_gemm_f32", ptr(a), a.stride(0), ptr(w), w.stride(0), m, n, k, int(epilogue), ptr(bias), ptr(resid),
         ptr(out), out.stride(0), stream())
    return out


def attention_f32(qkv, n_seq, L, heads, out, causal=False):
    _chk(qkv, torch.float32, "qkv"); _chk(out, torch.float32, "out")
    call("rlcf_attention_f32", ptr(qkv), n_seq, L, heads, int(bool(causal)), ptr(out), stream())
    return out


def accuracy_count(logits, target, hits):
    """hits[0] += top-1 hits, hits[1] += top-5 hits, hits[2] += rows  (utils/tools.py:84-98; device int64 [3])."""
    _chk(logits, torch.float32, "logits"); _chk(target, torch.int64, "target"); _chk(hits, torch.int64, "hits")
    if logits.dim() != 2 or target.numel() != logits.shape[0] or hits.numel() != 3:
        raise _lib.RlcfError("accuracy_count: logits [n, C], target [n], hits [3]")
    call("rlcf_accuracy_count", ptr(logits), ptr(target), logits.shape[0], logits.shape[1], ptr(hits), stream())
    return hits


def add_rows(a, a_stride, b, b_stride, n_sets, n, x):
    """x[g, :n] = a[g*a_stride : +n] + b[g*b_stride : +n]  (a, b base pointers of strided per-query vectors)."""
    _chk(a, torch.float32, "a", strided=True); _chk(b, torch.float32, "b", strided=True); _chk(x, torch.float32, "x")
    call("rlcf_add_rows", ptr(a), a_stride, ptr(b), b_stride, n_sets, n, ptr(x), stream())
    return x


def scale_rows_exp(src, ls, ls_stride, out):
    """out[q, :] = src[q, :] * exp(ls[q*ls_stride])."""
    _chk(src, torch.float32, "src"); _chk(ls, torch.float32, "ls", strided=True); _chk(out, torch.float32, "out")
    call("rlcf_scale_rows_exp", ptr(src), ptr(ls), ls_stride, src.shape[0], src.shape[1], ptr(out), stream())
    return out


def tied_rows_grad(dx, tokens, n_sets, L, d, g_tok, g_pos, out_stride):
    _chk(dx, torch.float32, "dx"); _chk(tokens, torch.int64, "tokens")
    _chk(g_tok, torch.float32, "g_tok", strided=True); _chk(g_pos, torch.float32, "g_pos", strided=True)
    call("rlcf_tied_rows_grad", ptr(dx), ptr(tokens), n_sets, L, d, ptr(g_tok), ptr(g_pos), out_stride, stream())


def adamw_full(params, m, v, grads, n_sets, p_total, lr, step, params_in, params_in_stride, fresh, w16=None, n16=0,
               beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, loss_scale=1.0, set_stride=0):
    """Streaming AdamW over whole per-sample parameter vectors [n_sets, p_total]; also writes the fp16 copy of the first
    n16 parameters of every sample into w16 [n_sets, >= n16].  set_stride > p_total: params / m / v / grads are base
    pointers into larger per-sample vectors and only p_total entries of each are updated."""
    for t, nm in ((params, "params"), (m, "m"), (v, "v"), (grads, "grads")):
        _chk(t, torch.float32, nm, strided=bool(set_stride))
    _chk(params_in, torch.float32, "params_in", strided=True); _chk(w16, torch.float16, "w16")
    call("rlcf_adamw_full", ptr(params), ptr(m), ptr(v), ptr(grads), n_sets, p_total, float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(step), float(loss_scale), ptr(params_in), params_in_stride,
         int(bool(fresh)), ptr(w16), 0 if w16 is None else w16.stride(0), n16, set_stride, stream())


def transpose_f16_sets(src, rows, cols, out, n_sets, set_stride):
    _chk(src, torch.float16, "src", strided=True); _chk(out, torch.float16, "out", strided=True)
    call("rlcf_transpose_f16_sets", ptr(src), rows, cols, ptr(out), n_sets, set_stride, stream())


def resample_u8(src, hdr, hb, hk, vb, vk, out_h, out_w, tmp, out):
    """PIL-exact crops + resizes of one uint8 HWC image (see rlcf_resample_u8); all plan tensors int32 on the device."""
    _chk(src, torch.uint8, "src"); _chk(tmp, torch.uint8, "tmp"); _chk(out, torch.uint8, "out")
    for t, nm in ((hdr, "hdr"), (hb, "hb"), (hk, "hk"), (vb, "vb"), (vk, "vk")):
        _chk(t, torch.int32, nm)
    H, W, _ = src.shape
    V = hdr.shape[0]
    if tuple(out.shape) != (V, out_h, out_w, 3) or tmp.shape[0] < V or tuple(tmp.shape[2:]) != (out_w, 3):
        raise _lib.RlcfError("resample_u8: out must be [V, out_h, out_w, 3] and tmp [V, rows, out_w, 3]")
    call("rlcf_resample_u8", ptr(src), H, W, V, ptr(hdr), ptr(hb), ptr(hk), hk.shape[-1], ptr(vb), ptr(vk), vk.shape[-1],
         out_h, out_w, ptr(tmp), tmp.shape[1], ptr(out), stream())
    return out


def augmix_views(x_orig, vflag, wts, omm, n_ops, op_codes, mats, mean, std, out):
    _chk(x_orig, torch.uint8, "x_orig"); _chk(vflag, torch.int32, "vflag"); _chk(wts, torch.float32, "wts")
    _chk(omm, torch.float32, "omm"); _chk(n_ops, torch.int32, "n_ops"); _chk(op_codes, torch.int32, "ops")
    _chk(mats, torch.float64, "mats"); _chk(out, torch.float32, "out")
    V = x_orig.shape[0]
    if tuple(x_orig.shape[1:]) != (224, 224, 3) or tuple(out.shape) != (V, 3, 224, 224):
        raise _lib.RlcfError("augmix_views: x_orig must be [V,224,224,3] uint8 and out [V,3,224,224] fp32")
    call("rlcf_augmix_views", ptr(x_orig), V, ptr(vflag), ptr(wts), ptr(omm), ptr(n_ops), ptr(op_codes), ptr(mats),
         float(mean[0]), float(mean[1]), float(mean[2]), float(std[0]), float(std[1]), float(std[2]), ptr(out), stream())
    return out


def resample_taps(geom, out, ks_h, ks_v, hdr):
    """Pillow's tap tables on the device (see rlcf_resample_taps): returns (hb, hk, vb, vk) and fills hdr[:, 2:4]."""
    _chk(geom, torch.int32, "geom"); _chk(hdr, torch.int32, "hdr")
    V, dev = geom.shape[0], geom.device
    hb = torch.empty(V, out, 2, dtype=torch.int32, device=dev); vb = torch.empty_like(hb)
    hk = torch.empty(V, out, ks_h, dtype=torch.int32, device=dev)
    vk = torch.empty(V, out, ks_v, dtype=torch.int32, device=dev)
    call("rlcf_resample_taps", ptr(geom), V, out, ks_h, ks_v, ptr(hdr), ptr(hb), ptr(hk), ptr(vb), ptr(vk), stream())
    return hb, hk, vb, vk


def gemm_wgrad_adamw(a, b, params, m, v, params_in, params_in_gs, fresh, w16, lr, step, beta1=0.9, beta2=0.999, eps=1e-8,
                     weight_decay=1e-2, loss_scale=1.0):
    """Fused weight gradient + AdamW (rlcf_gemm_wgrad_adamw): a [G, n_out, K], b [G, n_in, K] fp16 views; params / m / v
    fp32 [G, n_out, n_in] views of one layout (params is written, params_in -- a base pointer, group stride params_in_gs
    floats -- is read); w16 fp16 [G, n_out, n_in] view receiving the updated weights (or None)."""
    for t, nm in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.float16 or t.dim() != 3 or t.stride(2) != 1:
            raise _lib.RlcfError(f"gemm_wgrad_adamw: {nm} must be a 3-D CUDA fp16 tensor with unit inner stride")
    G, n_out, k = a.shape
    n_in = b.shape[1]
    for t, nm in ((params, "params"), (m, "m"), (v, "v")):
        if t.dtype != torch.float32 or tuple(t.shape) != (G, n_out, n_in) or t.stride() != params.stride() or t.stride(2) != 1:
            raise _lib.RlcfError(f"gemm_wgrad_adamw: {nm} must be fp32 [G, n_out, n_in] with params' strides")
    if w16 is not None and (w16.dtype != torch.float16 or tuple(w16.shape) != (G, n_out, n_in)
                            or w16.stride(1) != params.stride(1) or w16.stride(2) != 1):
        raise _lib.RlcfError("gemm_wgrad_adamw: w16 must be fp16 [G, n_out, n_in] with params' row stride")
    _chk(params_in, torch.float32, "params_in", strided=True)
    call("rlcf_gemm_wgrad_adamw", ptr(a), a.stride(1), a.stride(0), ptr(b), b.stride(1), b.stride(0), G, n_out, n_in, k,
         ptr(params), ptr(m), ptr(v), params.stride(1), params.stride(0), ptr(params_in), params_in_gs, int(bool(fresh)),
         ptr(w16), 0 if w16 is None else w16.stride(0), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), int(step), float(loss_scale), stream())


def bicubic_resize(images, view_idx, n_views, out):
    """out [n_views, C, oh, ow] = interpolate(images[view_idx], size=(oh, ow), mode="bicubic", align_corners=True)."""
    _chk(images, torch.float32, "images"); _chk(view_idx, torch.int32, "view_idx"); _chk(out, torch.float32, "out")
    _, c, h, w = images.shape
    if out.shape[0] < n_views or out.shape[1] != c:
        raise _lib.RlcfError("bicubic_resize: out must be [>= n_views, C, oh, ow]")
    call("rlcf_bicubic_resize", ptr(images), ptr(view_idx), n_views, c, h, w, out.shape[2], out.shape[3], ptr(out), stream())
    return out
