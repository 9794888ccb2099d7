"""Training path of the bridge: autograd Functions over the C ABI.

The reference trains the projector with plain autograd through nn.LayerNorm / nn.Linear / nn.SiLU
(Multitask/model/projector.py:149-151) and through the index_put of the splice
(Multitask/model/ps-slm.py:833-869); the loss comes back as the gradient of ``inputs_embeds``
(Multitask/utils/deepspeed_utils.py:235).  Here:

* ``LinearSiLUFunction``  — forward: LN-folded GEMM-1 (pre-activation kept in fp32), SiLU, GEMM-2;
  backward: dW2 = dyᵀ·h, dh = dy·W2, dz = dh·silu'(z), G = (rstd·dz)ᵀ·x on the tensor cores, then
  dW1/dγ/dβ from G by the LayerNorm-fold algebra (no third big GEMM), db1, db2.  The input (a
  posterior) never needs a gradient.
* ``SpliceFunction``      — forward: tasu_splice_scatter; backward: gather of the upstream gradient at
  the audio slots (tasu_gather_rows through the scatter's audio_dest map) and, when the text embeddings require it
  (freeze_llm=False, or PEFT with trained embed_tokens, ps-slm.py:119-123), the scatter of the gradient back to the
  text tokens (tasu_splice_text_grad).
"""
import torch

from . import _lib as L
from . import ops
from .bridge import cast_weight_bf16


class LinearSiLUFunction(torch.autograd.Function):
    """y = W2·silu(W1·LN(x) + b1) + b2 on prepared operands (xb bf16 [N, pad64(V)], mean, rstd)."""

    @staticmethod
    def forward(ctx, xb, mean, rstd, n_rows, gamma, beta, w1, b1, w2, b2, out_dtype):
        V = w1.shape[1]
        Hb, H = w1.shape[0], w2.shape[0]
        dev = xb.device
        w1g, colsum, dbias = ops.fold_layernorm(w1.detach(), gamma.detach(), beta.detach(), b1.detach())
        w2b = cast_weight_bf16(w2)
        z = torch.empty(n_rows, Hb, dtype=torch.float32, device=dev)
        ops.gemm_bf16_tn(xb, w1g, n_rows, Hb, V, z, L.EPI_LNFOLD, dbias, rstd, mean, colsum)
        h = ops.silu_fwd(z)
        y = torch.empty(n_rows, H, dtype=out_dtype, device=dev)
        ops.gemm_bf16_tn(h, w2b, n_rows, H, Hb, y, L.EPI_BIAS, b2.detach().float().contiguous())
        ctx.save_for_backward(xb, mean, rstd, z, h, gamma, beta, w1, w2)
        ctx.n_rows = n_rows
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, mean, rstd, z, h, gamma, beta, w1, w2 = ctx.saved_tensors
        N = ctx.n_rows
        Hb, V = w1.shape
        H = w2.shape[0]
        dev = dy.device
        dy = dy.contiguous()
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        db2 = ops.colsum(dy)
        dyb, _, _ = ops.cast_rows(dy, torch.bfloat16) if dy.dtype != torch.bfloat16 else (dy, None, None)
        # the three backward contractions read dy, h, W2 and x where they lie (MN-major GEMM operands): no transposes
        dw2 = torch.empty(H, Hb, dtype=torch.float32, device=dev)
        ops.gemm_bf16_f32(dyb, True, h, True, H, Hb, N, dw2)      # dW2 = dyᵀ·h
        dh = torch.empty(N, Hb, dtype=torch.float32, device=dev)
        ops.gemm_bf16_f32(dyb, False, cast_weight_bf16(w2), True, N, Hb, H, dh)   # dh = dy·W2
        dzsT, db1, g0 = ops.silu_bwd(dh, z, rstd, mean)           # [Hb, N] = (rstd·dz)ᵀ
        ldv = ops.pad_to(V, 4)
        G = torch.empty(Hb, ldv, dtype=torch.float32, device=dev)
        ops.gemm_bf16_f32(dzsT, False, xb, True, Hb, V, N, G)     # G = (rstd·dz)ᵀ·x
        dw1, dgamma, dbeta = ops.linear_silu_wgrad_finish(G, w1.detach().float().contiguous(),
                                                          gamma.detach().float().contiguous(),
                                                          beta.detach().float().contiguous(), g0, db1)
        return (None, None, None, None, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), dw1.to(w1.dtype),
                db1.to(w1.dtype), dw2.to(w2.dtype), db2.to(w2.dtype), None)


class TokenRowLinearSiLUFunction(torch.autograd.Function):
    """y = W2·silu(W1·LN(x) + b1) + b2 for TEXT-SIMULATED rows x given only as descriptors
    (``ops.TokenRows``): GEMM-1 and its two backward contractions collapse to a column gather / scatter of the
    fp32 W1 (csrc/tokrows.cu); GEMM-2, dW2 and dh stay on the tensor cores.  Replaces
    ps-slm.py:360-409 + projector.py:149-151 (+ autograd) for the text-only training step."""

    @staticmethod
    def forward(ctx, rows, gamma, beta, w1, b1, w2, b2, eps, out_dtype):
        ws = ops.tokrow_train_workspace(rows, w1.shape[0], w2.shape[0], w1.device)
        y, z, h, row_a, row_e = ops.tokrow_linear_silu_fwd(rows, gamma, beta, w1, b1, w2, b2, eps, out_dtype, ws)
        ctx.save_for_backward(z, h, row_a, row_e, gamma, beta, w1, w2)
        ctx.rows, ctx.ws = rows, ws
        return y

    @staticmethod
    def backward(ctx, dy):
        z, h, row_a, row_e, gamma, beta, w1, w2 = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        from . import dist as D
        dgamma, dbeta, dw1, db1, dw2, db2 = ops.tokrow_linear_silu_bwd(dy, ctx.rows, z, h, row_a, row_e, gamma, beta,
                                                                       w1, w2, ctx.ws, between=D.overlap_hook())
        ctx.ws = None
        return (None, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), dw1.to(w1.dtype), db1.to(w1.dtype),
                dw2.to(w2.dtype), db2.to(w2.dtype), None, None)


def linear_silu_train_rows(module, xb, mean, rstd, n_rows, out_dtype=torch.float32):
    """Trainable projector forward on prepared rows (used by the packed text-only training step)."""
    return LinearSiLUFunction.apply(xb, mean, rstd, n_rows, module.norm.weight, module.norm.bias,
                                    module.ffn[0].weight, module.ffn[0].bias, module.ffn[2].weight,
                                    module.ffn[2].bias, out_dtype)


def linear_silu_train(module, x):
    """Trainable ``EncoderProjectorLinearSiLU.forward`` for a dense [B, T, V] input."""
    if x.requires_grad:
        raise NotImplementedError("the bridge projector does not propagate a gradient to its (posterior) input")
    B, T, D = x.shape
    xb, mean, rstd = ops.cast_rows(x.reshape(B * T, D), torch.bfloat16, ops.pad_to(D), want_ln=True,
                                   ln_eps=module.norm.eps)
    out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
    y = linear_silu_train_rows(module, xb, mean, rstd, B * T, out_dtype)
    return y.view(B, T, -1)


class LinearFunction(torch.autograd.Function):
    """y = act(x·Wᵀ + b) on the tensor cores with a full backward (act ∈ {none, relu}); used by the
    ``linear`` / ``simple_linear`` projectors (projector.py:39-50, :18-26) when they are trained."""

    @staticmethod
    def forward(ctx, x2, w, b, relu, out_dtype):
        N, K = x2.shape
        O = w.shape[0]
        xb, _, _ = ops.cast_rows(x2.detach(), torch.bfloat16, ops.pad_to(K))
        wb = cast_weight_bf16(w)
        y = torch.empty(N, ops.pad_to(O, 8), dtype=out_dtype, device=x2.device)
        ops.gemm_bf16_tn(xb, wb, N, O, K, y, L.EPI_BIAS_RELU if relu else L.EPI_BIAS, b.detach().float().contiguous())
        y = y[:, :O]
        ctx.save_for_backward(xb, w, y if relu else None)
        ctx.relu, ctx.K, ctx.need_dx = relu, K, x2.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w, y = ctx.saved_tensors
        N, K = xb.shape[0], ctx.K
        O = w.shape[0]
        dy = dy.float()
        if ctx.relu:
            dy = dy * (y > 0)                               # elementwise mask (plumbing), the contractions are ours
        dy = dy.contiguous()
        db = ops.colsum(dy)
        dyT = ops.transpose_cast(dy, N, O)                  # [O, N]
        xT = ops.transpose_cast(xb, N, K)                   # [K, N]
        dw = torch.empty(O, ops.pad_to(K, 4), dtype=torch.float32, device=dy.device)
        ops.gemm_bf16_tn(dyT, xT, O, K, N, dw)              # dW = dyᵀ·x
        dx = None
        if ctx.need_dx:
            dyb, _, _ = ops.cast_rows(dy, torch.bfloat16, ops.pad_to(O, 8))
            wT = ops.transpose_cast(w.detach(), O, K)       # [K, O]
            dx = torch.empty(N, ops.pad_to(K, 4), dtype=torch.float32, device=dy.device)
            ops.gemm_bf16_tn(dyb, wT, N, K, O, dx)          # dx = dy·W
            dx = dx[:, :K]
        return dx, dw[:, :K].to(w.dtype), db.to(w.dtype), None, None


class SpliceFunction(torch.autograd.Function):
    """Differentiable splice (ps-slm.py:833-869): the gradient of ``inputs_embeds`` flows to the audio rows (gather at the
    audio slots) and — when ``text_src`` is the ``[B, S, H]`` text embedding tensor (text_mode 0) and requires it — to the
    text embeddings (every text token receives the gradient of the row it was copied to).  With the fused embedding
    lookup (text_mode 1) the table is a constant: callers that train ``embed_tokens`` materialise ``inputs_embeds`` first
    (bridge.merge_packed_audio_rows does)."""

    @staticmethod
    def forward(ctx, audio_rows, text_src, plan, spliced_len, text_mode, audio_layout, audio_max_len, labels,
                pad_id, ignore_id):
        emb, mask, out_labels, pos, fids = ops.splice_scatter(plan, spliced_len, text_src.detach(), text_mode,
                                                              audio_rows.detach(), audio_layout, audio_max_len,
                                                              labels, pad_id, ignore_id,
                                                              left_padding=getattr(plan, "left_padding", None),
                                                              want_audio_dest=True)
        ctx.plan, ctx.layout, ctx.max_len, ctx.text_mode = plan, audio_layout, audio_max_len, text_mode
        ctx.shape, ctx.text_shape = tuple(audio_rows.shape), tuple(text_src.shape)
        ctx.mark_non_differentiable(mask, pos, fids)
        if out_labels is not None:
            ctx.mark_non_differentiable(out_labels)
        return emb, mask, out_labels, pos, fids

    @staticmethod
    def backward(ctx, demb, *_):
        ga = gt = None
        if ctx.needs_input_grad[0]:
            n_rows = ctx.shape[0] if ctx.layout == 0 else 0
            ga = ops.splice_audio_grad(ctx.plan, demb, ctx.layout, ctx.max_len, n_rows).view(ctx.shape)
        if ctx.needs_input_grad[1]:
            if ctx.text_mode != 0:
                raise NotImplementedError("gradient to the embedding table through the fused lookup: pass inputs_embeds")
            gt = ops.splice_text_grad(ctx.plan, demb).view(ctx.text_shape)
        return (ga, gt, None, None, None, None, None, None, None, None)
