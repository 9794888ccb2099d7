"""The ``encoder_projector`` plugin family, B200-native.

Same class names, constructor signature ``Cls(config)`` (reads ``config.encoder_dim``,
``config.llm_dim``, ``config.encoder_projector_ds_rate``), ``forward`` shape contract
``[B, T, D] → [B, T//k, llm_dim]``, attribute ``k``, parameter names/shapes and initialisers as
Multitask/model/projector.py, so reference checkpoints load with ``load_state_dict`` and
DeepSpeed/AdamW own ordinary ``nn.Parameter``s:

* ``EncoderProjectorLinearSiLU`` ("linear-silu", default) ← projector.py:129-151
* ``EncoderProjectorConcat``     ("linear")              ← projector.py:29-50
* ``EncoderProjectorLinear``     ("simple_linear")       ← projector.py:10-26

The Linear layers run as bf16 tcgen05 GEMMs with fp32 accumulation (libtasu_bridge.so); the
LayerNorm of the default projector is folded into GEMM-1's epilogue.  There is no PyTorch
fallback: CPU tensors raise.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .bridge import ProjectorCache, cast_weight_bf16, linear_silu_forward


def _rows_bf16(x2: torch.Tensor, want_ln: bool, eps: float = 1e-5):
    """[rows, K] activations → bf16 with a 64-padded pitch (+ LayerNorm statistics)."""
    K = x2.shape[1]
    return ops.cast_rows(x2, torch.bfloat16, ops.pad_to(K), want_ln=want_ln, ln_eps=eps)


def _downsample(x: torch.Tensor, k: int) -> torch.Tensor:
    """drop trailing T % k frames and concat k neighbours (projector.py:19-24, :40-46)."""
    B, T, D = x.shape
    discard = T % k
    if discard:
        x = x[:, :-discard, :]
    T = x.size(1)
    return x.contiguous().view(B, T // k, D * k)


class EncoderProjectorLinearSiLU(nn.Module):
    """LayerNorm(in) → Linear(in, 2048) → SiLU → Linear(2048, out) — projector.py:129-151."""

    def __init__(self, config, bottleneck=2048):
        super().__init__()
        in_dim = config.encoder_dim
        out_dim = config.llm_dim
        self.norm = nn.LayerNorm(in_dim)
        self.ffn = nn.Sequential(
            nn.Linear(in_dim, bottleneck, bias=True),
            nn.SiLU(),
            nn.Linear(bottleneck, out_dim, bias=True),
        )
        nn.init.kaiming_uniform_(self.ffn[0].weight, a=math.sqrt(5))
        nn.init.zeros_(self.ffn[2].bias)
        self.k = 1
        self._cache = ProjectorCache()
        # "bf16" (default): bf16 operands, fp32 accumulation (≤1e-2 of the fp32 reference);
        # "fp32x3": three-term bf16 split on the same tensor cores, fp32-accurate (~1e-6), 6x the MMA work
        self.precision = "bf16"
        self._cache3 = ProjectorCache()
        self._cache_rows = ProjectorCache()

    def _forward_fp32x3(self, x):
        """fp32-accurate inference path (reference numerics: fp32 LayerNorm/Linear/SiLU/Linear)."""
        B, T, D = x.shape
        Hb, H = self.ffn[0].weight.shape[0], self.ffn[2].weight.shape[0]
        params = [self.norm.weight, self.norm.bias, self.ffn[0].weight, self.ffn[0].bias, self.ffn[2].weight, self.ffn[2].bias]

        def build():
            with torch.no_grad():
                w1s, k1, _, _, colsum = ops.split_bf16x3(self.ffn[0].weight.detach().float(), 1,
                                                         col_scale=self.norm.weight.detach().float().contiguous(), want_rowsum=True)
                dbias = (self.ffn[0].weight.detach().double() @ self.norm.bias.detach().double()
                         + self.ffn[0].bias.detach().double()).float()       # tiny host-side GEMV, once per weight version
                w2s, k2, _, _, _ = ops.split_bf16x3(self.ffn[2].weight.detach().float(), 1)
                return w1s, k1, colsum, dbias, w2s, k2, self.ffn[2].bias.detach().float().contiguous()
        w1s, k1, colsum, dbias, w2s, k2, b2 = self._cache3.get(params, build)
        xs, kx, mean, rstd, _ = ops.split_bf16x3(x.reshape(B * T, D).float(), 0, want_ln=True, ln_eps=self.norm.eps)
        h = torch.empty(B * T, Hb, dtype=torch.float32, device=x.device)
        ops.gemm_fp32x3(xs, w1s, B * T, Hb, kx, h, L.EPI_LNFOLD_SILU, dbias, rstd, mean, colsum)
        hs, kh, _, _, _ = ops.split_bf16x3(h, 0)
        y = torch.empty(B * T, H, dtype=torch.float32, device=x.device)
        ops.gemm_fp32x3(hs, w2s, B * T, H, kh, y, L.EPI_BIAS, b2)
        return y.view(B, T, H).to(x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)

    def folded_weights(self):
        """(W1·γ bf16 [2048, pad64(in)], colsum, W1β+b1, W2 bf16, b2 fp32), cached per parameter version."""
        params = [self.norm.weight, self.norm.bias, self.ffn[0].weight, self.ffn[0].bias,
                  self.ffn[2].weight, self.ffn[2].bias]

        def build():
            with torch.no_grad():
                w1g, colsum, dbias = ops.fold_layernorm(self.ffn[0].weight.detach(), self.norm.weight.detach(),
                                                        self.norm.bias.detach(), self.ffn[0].bias.detach())
                w2 = cast_weight_bf16(self.ffn[2].weight)
                b2 = self.ffn[2].bias.detach().float().contiguous()
            return w1g, colsum, dbias, w2, b2
        return self._cache.get(params, build)

    def forward_token_rows(self, rows, out_dtype=torch.float32):
        """Projector output ``[n_rows, out_dim]`` (packed) for text-simulated rows given as descriptors
        (``ops.TokenRows`` from ``sim``): same function as ``forward`` on the dense
        ``ctc_pseudo_posterior[_noise]`` tensor (ps-slm.py:337-409 → projector.py:149-151), computed as a column
        gather of W1 in fp32 instead of a GEMM over a [rows, 25055] matrix that is one-hot plus a constant."""
        norm, l1, l2 = self.norm, self.ffn[0], self.ffn[2]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import TokenRowLinearSiLUFunction
            return TokenRowLinearSiLUFunction.apply(rows, norm.weight, norm.bias, l1.weight, l1.bias, l2.weight,
                                                    l2.bias, norm.eps, out_dtype)
        params = [norm.weight, norm.bias, l1.weight, l1.bias, l2.weight, l2.bias]

        def build():
            with torch.no_grad():
                S, D = ops.linear_rowdots(l1.weight.detach(), norm.weight.detach(), norm.bias.detach(), l1.bias.detach())
                return S, D, cast_weight_bf16(l2.weight), l2.bias.detach().float().contiguous()
        S, D, w2, b2 = self._cache_rows.get(params, build)
        with torch.no_grad():
            _, h, _, _ = ops.tokrow_fwd(l1.weight.detach(), norm.weight.detach(), S, D, rows, norm.eps, want_z=False)
            n = rows.n_rows
            y = torch.empty(max(n, 1), l2.weight.shape[0], dtype=out_dtype, device=h.device)[:n]
            ops.gemm_bf16_tn(h, w2, n, l2.weight.shape[0], l1.weight.shape[0], y, L.EPI_BIAS, b2)
        return y

    def forward(self, x):                  # (B, T, in_dim)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import linear_silu_train
            return linear_silu_train(self, x)
        if self.precision == "fp32x3":
            return self._forward_fp32x3(x)
        B, T, D = x.shape
        w1g, colsum, dbias, w2, b2 = self.folded_weights()
        xb, mean, rstd = _rows_bf16(x.reshape(B * T, D), True, self.norm.eps)
        y = linear_silu_forward(xb, B * T, D, mean, rstd, w1g, colsum, dbias, w2, b2, x.dtype
                                if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        return y.view(B, T, -1)


class EncoderProjectorConcat(nn.Module):
    """k-frame concat → Linear(D·k, 2048) → ReLU → Linear(2048, llm_dim) — projector.py:29-50."""

    def __init__(self, config):
        super().__init__()
        self.k = config.encoder_projector_ds_rate
        self.encoder_dim = config.encoder_dim
        self.llm_dim = config.llm_dim
        self.linear1 = nn.Linear(self.encoder_dim * self.k, 2048)
        self.relu = nn.ReLU()
        self.linear2 = nn.Linear(2048, config.llm_dim)
        self._cache = ProjectorCache()

    def forward(self, x):
        x = _downsample(x, self.k)
        B, T, D = x.shape
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import LinearFunction
            out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
            h = LinearFunction.apply(x.reshape(B * T, D), self.linear1.weight, self.linear1.bias, True, torch.float32)
            y = LinearFunction.apply(h, self.linear2.weight, self.linear2.bias, False, out_dtype)
            return y.reshape(B, T, self.llm_dim)
        params = [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias]
        w1, b1, w2, b2 = self._cache.get(params, lambda: (
            cast_weight_bf16(self.linear1.weight), self.linear1.bias.detach().float().contiguous(),
            cast_weight_bf16(self.linear2.weight), self.linear2.bias.detach().float().contiguous()))
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        xb, _, _ = _rows_bf16(x.reshape(B * T, D), False)
        h = torch.empty(B * T, 2048, dtype=torch.bfloat16, device=x.device)
        ops.gemm_bf16_tn(xb, w1, B * T, 2048, D, h, L.EPI_BIAS_RELU, b1)
        y = torch.empty(B * T, self.llm_dim, dtype=out_dtype, device=x.device)
        ops.gemm_bf16_tn(h, w2, B * T, self.llm_dim, 2048, y, L.EPI_BIAS, b2)
        return y.view(B, T, self.llm_dim)


class EncoderProjectorLinear(nn.Module):
    """k-frame concat → Linear(D·k, llm_dim) — projector.py:10-26 (CTC head over the LLM vocab)."""

    def __init__(self, config):
        super().__init__()
        self.k = config.encoder_projector_ds_rate
        self.encoder_dim = config.encoder_dim
        self.llm_vocab = config.llm_dim
        self.map = nn.Linear(self.encoder_dim * self.k, self.llm_vocab, bias=True)
        self._cache = ProjectorCache()

    def forward(self, x):
        x = _downsample(x, self.k)
        B, T, D = x.shape
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import LinearFunction
            out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
            y = LinearFunction.apply(x.reshape(B * T, D), self.map.weight, self.map.bias, False, out_dtype)
            return y.reshape(B, T, self.llm_vocab)
        w, b = self._cache.get([self.map.weight, self.map.bias], lambda: (
            cast_weight_bf16(self.map.weight), self.map.bias.detach().float().contiguous()))
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        xb, _, _ = _rows_bf16(x.reshape(B * T, D), False)
        ld = ops.pad_to(self.llm_vocab, 8)
        y = torch.empty(B * T, ld, dtype=out_dtype, device=x.device)
        ops.gemm_bf16_tn(xb, w, B * T, self.llm_vocab, D, y, L.EPI_BIAS, b)
        return y.view(B, T, ld)[:, :, :self.llm_vocab]


PROJECTORS = {
    "linear": EncoderProjectorConcat,
    "linear-silu": EncoderProjectorLinearSiLU,
    "simple_linear": EncoderProjectorLinear,
}
