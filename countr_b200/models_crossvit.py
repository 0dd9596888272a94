"""Drop-in for the reference's models_crossvit.py: same class names, constructor arguments and
state_dict keys (models_crossvit.py:9-156).  The nn.Linear / nn.LayerNorm children only own the
fp32 master parameters; the arithmetic runs in the sm_100a kernels of libcountr_sm100.so.

The stand-alone `forward` of each class is an inference path through the same kernels (used by
the module-level parity tests); the training path goes through SupervisedMAE.forward, whose
decoder is one autograd node (countr_b200/models_mae_cross.py).
"""
import collections.abc
from itertools import repeat

import torch
import torch.nn as nn

from . import ops
from .engine import F16, F32, _contig32, engine


def drop_path(x, drop_prob: float = 0., training: bool = False, scale_by_keep: bool = True):
    """Stochastic depth per sample (models_crossvit.py:9-25).  The reference never enables it
    (every CrossAttentionBlock / Block is built with drop_path=0), so this is host-side glue."""
    if drop_prob == 0. or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
    if keep_prob > 0.0 and scale_by_keep:
        random_tensor.div_(keep_prob)
    return x * random_tensor


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training, self.scale_by_keep)


def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable):
            return x
        return tuple(repeat(x, n))
    return parse


to_2tuple = _ntuple(2)


def _no_dropout(*ps):
    for p in ps:
        if p:
            raise NotImplementedError("countr_b200: dropout > 0 is not on the reference's hot path (always 0 there)")


def _tokens16(x):
    """[B, N, C] float tensor -> fp16 [B*N, C] GEMM operand."""
    B, N, C = x.shape
    x32 = x.detach().to(F32).contiguous().view(B * N, C)
    x16 = torch.empty(B * N, C, dtype=F16, device=x.device)
    ops.cast16(x32, x16)
    return x16


class Mlp(nn.Module):
    """fc1 -> exact GELU -> fc2 (models_crossvit.py:46-67)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("countr_b200 Mlp implements the reference's nn.GELU only")
        _no_dropout(*to_2tuple(drop))
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    @torch.no_grad()
    def forward(self, x):
        B, N, _ = x.shape
        wc = engine().wc
        x16 = _tokens16(x)
        u = torch.empty(B * N, self.fc1.weight.shape[0], dtype=F16, device=x.device)
        ops.linear(x16, wc.w16(self.fc1.weight), u, bias=_contig32(self.fc1.bias), act=1)
        y = torch.empty(B * N, self.fc2.weight.shape[0], dtype=F32, device=x.device)
        ops.linear(u, wc.w16(self.fc2.weight), y, bias=_contig32(self.fc2.bias))
        return y.view(B, N, -1).to(x.dtype)


class Attention(nn.Module):
    """Multi-head self-attention (models_crossvit.py:69-94)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        _no_dropout(attn_drop, proj_drop)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    @torch.no_grad()
    def forward(self, x):
        B, N, C = x.shape
        wc = engine().wc
        x16 = _tokens16(x)
        qkv = torch.empty(B * N, 3 * C, dtype=F16, device=x.device)
        ops.linear(x16, wc.w16(self.qkv.weight), qkv, bias=None if self.qkv.bias is None else _contig32(self.qkv.bias))
        att = torch.empty(B * N, C, dtype=F16, device=x.device)
        ops.attention_fwd(qkv, att, B, N, self.num_heads, C // self.num_heads, self.scale)
        y = torch.empty(B * N, C, dtype=F32, device=x.device)
        ops.linear(att, wc.w16(self.proj.weight), y, bias=_contig32(self.proj.bias))
        return y.view(B, N, C).to(x.dtype)


class CrossAttention(nn.Module):
    """Image tokens attend to the exemplar tokens (models_crossvit.py:96-128)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        _no_dropout(attn_drop, proj_drop)
        self.wq = nn.Linear(dim, dim, bias=qkv_bias)
        self.wk = nn.Linear(dim, dim, bias=qkv_bias)
        self.wv = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    @torch.no_grad()
    def forward(self, x, y):
        B, Nx, C = x.shape
        Ny = y.shape[1]
        wc = engine().wc
        b = lambda lin: None if lin.bias is None else _contig32(lin.bias)  # noqa: E731
        x16, y16 = _tokens16(x), _tokens16(y)
        q = torch.empty(B * Nx, C, dtype=F16, device=x.device)
        ops.linear(x16, wc.w16(self.wq.weight), q, bias=b(self.wq))
        k = torch.empty(B * Ny, C, dtype=F32, device=x.device)
        v = torch.empty(B * Ny, C, dtype=F32, device=x.device)
        ops.linear(y16, wc.w16(self.wk.weight), k, bias=b(self.wk))
        ops.linear(y16, wc.w16(self.wv.weight), v, bias=b(self.wv))
        c = torch.empty(B * Nx, C, dtype=F16, device=x.device)
        ops.cross_attn_core(q, k, v, c, B, Nx, Ny, C, C // self.num_heads, self.scale)
        out = torch.empty(B * Nx, C, dtype=F32, device=x.device)
        ops.linear(c, wc.w16(self.proj.weight), out, bias=_contig32(self.proj.bias))
        return out.view(B, Nx, C).to(x.dtype)


class CrossAttentionBlock(nn.Module):
    """x += selfattn(norm0 x); x += attn(norm1 x, y); x += mlp(norm2 x) (models_crossvit.py:130-156)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop_path:
            raise NotImplementedError("countr_b200: drop_path > 0 is not on the reference's hot path")
        self.norm0 = norm_layer(dim)
        self.selfattn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                  proj_drop=drop)
        self.drop_path0 = nn.Identity()
        self.norm1 = norm_layer(dim)
        self.attn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                   proj_drop=drop)
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.drop_path2 = nn.Identity()

    def _ln(self, norm, x):
        B, N, C = x.shape
        x32 = x.detach().to(F32).contiguous().view(B * N, C)
        y32 = torch.empty_like(x32)
        ops.layernorm_fwd(x32, _contig32(norm.weight), _contig32(norm.bias), norm.eps, y32=y32)
        return y32.view(B, N, C)

    @torch.no_grad()
    def forward(self, x, y):
        x = x.to(F32)
        x = x + self.selfattn(self._ln(self.norm0, x))
        x = x + self.attn(self._ln(self.norm1, x), y)
        x = x + self.mlp(self._ln(self.norm2, x))
        return x
