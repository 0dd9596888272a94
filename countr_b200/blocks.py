"""Forward / backward kernel schedules of one pre-norm ViT block (timm 0.4.9 `Block`:
x += attn(norm1 x); x += mlp(norm2 x)) on the fp32 residual stream, with every activation the backward
needs kept in its own buffer.  Used by the MAE pre-training model (models_mae_noct), where the encoder
is trained too; the fine-tune path has its own in-place (frozen encoder) and FIM schedules in engine.py.
"""
import torch

from . import ops
from .backward import attention_backward
from .engine import F16, F32, _contig32


def vit_block_forward(wc, blk, x, B, L, save):
    """x fp32 [B*L, D] -> new fp32 [B*L, D]; `save` (list) receives the tape entry."""
    dev = x.device
    M, D = x.shape
    H = blk.attn.num_heads
    dh = D // H
    hid = blk.mlp.fc1.weight.shape[0]
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
    h1, mean1, rstd1 = e((M, D), F16), e((M,), F32), e((M,), F32)
    ops.layernorm_fwd(x, _contig32(blk.norm1.weight), _contig32(blk.norm1.bias), blk.norm1.eps, y16=h1, mean=mean1, rstd=rstd1)
    qkv = e((M, 3 * D), F16)
    ops.linear(h1, wc.w16(blk.attn.qkv.weight), qkv, bias=_contig32(blk.attn.qkv.bias))
    att, lse = e((M, D), F16), e((B, H, L), F32)
    ops.attention_fwd(qkv, att, B, L, H, dh, blk.attn.scale, lse=lse)
    x1 = e((M, D), F32)
    ops.linear(att, wc.w16(blk.attn.proj.weight), x1, bias=_contig32(blk.attn.proj.bias), residual=x)
    h2, mean2, rstd2 = e((M, D), F16), e((M,), F32), e((M,), F32)
    ops.layernorm_fwd(x1, _contig32(blk.norm2.weight), _contig32(blk.norm2.bias), blk.norm2.eps, y16=h2, mean=mean2, rstd=rstd2)
    u, pre = e((M, hid), F16), (e((M, hid), F16) if save is not None else None)     # the pre-activation is only kept for the backward
    ops.linear(h2, wc.w16(blk.mlp.fc1.weight), u, bias=_contig32(blk.mlp.fc1.bias), act=1, aux=pre)
    x2 = e((M, D), F32)
    ops.linear(u, wc.w16(blk.mlp.fc2.weight), x2, bias=_contig32(blk.mlp.fc2.bias), residual=x1)
    if save is not None:
        save.append(dict(x0=x, h1=h1, mean1=mean1, rstd1=rstd1, qkv=qkv, att=att, lse=lse, x1=x1, h2=h2, mean2=mean2, rstd2=rstd2,
                         u=u, pre=pre, B=B, L=L))
    return x2


def vit_block_backward(wc, blk, s, g, g16, G, jobs, next_bias=None):
    """g (fp32, updated in place) holds dL/dx_out on entry and dL/dx_in on exit; g16 is the fp16 copy of the entry gradient.
    Returns the fp16 copy of the exit gradient (a fresh buffer: the deferred weight-gradient jobs keep reading the old ones).
    G(param) -> the fp32 gradient view to accumulate into; jobs: backward.GradJobs (weight / bias gradients are deferred);
    next_bias: bias of the Linear that consumed x_in's producer (fc2 of the block below, patch-embed / decoder_embed), its
    gradient = column sums of the exit gradient, emitted by the last LayerNorm backward.  The caller provides the column sums
    of the ENTRY gradient (this block's fc2.bias) the same way."""
    dev = g.device
    M, D = g.shape
    H = blk.attn.num_heads
    dhd = D // H
    hid = blk.mlp.fc1.weight.shape[0]
    B, L = s["B"], s["L"]
    dh = torch.empty(M, D, dtype=F32, device=dev)
    # MLP
    jobs.dW(g16, s["u"], G(blk.mlp.fc2.weight))
    dpre = torch.empty(M, hid, dtype=F16, device=dev)
    ops.linear(g16, wc.w16_t(blk.mlp.fc2.weight), dpre, act=2, aux=s["pre"])
    jobs.dB(dpre, G(blk.mlp.fc1.bias))
    jobs.dW(dpre, s["h2"], G(blk.mlp.fc1.weight))
    ops.linear(dpre, wc.w16_t(blk.mlp.fc1.weight), dh)
    g16 = torch.empty(M, D, dtype=F16, device=dev)
    ops.layernorm_bwd(dh, s["x1"], _contig32(blk.norm2.weight), s["mean2"], s["rstd2"], g, G(blk.norm2.weight), G(blk.norm2.bias),
                      accumulate=True, dx16=g16, dx_colsum=G(blk.attn.proj.bias), jobs=jobs)
    # attention
    jobs.dW(g16, s["att"], G(blk.attn.proj.weight))
    datt = torch.empty(M, D, dtype=F16, device=dev)
    ops.linear(g16, wc.w16_t(blk.attn.proj.weight), datt)
    dqkv = attention_backward(s["qkv"], s["lse"], datt, B, L, H, dhd, blk.attn.scale, att=s["att"])
    jobs.dB(dqkv, G(blk.attn.qkv.bias))
    jobs.dW(dqkv, s["h1"], G(blk.attn.qkv.weight))
    ops.linear(dqkv, wc.w16_t(blk.attn.qkv.weight), dh)
    g16 = torch.empty(M, D, dtype=F16, device=dev)
    ops.layernorm_bwd(dh, s["x0"], _contig32(blk.norm1.weight), s["mean1"], s["rstd1"], g, G(blk.norm1.weight), G(blk.norm1.bias),
                      accumulate=True, dx16=g16, dx_colsum=None if next_bias is None else G(next_bias), jobs=jobs)
    return g16
