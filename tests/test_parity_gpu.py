"""Model-level parity on the GPU: the sm_100a path (through models_mae_cross.SupervisedMAE and the
C ABI) against (a) the committed goldens produced by the reference itself and (b) the CPU oracle
evaluated on the same seeded inputs.

Tolerance (BASELINE.json north_star): density map within 1e-3 relative (global rel-L2 over the
[N,384,384] map) and the count sum/60 within 1e-3 relative, operands fp16 / accumulation fp32.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import countr_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAP_TOL = 1e-3
COUNT_TOL = 1e-3


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(cfg_name, seed, dev):
    import models_mae_cross as M
    from functools import partial
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg, seed=seed)
    m = M.SupervisedMAE(img_size=cfg["img_size"], patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                        num_heads=cfg["num_heads"], decoder_embed_dim=cfg["decoder_embed_dim"],
                        decoder_depth=cfg["decoder_depth"], decoder_num_heads=cfg["decoder_num_heads"],
                        mlp_ratio=cfg["mlp_ratio"], norm_layer=partial(torch.nn.LayerNorm, eps=cfg["eps"]))
    m.load_state_dict(sd, strict=True)
    return m.to(dev), sd, cfg


def test_base_c1_matches_reference_golden(cuda):
    """BASELINE config 1 (demo.py path): one image, 3 exemplars, base model, fp32 in/out."""
    g = np.load(os.path.join(GOLD, "base_c1.npz"))
    m, sd, cfg = build("base", 0, cuda)
    m.eval()
    imgs, boxes = synth.make_inputs(1, seed=1234)
    with torch.no_grad():
        lat = m.forward_encoder(imgs.to(cuda))
        out, = m(imgs.to(cuda), boxes.to(cuda), 3)          # `output, = model(...)` as demo.py:139
        out0 = m(imgs.to(cuda), torch.empty(1, 0, device=cuda), 0)
    assert out.shape == (384, 384) and out.dtype == torch.float32
    e_lat = rel(lat[0, :8, :32], g["latent_head"])
    e_map = rel(out, g["out"][0])
    e_cnt = abs(out.sum().item() - g["out"].sum()) / abs(g["out"].sum())
    e_zero = rel(F.avg_pool2d(out0[:, None], 8)[:, 0], g["out_zero_pool8"])
    print(f"\n[parity base C1] latent relL2={e_lat:.3e} map relL2={e_map:.3e} count rel={e_cnt:.3e} zero-shot pool8 relL2={e_zero:.3e}")
    assert e_map < MAP_TOL and e_cnt < COUNT_TOL and e_zero < MAP_TOL
    assert e_lat < 5e-3


@pytest.mark.parametrize("shot", [0, 1, 2, 3, 5])
def test_small_all_shots_match_golden_and_oracle(cuda, shot):
    g = np.load(os.path.join(GOLD, "small_fwd.npz"))
    m, sd, cfg = build("small", 1, cuda)
    m.eval()
    imgs, boxes = synth.make_inputs(2, seed=77, shots=5)
    bx = boxes.to(cuda) if shot else torch.empty(2, 0, device=cuda)
    with torch.no_grad():
        out = m(imgs.to(cuda), bx, shot)
        ref = O.forward(sd, cfg, imgs, boxes if shot else torch.empty(2, 0), shot)
    e_or = rel(out, ref)
    e_pool = rel(F.avg_pool2d(out[:, None], 8)[:, 0], g[f"out_pool8_s{shot}"])
    e_sum = rel(out.sum((1, 2)), g[f"sum_s{shot}"])
    print(f"\n[parity small shot={shot}] vs oracle relL2={e_or:.3e} vs golden pool8={e_pool:.3e} sums={e_sum:.3e}")
    assert e_or < MAP_TOL and e_pool < MAP_TOL and e_sum < COUNT_TOL


def test_input_conventions(cuda):
    """fp16 inputs (FSC_finetune_cross.py:273-275), a non-contiguous window view (demo.py:139) and
    batch-independence (same image alone or inside a batch)."""
    m, sd, cfg = build("small", 1, cuda)
    m.eval()
    imgs, boxes = synth.make_inputs(3, seed=5)
    wide = torch.rand(3, 3, 384, 512)
    wide[:, :, :, 64:448] = imgs
    with torch.no_grad():
        a = m(imgs.to(cuda), boxes.to(cuda), 3)
        b = m(wide.to(cuda)[:, :, :, 64:448], boxes.to(cuda), 3)
        c = m(imgs.to(cuda).half(), boxes.to(cuda).half(), 3)
        d = m(imgs[1:2].to(cuda), boxes[1:2].to(cuda), 3)
    assert torch.equal(a, b)          # the forward pass is bit-reproducible (no order-dependent fp32 atomics)
    assert c.dtype == torch.float16 and rel(c.float(), a) < 2e-3
    assert rel(d[0], a[1]) < 1e-3     # other tile shapes / statistics splits at batch 1: same math, different rounding
    with pytest.raises(AssertionError):
        m(torch.rand(1, 3, 256, 256, device=cuda), boxes[:1].to(cuda), 3)   # timm PatchEmbed's size assert


def test_submodules_standalone(cuda):
    """models_crossvit classes keep working on their own (inference) and match the oracle."""
    import models_crossvit as X
    torch.manual_seed(0)
    blk = X.CrossAttentionBlock(512, 16, 4., qkv_bias=True, norm_layer=torch.nn.LayerNorm).to(cuda)
    for p in blk.parameters():
        torch.nn.init.normal_(p, std=0.05)
    sd = {"b." + k: v.detach().cpu() for k, v in blk.state_dict().items()}
    x = torch.randn(2, 576, 512)
    y = torch.randn(2, 3, 512)
    out = blk(x.to(cuda), y.to(cuda))
    ref = O.fim_block(x, y, sd, "b", 16, 1e-5)
    assert rel(out, ref) < 2e-3


def test_large_width_and_deeper_fim_variant(cuda):
    """Kernel templates generalise over (D, heads, FIM depth): ViT-L width with a 4-block FIM."""
    m, sd, cfg = build("large_fim4", 3, cuda)
    m.eval()
    imgs, boxes = synth.make_inputs(1, seed=9)
    with torch.no_grad():
        out = m(imgs.to(cuda), boxes.to(cuda), 2)
        ref = O.forward(sd, cfg, imgs, boxes, 2)
    e = rel(out, ref)
    print(f"\n[parity large_fim4] relL2={e:.3e}")
    assert e < MAP_TOL


def test_huge_patch14_geometry_runs_and_matches(cuda):
    """mae_vit_huge_patch14's geometry: 14-px patches that do not divide the 384-px frame (27 x 27 tokens,
    zero-padded 588 -> 592 patch rows), 80-channel heads (generic attention kernel), odd token grids through the FIM and the
    density head (27 -> 54 -> 108 -> 216 -> 432 px).  Inference only."""
    m, sd, cfg = build("huge_d2", 5, cuda)      # (seed 4 draws a final 1x1 conv whose terms cancel 7x: a conditioning, not a kernel, effect)
    m.eval()
    imgs, boxes = synth.make_inputs(1, seed=10)
    with torch.no_grad():
        out = m(imgs.to(cuda), boxes.to(cuda), 3)
        ref = O.forward(sd, cfg, imgs, boxes, 3)
    assert out.shape == ref.shape == (1, 432, 432)
    e = rel(out, ref)
    print(f"\n[parity huge_d2] relL2={e:.3e}")
    assert e < MAP_TOL
