"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python mirror of the descriptor struct matches the header, the module surface matches
the reference's (names, state_dict keys), and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "countr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(countr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from countr_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/countr_b200.h but not exported"
    assert lib.countr_version().startswith(b"countr_b200")


def test_gemm_desc_mirror_matches_header():
    from countr_b200._lib import GemmDesc
    src = open(os.path.join(ROOT, "include", "countr_b200.h")).read()
    body = src[src.index("typedef struct countr_gemm_desc {") + len("typedef struct countr_gemm_desc {"):src.index("} countr_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        decl = re.sub(r"^(const\s+)?(void|float|double|int32_t|int64_t)\s*\*?", "", decl).strip()
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f[0] for f in GemmDesc._fields_]


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from countr_b200 import _lib
    import models_mae_cross as M
    with pytest.raises(_lib.CountrError):
        _lib.require_device()
    m = M.SupervisedMAE(embed_dim=128, depth=1, num_heads=2)
    with pytest.raises(_lib.CountrError):
        m(torch.rand(1, 3, 384, 384), torch.rand(1, 3, 3, 64, 64), 3)


def test_module_surface_matches_reference():
    import models_crossvit as X
    import models_mae_cross as M
    from oracle import synth
    for name in ("mae_vit_base_patch16", "mae_vit_base4_patch16", "mae_vit_base6_patch16", "mae_vit_large_patch16",
                 "mae_vit_huge_patch14", "mae_vit_base_patch16_dec512d8b", "mae_vit_large_patch16_dec512d8b",
                 "mae_vit_huge_patch14_dec512d8b", "mae_vit_base_patch16_fim4", "mae_vit_base_patch16_fim6", "SupervisedMAE"):
        assert hasattr(M, name)
    for name in ("drop_path", "DropPath", "to_2tuple", "Mlp", "Attention", "CrossAttention", "CrossAttentionBlock"):
        assert hasattr(X, name)
    m = M.__dict__["mae_vit_base_patch16"](norm_pix_loss="store_true")          # demo.py:193 passes a string
    spec = synth.state_dict_spec(synth.CONFIGS["base"])
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    assert all(tuple(sd[k].shape) == tuple(s) for k, s, _ in spec)
    assert sum(p.numel() for p in m.parameters()) == 99_690_625
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 98_953_345
    assert not m.pos_embed.requires_grad and not m.decoder_pos_embed.requires_grad
    # init-time arithmetic: sin-cos tables identical to the oracle restatement of util/pos_embed.py
    from oracle import countr_oracle as O
    assert torch.equal(m.pos_embed[0], O.sincos_2d(768, 24)) and torch.equal(m.decoder_pos_embed[0], O.sincos_2d(512, 24))
    # parameters that get a gradient for each shot count (DDP find_unused_parameters contract)
    n3, _ = m._decoder_params(3)
    n0, _ = m._decoder_params(0)
    assert sorted(n3) == sorted(O.decoder_param_names(dict(sd), 3)) and sorted(n0) == sorted(O.decoder_param_names(dict(sd), 0))
    named = dict(m.named_parameters())
    assert sum(named[n].numel() for n in n3) == 13_306_241      # 53.2 MB of fp32 gradients
    assert sum(named[n].numel() for n in n0) == 11_755_777      # zero-shot: 47.0 MB


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "countr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_decoder_weight_plan_covers_every_gemm_operand():
    """The batched weight-refresh plan (Engine.decoder_weight_plan) must name every decoder weight that is consumed as a
    16-bit GEMM / conv operand — a tensor missing from the plan falls back to a per-tensor refresh launch (and, before the
    plan existed, fused optimizers left such copies stale).  Pure host logic: runs without a GPU."""
    import models_mae_cross as M
    from countr_b200.engine import Engine
    m = M.SupervisedMAE(embed_dim=128, depth=1, num_heads=2, decoder_depth=2)
    fp32_direct = {"decoder_proj1.0.weight", "decode_head3.3.weight"}       # direct conv / 1x1 dot kernels read the fp32 masters
    names = {id(p): n for n, p in m.named_parameters()}
    for shot, train in ((3, True), (0, True), (2, False), (0, False)):
        plan = Engine.decoder_weight_plan(m, shot, train)
        assert len(plan) == len({(id(p), k) for p, k in plan}), "duplicate (parameter, layout) in the plan"
        kinds = {}
        for p, k in plan:
            kinds.setdefault(names[id(p)], set()).add(k)
        want_w = [n for n, p in m.named_parameters()
                  if p.ndim >= 2 and n not in fp32_direct and not n.startswith(("blocks.", "patch_embed.")) and "pos_embed" not in n
                  and (shot > 0 or not n.startswith("decoder_proj"))]
        for n in want_w:
            fwd = "c0" if (n.startswith("decoder_proj") or n.startswith("decode_head")) else "w"
            assert fwd in kinds.get(n, ()), (shot, train, n, kinds.get(n))
            if train and n != "decoder_embed.weight":      # decoder_embed needs no dX: its input is the frozen encoder's output
                assert ("c1" if fwd == "c0" else "wt") in kinds[n], (shot, train, n, kinds[n])
        assert ("shot_token" in kinds) == (shot == 0)
        assert not any(n.startswith(("blocks.", "patch_embed.")) for n in kinds), "encoder weights are frozen: not part of the per-step plan"
