"""Run-to-run reproducibility of one fine-tune step (SURVEY.md §7 "loss curves track step-for-step"): the forward pass and the
loss are bit-identical between two runs on identical inputs; the gradients are accumulated with fp32 atomics / TMA reduce-adds
(split-K weight gradients, GroupNorm parameter gradients), so they agree to rounding-order noise, not bitwise."""
import pytest
import torch

from oracle import synth
from test_parity_gpu import build

pytestmark = pytest.mark.gpu


def _step(m, imgs, boxes, gt, mask):
    for p in m.parameters():
        p.grad = None
    out = m(imgs, boxes, 3)
    loss = ((out - gt) ** 2 * mask / (384 * 384)).sum() / imgs.shape[0]
    (loss * 4096.0).backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    return out.detach().clone(), loss.detach().clone(), grads


def test_step_is_reproducible_run_to_run(cuda):
    m, sd, cfg = build("small", 1, cuda)
    m.train()
    imgs, boxes = synth.make_inputs(2, seed=3)
    gt, mask = synth.make_targets(2, seed=4)
    args = [t.to(cuda) for t in (imgs, boxes, gt, mask)]
    out1, loss1, g1 = _step(m, *args)
    out2, loss2, g2 = _step(m, *args)
    assert torch.equal(out1, out2), "the forward pass must be bit-reproducible"
    assert torch.equal(loss1, loss2)
    assert g1.keys() == g2.keys()
    num = sum((g1[k].double() - g2[k].double()).pow(2).sum().item() for k in g1)
    den = sum(g1[k].double().pow(2).sum().item() for k in g1)
    rel = (num / den) ** 0.5
    exact = sum(int(torch.equal(g1[k], g2[k])) for k in g1)
    print(f"\n[repro] gradients: relL2 between two runs {rel:.2e}; {exact}/{len(g1)} tensors bit-identical")
    print("[repro] not bit-identical:", ", ".join(k for k in g1 if not torch.equal(g1[k], g2[k])))
    assert rel < 1e-5
