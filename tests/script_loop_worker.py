"""Script-faithful fine-tune loop, run in a SUBPROCESS by tests/test_script_loop_gpu.py (not collected by pytest).

Everything FSC_finetune_cross.py does around `model(...)` is executed as the script does it, with the script's own
conditions — none of which the kernel-level tests exercise:
  * CUDA_LAUNCH_BLOCKING=1 in the environment before CUDA starts (:110)
  * a 1-rank NCCL process group and DDP(model, device_ids=[gpu], find_unused_parameters=True) (:228-231)
  * timm add_weight_decay groups + torch.optim.AdamW (:234-235)
  * fp16 samples / gt_density / boxes (:273-275), forward under torch.cuda.amp.autocast() (:286-287)
  * the int64 mask tiled to [B, 384, 384] and the loss expression in fp16 (:290-295) -> an fp16 grad_out reaches the decoder
  * counts / MAE bookkeeping (:298-303), NativeScalerWithGradNormCount = real GradScaler (dynamic scale, unscale_, inf skip)
    + get_grad_norm_ (util/misc.py:260-301), optimizer.zero_grad(), lr written into the param groups every iteration
  * shot_num changing from step to step (:278-284; here the fixed sequence of the reference golden curve)
GradScaler starts at 65536: the fp16 loss gradient overflows on the first iteration (65536 is not representable in fp16), so
the first optimizer step is SKIPPED and the scale backs off — the inf path is exercised without any injection.  A skipped
iteration is repeated with the same batch, so the sequence of applied updates is the golden curve's.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def get_grad_norm_(parameters, norm_type=2.0):          # util/misc.py:289-301
    parameters = [p for p in parameters if p.grad is not None]
    if len(parameters) == 0:
        return torch.tensor(0.)
    device = parameters[0].grad.device
    return torch.norm(torch.stack([torch.norm(p.grad.detach(), norm_type).to(device) for p in parameters]), norm_type)


class NativeScalerWithGradNormCount:                     # util/misc.py:257-286
    def __init__(self):
        self._scaler = torch.cuda.amp.GradScaler()

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if update_grad:
            self._scaler.unscale_(optimizer)
            norm = get_grad_norm_(parameters)
            self._scaler.step(optimizer)
            self._scaler.update()
        else:
            norm = None
        return norm


def main(out_path):
    assert os.environ.get("CUDA_LAUNCH_BLOCKING") == "1"
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from oracle import synth
    from test_parity_gpu import build
    torch.cuda.set_device(0)
    device = torch.device("cuda:0")
    dist.init_process_group("nccl", rank=0, world_size=1)
    C = synth.CURVE
    m, sd, cfg = build("small", 1, device)
    m.train()
    start = {n: p.detach().clone() for n, p in m.named_parameters()}
    model = DDP(m, device_ids=[0], find_unused_parameters=True)
    model_without_ddp = model.module
    optimizer = torch.optim.AdamW(synth.weight_decay_groups(model_without_ddp.named_parameters(), C["weight_decay"]), lr=C["lr"], betas=C["betas"])
    loss_scaler = NativeScalerWithGradNormCount()
    batches = synth.curve_batches()
    records = []
    train_mae = torch.tensor([0], dtype=torch.float64, device=device)
    optimizer.zero_grad()
    it, iters = 0, 0
    while it < C["steps"] and iters < C["steps"] + 8:
        iters += 1
        for group in optimizer.param_groups:            # lr_sched.adjust_learning_rate writes the lr every iteration (:270)
            group["lr"] = C["lr"]
        imgs, boxes, gt, mask = batches[it % 2]
        samples = imgs.to(device, non_blocking=True, dtype=torch.half)
        gt_density = gt.to(device, non_blocking=True, dtype=torch.half)
        boxes_h = boxes.to(device, non_blocking=True, dtype=torch.half)
        shot_num = C["shots"][it]
        with torch.cuda.amp.autocast():
            output = model(samples, boxes_h, shot_num)
        assert output.dtype == torch.float16
        mask_np = mask.numpy().astype(np.int64)          # np.random.binomial(n=1, p=0.8, size=[384, 384]) stand-in
        masks = np.tile(mask_np, (output.shape[0], 1))
        masks = masks.reshape(output.shape[0], 384, 384)
        masks = torch.from_numpy(masks).to(device)
        loss = (output - gt_density) ** 2
        loss = (loss * masks / (384 * 384)).sum() / output.shape[0]
        with torch.no_grad():
            pred_cnt = (output.view(len(samples), -1)).sum(1) / 60
            gt_cnt = (gt_density.view(len(samples), -1)).sum(1) / 60
            cnt_err = torch.abs(pred_cnt - gt_cnt).float()
            batch_mae = cnt_err.double().mean()
            # (the script's fp16 `.sum(1)` overflows to inf once a count exceeds 65504 / 60 — the seeded random weights of this
            # test produce such maps; the fp32 sum of the same fp16 map is what is compared with the golden counts)
            count32 = output.float().view(len(samples), -1).sum(1) / 60
        train_mae += torch.nan_to_num(batch_mae, posinf=0.0)
        assert torch.isfinite(loss), "Loss is {}, stopping training".format(loss)
        scale_before = loss_scaler._scaler.get_scale()
        # fp32 restatement of the same loss from the fp16 map (the fp16 expression above quantises every pixel term to the
        # subnormal grid; this is the number compared with the golden curve)
        loss32 = (((output.float() - gt.to(device)) ** 2) * mask.to(device) / (384 * 384)).sum() / output.shape[0]
        norm = loss_scaler(loss, optimizer, parameters=model.parameters(), update_grad=True)
        n_grads = sum(1 for p in model.parameters() if p.grad is not None)
        optimizer.zero_grad()
        scale_after = loss_scaler._scaler.get_scale()
        skipped = scale_after < scale_before
        records.append(dict(it=it, shot=shot_num, loss16=float(loss), loss32=float(loss32), grad_norm=float(norm), scale=scale_before,
                            skipped=bool(skipped), n_grads=n_grads, count=[float(c) for c in count32.cpu()]))
        if not skipped:
            it += 1
    deltas = {n: float((p.detach() - start[n]).norm()) for n, p in model_without_ddp.named_parameters() if p.requires_grad}
    with open(out_path, "w") as f:
        json.dump(dict(records=records, deltas=deltas, train_mae=float(train_mae)), f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
