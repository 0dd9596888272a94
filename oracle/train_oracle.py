"""CPU ORACLE (test infrastructure only) for the script-side pieces of the fine-tune step
(FSC_finetune_cross.py:290-303, util/misc.py:260-301); see oracle/countr_oracle.py for the rules that apply."""
import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def bernoulli_mask(seed, step, hw, keep_prob=0.8):
    """The device-side draw of `countr_finetune_loss` (mask == NULL): splitmix64 finaliser over (seed, step, pixel), keep when
    the top 24 bits fall below keep_prob * 2^24.  Stands in for np.random.binomial(n=1, p=0.8, size=[384, 384])
    (FSC_finetune_cross.py:290): same distribution, a stream that lives on the device."""
    with np.errstate(over="ignore"):
        pix = np.arange(hw, dtype=np.uint64)
        z = np.uint64(seed) + np.uint64(step) * np.uint64(0xD1B54A32D192ED03) + (pix + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    thr = np.uint64(int(float(np.float32(keep_prob)) * 16777216.0))
    return ((z >> np.uint64(40)) < thr).astype(np.uint8)


def loss_and_counts(out, gt, mask):
    """FSC_finetune_cross.py:290-303 for one batch: (loss, pred_cnt, gt_cnt, batch_mae, batch_mse)."""
    B = out.shape[0]
    loss = ((out - gt) ** 2 * mask / (out.shape[1] * out.shape[2])).sum() / B
    pred = out.reshape(B, -1).sum(1) / 60
    gtc = gt.reshape(B, -1).sum(1) / 60
    err = (pred - gtc).abs().float()
    return loss, pred, gtc, err.double().mean(), (err ** 2).double().mean()


def grad_norm(grads):
    """util/misc.py:289-301 get_grad_norm_ (norm_type 2)."""
    return torch.norm(torch.stack([torch.norm(g.detach(), 2.0) for g in grads]), 2.0)


class ScalerState:
    """torch.cuda.amp.GradScaler's update rule (init 65536, growth 2, backoff 0.5, interval 2000)."""

    def __init__(self, scale=65536.0, growth=2.0, backoff=0.5, interval=2000):
        self.scale, self.growth, self.backoff, self.interval, self.tracker = scale, growth, backoff, interval, 0

    def update(self, found_inf):
        if found_inf:
            self.scale *= self.backoff
            self.tracker = 0
        else:
            self.tracker += 1
            if self.tracker == self.interval:
                self.scale *= self.growth
                self.tracker = 0
