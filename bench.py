#!/usr/bin/env python
"""Headline benchmark: 384x384 images/sec of the CounTR hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload finetune|infer0|pretrain]

Workloads (BASELINE.json configs, per GPU; weak scaling — the per-GPU batch is fixed):
  finetune (default, configs[1]/[2]): ViT-B/16 SupervisedMAE, batch 8, 3 exemplar shots; one step = frozen-encoder forward +
            decoder forward + masked-MSE loss + decoder backward (+ gradient all-reduce for N > 1) + inf check / unscale /
            AdamW / loss-scale update, on synthetic data with seeded random weights
  infer0   (configs[3]): zero-shot inference (0 exemplars, learnable shot token), batch 128, forward only
  pretrain (configs[4]): MAE pre-training step of models_mae_noct (mask 0.5), batch 32 per GPU, forward + full backward +
            (all-reduce) + AdamW

  value : images/sec with the step's inputs already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same step driven through the public API with HOST (pinned) inputs: H2D of every input inside the timed region
          (prefetched on a copy stream while the previous step computes, as an input pipeline does) and a D2H read of the
          step's result (loss / counts) every step
  --impl reference : the reference's own algorithm (oracle/ port, fp32, torch CPU kernels, all host cores) on a bounded
          sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.md §2 (2*MAC, attention counted dense)
GFLOP_PER_IMG = {"finetune": 320.67, "infer0": 179.48, "pretrain": 262.62}
PER_GPU_BATCH = {"finetune": 8, "infer0": 128, "pretrain": 32}
SHOTS = 3
WORKLOAD_NAME = {
    "finetune": "FSC147 fine-tune: ViT-B/16, 384x384, 3 shots, batch 8 per GPU (BASELINE configs[1])",
    "infer0": "zero-shot inference (0 exemplars, learnable shot token), ViT-B/16, 384x384, batch 128 per GPU (BASELINE configs[3])",
    "pretrain": "MAE pre-train (models_mae_noct) 384x384, mask_ratio 0.5, batch 32 per GPU (BASELINE configs[4])",
}
METRIC = {"finetune": "images/sec (fine-tune step, 384x384)", "infer0": "images/sec (zero-shot forward, 384x384)",
          "pretrain": "images/sec (MAE pre-train step, 384x384)"}
NCU_SUMMARY = os.path.join("profiles", "r2_conv_h3_ncu_summary.json")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed `ncu --set full`
    capture of the same launch (ncu cannot run inside a timed bench); (bytes, source file) or (None, None)."""
    for rel in (NCU_SUMMARY, os.path.join("profiles", "r1_conv_h3_ncu_summary_v2.json")):
        try:
            with open(os.path.join(ROOT, rel)) as f:
                return json.load(f).get("dram_bytes_per_launch"), rel
        except Exception:
            continue
    return None, None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_steps(workload, batch, steps, warmup):
    """(images/sec, cores, seconds per step, sample description) of the oracle port on the host CPU."""
    import torch
    from oracle import countr_oracle as O
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    times = []
    if workload == "pretrain":
        from oracle import noct_oracle as NO
        cfg = dict(img_size=384, patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
                   decoder_num_heads=16, mlp_ratio=4, eps=1e-6)
        sd = NO.make_state_dict(cfg, seed=0)
        params = []
        for k in sd:
            if k not in ("pos_embed", "decoder_pos_embed"):
                sd[k] = sd[k].clone().requires_grad_(True)
                params.append(sd[k])
        opt = torch.optim.AdamW(params, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05)
        imgs, _ = synth.make_inputs(batch, seed=1)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            loss, _, _ = NO.forward(sd, cfg, imgs, 0.5, torch.rand(batch, 576), True)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        what = "MAE pre-train step (fwd + bwd + AdamW)"
    else:
        cfg = synth.CONFIGS["base"]
        sd = synth.make_state_dict(cfg, seed=0)
        if workload == "infer0":
            imgs, _ = synth.make_inputs(batch, seed=1)
            with torch.no_grad():
                for it in range(warmup + steps):
                    t0 = time.perf_counter()
                    O.forward(sd, cfg, imgs, torch.empty(batch, 0), 0)
                    if it >= warmup:
                        times.append(time.perf_counter() - t0)
            what = "zero-shot forward"
        else:
            names = O.decoder_param_names(sd, SHOTS)
            params = []
            for n in names:
                sd[n] = sd[n].clone().requires_grad_(True)
                params.append(sd[n])
            opt = torch.optim.AdamW(params, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05)
            imgs, boxes = synth.make_inputs(batch, seed=1)
            gt, mask = synth.make_targets(batch, seed=2)
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                out = O.forward(sd, cfg, imgs, boxes, SHOTS)
                loss = O.finetune_loss(out, gt, mask)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
            what = "fine-tune step (fwd + decoder bwd + AdamW)"
    sec = sum(times) / len(times)
    sample = f"oracle port of the reference {what}, fp32 torch CPU, batch {batch} per step, {steps} timed steps after {warmup} warm-up"
    return batch / sec, cores, sec, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    batch = {"finetune": 2, "infer0": 4, "pretrain": 2}[wl]          # a bounded sample of the per-GPU batch: ~0.4-1 s per step
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    ips, cores, sec, sample = cpu_steps(wl, batch, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC[wl], "value": round(ips, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[wl]},
        "cpu_baseline": {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.power = []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_max": round(max(self.power), 1) if self.power else None, "samples": len(s)}


def param_groups(model, weight_decay):
    """timm optim_factory.add_weight_decay (FSC_finetune_cross.py:234): no decay for 1-D params / biases."""
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or n.endswith(".bias")) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class Env:
    """Process-level plumbing shared by the workloads: ranks, NCCL, barrier, device-timed loops."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from countr_b200 import _lib
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        _lib.require_device()
        self.sm_budget = None
        if self.world > 1:
            import datetime
            # The gradient all-reduce overlaps compute (the next batch's frozen encoder / the rest of the backward).  Our GEMM and
            # attention kernels are persistent — one CTA (pair) per SM — so a collective that needs SMs while they run would
            # push part of every grid into a second wave.  Leave four SMs to NCCL (144-CTA grids lose nothing on the Linear and
            # attention shapes: 432 tiles = 3 x 144, 288 attention items = 2 x 144) and cap the collective at four CTAs.
            # (fine-tune only: measured at N = 2, 4.32 -> 4.23 ms per step.  The pre-training step's 446.6 MB all-reduce needs NCCL's
            # full CTA count; its sliced overlap — models_mae_noct grad_slice_hook, --overlap-pretrain — did not pay at N = 2:
            # 20.2 ms against 19.7 ms for one all-reduce between the backward and the update graph.)
            overlap = (args.workload == "finetune" and not args.no_overlap) or (args.workload == "pretrain" and args.overlap_pretrain)
            if overlap:
                os.environ.setdefault("NCCL_MAX_CTAS", "4")
                os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
                self.sm_budget = int(_lib.lib().countr_set_sm_budget(int(os.environ.get("COUNTR_SM_BUDGET", "144"))))
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=180))   # a stuck collective aborts
        self.steps, self.warmup = args.steps, max(args.warmup, 3)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, n):
        """n calls of fn(i) bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks (ms)."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for i in range(n):
            fn(i)
        ev1.record()
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def agree(self, ok):
        """All ranks must take the same path (graph replay vs eager)."""
        if self.world == 1:
            return ok
        flag = self.torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        return flag.item() != 0

    def done(self):
        """Leave without tearing NCCL down: destroy_process_group() can block forever while CUDA graphs that captured
        collectives are alive (observed at N = 2: the JSON line was out, the processes never exited).  Everything is flushed
        and every rank has passed the last barrier, so a hard exit loses nothing."""
        if self.world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            try:
                self.barrier()
            finally:
                os._exit(0)


def eager_warmup(torch, fn, n=3):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(n):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()


def dominant_kernel_roofline(env, B):
    """The largest single kernel of the step: the decode_head3 3x3 conv as an implicit GEMM (43.5 of 321 GFLOP/img forward
    alone), timed alone with CUDA events on the launching stream (burst clocks -> burst peak)."""
    torch = env.torch
    from countr_b200 import ops
    dev = env.dev
    Bk = min(B, 8)
    x16 = torch.randn(Bk, 192, 192, 256, device=dev).half()
    w16 = torch.randn(256, 9 * 256, device=dev).half()
    y16 = torch.empty(Bk, 192, 192, 256, device=dev, dtype=torch.float16)
    bias = torch.zeros(256, device=dev)
    stats = torch.zeros(Bk, 8, 2, device=dev, dtype=torch.float64)
    for _ in range(3):
        ops.conv3x3(x16, w16, y16, bias=bias, gn_stats=stats)
    reps = 10
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(reps):
        ops.conv3x3(x16, w16, y16, bias=bias, gn_stats=stats)
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / reps
    k_flop = 2.0 * Bk * 192 * 192 * 256 * 2304
    k_tflops = k_flop / (k_ms * 1e-3) / 1e12
    peaks, peak_src = load_peaks()
    traffic, traffic_src = ncu_traffic()
    return {"bound": "tensor", "kernel": "gemm_kernel<pair> (decode_head3 conv3x3 implicit GEMM, M=%d N=256 K=2304)" % (Bk * 192 * 192),
            "achieved": round(k_tflops, 1), "peak": peaks.get("bf16_tflops"), "unit": "TFLOP/s",
            "frac": round(k_tflops / peaks.get("bf16_tflops"), 4), "peak_source": peak_src + " bf16_tflops (burst: kernel timed alone)",
            "ms_per_launch": round(k_ms, 4), "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes": 2 * (2 * Bk * 192 * 192 * 256) + 2 * 256 * 2304}


def whole_step(value, world, wl, window_ms, sustained=None):
    """Whole-step tensor utilisation against the peak that matches the timing window: a sub-second window runs at burst
    clocks (burst peak); the multi-second window is compared with the sustained peak."""
    peaks, _ = load_peaks()
    tf = value / world * GFLOP_PER_IMG[wl] / 1e3
    out = {"achieved": round(tf, 1), "peak": peaks.get("bf16_tflops"), "frac": round(tf / peaks.get("bf16_tflops"), 4),
           "peak_kind": "burst (timed window %.0f ms)" % window_ms, "gflop_per_image": GFLOP_PER_IMG[wl]}
    if sustained is not None:
        tfs = sustained / world * GFLOP_PER_IMG[wl] / 1e3
        out["sustained"] = {"achieved": round(tfs, 1), "peak": peaks.get("bf16_tflops_sustained"),
                            "frac": round(tfs / peaks.get("bf16_tflops_sustained"), 4)}
    return out


def sustained_window(env, run_step, ms_per_step, batch):
    """>= 3 s of back-to-back steps (the default K steps last well under a second, i.e. burst clocks): the rate a long
    training run sees, with its own clock record."""
    n = max(10, int(3300.0 / max(ms_per_step, 1e-3)) + 1)
    sampler = ClockSampler(env.local)
    sampler.start()
    ms = env.timed(lambda i: run_step(), n)
    clocks = sampler.finish()
    return {"value": round(batch * env.world * n / (ms * 1e-3), 2), "unit": "images/s", "steps": n, "seconds": round(ms * 1e-3, 3),
            "ms_per_step": round(ms / n, 4), "clocks": clocks}


# ------------------------------------------------------------------------------------------------ fine-tune
def finetune_parity(env, model):
    """Forward of the benchmarked configuration (base model, batch 8, 3 shots, the seeded synthetic weights the bench
    trains from) against the UNMODIFIED reference's output on the same inputs (tests/golden/base_b8.npz, produced by
    scripts/gen_golden.py).  Outside every timed region; the golden holds the 8x8-pooled map and the per-image sums."""
    torch = env.torch
    import numpy as np
    import torch.nn.functional as F
    from oracle import synth            # seeded input generator only (the checker's inputs), no oracle arithmetic
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "base_b8.npz"))
    except Exception:
        return None
    imgs, boxes = synth.make_inputs(8, seed=1234)
    was_training = model.training
    model.eval()
    with torch.no_grad():
        out = model(imgs.to(env.dev), boxes.to(env.dev), 3).float().cpu()
    model.train(was_training)
    pool = F.avg_pool2d(out[:, None], 8)[:, 0].double().numpy()
    ref = g["out_pool8"].astype(np.float64)
    sums = out.sum((1, 2)).double().numpy()
    return {"map_rel_l2": float(np.linalg.norm(pool - ref) / np.linalg.norm(ref)),
            "count_rel": float(np.max(np.abs(sums - g["out_sum"]) / np.abs(g["out_sum"]))),
            "against": "reference output on the same seeded weights / inputs (tests/golden/base_b8.npz, 8x8-pooled map, B=8, 3 shots)",
            "tolerance": 1e-3}


def eager_b200_leg(env):
    """The same fine-tune step through stock PyTorch on this B200 (cuBLAS / cuDNN / ATen, fp16 autocast, fused AdamW): the
    'existing Blackwell software stack' the reference would run on.  A reported baseline like cpu_baseline."""
    torch = env.torch
    from oracle import countr_oracle as O
    from oracle import synth
    dev = env.dev
    cfg = synth.CONFIGS["base"]
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(cfg, 0).items()}
    names = O.decoder_param_names(sd, SHOTS)
    params = []
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
        params.append(sd[n])
    opt = torch.optim.AdamW(params, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    B = PER_GPU_BATCH["finetune"]
    imgs, boxes = synth.make_inputs(B, seed=1)
    gt, mask = synth.make_targets(B, seed=2)
    imgs, boxes, gt, mask = imgs.to(dev), boxes.to(dev), gt.to(dev), mask.to(dev)
    scale = 4096.0

    def step():
        with torch.autocast("cuda", dtype=torch.float16):
            out = O.forward(sd, cfg, imgs, boxes, SHOTS)
        loss = O.finetune_loss(out.float(), gt, mask)
        opt.zero_grad(set_to_none=True)
        (loss * scale).backward()
        torch._foreach_mul_([p.grad for p in params], 1.0 / scale)
        opt.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    del sd, params, opt
    torch.cuda.empty_cache()
    return {"value": round(B / ms * 1e3, 2), "unit": "images/s", "ms_per_step": round(ms, 3),
            "what": "reference algorithm (oracle port) in stock PyTorch eager on the same B200: fp16 autocast, cuBLAS/cuDNN/ATen, fused AdamW, batch 8"}


def script_mode_child():
    """Child process of the script_mode leg: the UNMODIFIED script's conditions around our model — CUDA_LAUNCH_BLOCKING=1
    (FSC_finetune_cross.py:110), eager launches (no CUDA graph), fp16 inputs under autocast, GradScaler, torch.optim.AdamW,
    host-side numpy mask + H2D and loss.item() every step (:273-316).  Wall-clock timed (every launch is synchronous)."""
    import numpy as np
    import torch
    import models_mae_cross
    assert os.environ.get("CUDA_LAUNCH_BLOCKING") == "1"
    dev = torch.device("cuda:0")
    B = PER_GPU_BATCH["finetune"]
    torch.manual_seed(0)
    model = models_mae_cross.mae_vit_base_patch16(norm_pix_loss=False).to(dev).train()
    opt = torch.optim.AdamW(param_groups(model, 0.05), lr=1e-5, betas=(0.9, 0.95))
    scaler = torch.cuda.amp.GradScaler()
    g = torch.Generator().manual_seed(1)
    samples = torch.rand(B, 3, 384, 384, generator=g)
    boxes = torch.rand(B, 3, 3, 64, 64, generator=g)
    gt = torch.rand(B, 384, 384, generator=g) * 0.5
    np.random.seed(0)
    steps, warm = 12, 4
    t0 = None
    for it in range(warm + steps):
        if it == warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        s = samples.to(dev, non_blocking=True, dtype=torch.half)
        gd = gt.to(dev, non_blocking=True, dtype=torch.half)
        bx = boxes.to(dev, non_blocking=True, dtype=torch.half)
        with torch.cuda.amp.autocast():
            output = model(s, bx, SHOTS)
        mask = np.random.binomial(n=1, p=0.8, size=[384, 384])
        masks = np.tile(mask, (output.shape[0], 1)).reshape(output.shape[0], 384, 384)
        masks = torch.from_numpy(masks).to(dev)
        loss = (output - gd) ** 2
        loss = (loss * masks / (384 * 384)).sum() / output.shape[0]
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        loss.item()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    print(json.dumps({"value": round(B / sec, 2), "unit": "images/s", "ms_per_step": round(sec * 1e3, 3), "steps": steps,
                      "what": "script conditions: CUDA_LAUNCH_BLOCKING=1, eager launches, autocast fp16, GradScaler, torch AdamW, "
                              "numpy mask + H2D + loss.item() per step (wall clock)"}))


def script_mode_leg():
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--script-mode-child"], env=env, capture_output=True, text=True,
                           timeout=240)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (r.stderr or r.stdout)[-300:]}
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_finetune(env, args):
    torch, dist = env.torch, env.dist
    import models_mae_cross
    from countr_b200 import ops
    from countr_b200.engine import engine
    from countr_b200.train import FineTuner
    dev, world, rank = env.dev, env.world, env.rank
    B = PER_GPU_BATCH["finetune"]
    torch.manual_seed(0)
    model = models_mae_cross.mae_vit_base_patch16(norm_pix_loss=False)
    parity = None
    try:
        from oracle import synth          # seeded weight generator: the parity check below needs the weights of the golden
        model.load_state_dict(synth.make_state_dict(synth.CONFIGS["base"], seed=0), strict=True)
        model = model.to(dev).train()
        if rank == 0 and not args.no_extras:
            parity = finetune_parity(env, model)
    except Exception as e:  # pragma: no cover
        print(f"[bench] parity check unavailable: {type(e).__name__}: {e}", file=sys.stderr)
        model = model.to(dev).train()
    eng = engine()
    eng.grad_allreduce = None      # the bench issues the all-reduce itself
    loss_scale = 4096.0
    use_tuner = not args.script_loop

    # host inputs (pinned), a few distinct batches rotated over the steps
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host = []
    for _ in range(n_host):
        host.append(dict(
            imgs=torch.rand(B, 3, 384, 384, generator=g).pin_memory(),
            boxes=torch.rand(B, SHOTS, 3, 64, 64, generator=g).pin_memory(),
            gt=(torch.rand(B, 384, 384, generator=g) * 0.5).pin_memory(),
            mask=(torch.rand(384, 384, generator=g) < 0.8).float().pin_memory()))     # np.random.binomial(1,.8) stand-in
    d = dict(imgs=torch.empty(B, 3, 384, 384, device=dev), boxes=torch.empty(B, SHOTS, 3, 64, 64, device=dev),
             gt=torch.empty(B, 384, 384, device=dev), mask=torch.empty(384, 384, device=dev))
    d_loss = torch.zeros((), device=dev)
    h_loss = torch.zeros((), pin_memory=True)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0].values())
    for k in d:
        d[k].copy_(host[0][k], non_blocking=True)
    torch.cuda.synchronize()

    comm = torch.cuda.Stream() if world > 1 else None
    pipelined = use_tuner and world > 1 and not args.no_overlap
    if use_tuner:
        # countr_b200.train.FineTuner: the reference loop's arithmetic with every kernel ours (fused loss + counts, inf check +
        # grad norm, flat-arena AdamW, GradScaler update; device-resident lr / scale)
        tuner = FineTuner(model, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), loss_scale=loss_scale)
        opt = None
        lat_cur = None

        def fwd_bwd():
            d_loss.copy_(tuner.forward_backward(d["imgs"], d["boxes"], d["gt"], d["mask"], SHOTS))

        def update():
            tuner.update()

        def arena():
            return tuner.arena
    else:
        # the unmodified script's loop: autograd through SupervisedMAE + torch.optim.AdamW (FSC_finetune_cross.py:286-316)
        tuner = None
        opt = torch.optim.AdamW(param_groups(model, 0.05), lr=1e-5, betas=(0.9, 0.95), fused=True, capturable=True)

        def fwd_bwd():
            out = model(d["imgs"], d["boxes"], SHOTS)
            loss = ((out - d["gt"]) ** 2 * d["mask"] / (384 * 384)).sum() / B
            (loss * loss_scale).backward()
            d_loss.copy_(loss.detach())

        def update():
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            torch._foreach_mul_(grads, 1.0 / loss_scale)
            opt.step()

        def arena():
            return eng.last_arena

    def zero_grad():
        if opt is not None:
            opt.zero_grad(set_to_none=True)

    def allreduce(stream=None):
        if world == 1:
            return
        a = arena()
        if opt is not None:      # script loop: autograd may have cloned the views (then reduce the clones: slower, still exact)
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            base = a.untyped_storage().data_ptr()
            if not all(gr.untyped_storage().data_ptr() == base for gr in grads):
                for gr in grads:
                    dist.all_reduce(gr, op=dist.ReduceOp.AVG)
                return
        if stream is None:
            dist.all_reduce(a, op=dist.ReduceOp.AVG)       # ONE collective over the flat 53 MB arena
        else:
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                dist.all_reduce(a, op=dist.ReduceOp.AVG)

    def step_eager():
        zero_grad()
        fwd_bwd()
        allreduce()
        update()

    eager_warmup(torch, step_eager, 3)

    # ---- CUDA graphs.  N == 1: the whole step is one graph.  N > 1: the NCCL all-reduce is not captured; the step is
    # [decoder fwd/bwd graph] -> all-reduce on a comm stream, OVERLAPPED with [encoder graph of the NEXT batch] (the frozen
    # encoder depends on no trainable parameter) -> [update graph].  Every timed step runs one encoder, one decoder
    # forward/backward, one all-reduce and one update; the encoder of the very first batch is the (untimed) pipeline fill.
    graphs, mode = None, "eager"
    launches_per_step = None
    if not args.no_graph:
        try:
            zero_grad()
            n0 = ops.LAUNCHES[0]
            if world == 1:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    fwd_bwd()
                    update()
                graphs, mode = (g1,), "one graph per step"
            elif pipelined:
                M = B * 576
                lat_ws = tuner.encode(d["imgs"])                     # workspace buffer the encoder graph writes
                lat_cur = torch.empty_like(lat_ws)
                torch.cuda.synchronize()
                g_enc, g_dec, g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_enc):
                    tuner.encode(d["imgs"])
                with torch.cuda.graph(g_dec, pool=g_enc.pool()):
                    lat_cur.copy_(lat_ws)
                    d_loss.copy_(tuner.forward_backward(d["imgs"], d["boxes"], d["gt"], d["mask"], SHOTS, lat16=lat_cur))
                with torch.cuda.graph(g_upd, pool=g_enc.pool()):
                    update()
                graphs, mode = (g_enc, g_dec, g_upd), "decoder graph | all-reduce overlapped with the next batch's encoder graph | update graph"
            else:
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    fwd_bwd()
                with torch.cuda.graph(g2, pool=g1.pool()):
                    update()
                graphs, mode = (g1, g2), "fwd/bwd graph | all-reduce | update graph"
            launches_per_step = ops.LAUNCHES[0] - n0
        except Exception as e:  # pragma: no cover
            print(f"[bench] rank {rank}: CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()
    if not env.agree(graphs is not None):
        graphs, mode, pipelined = None, "eager", False
    if graphs is None:
        pipelined = False

    def run_step(next_imgs=None):
        """One step on the resident inputs; pipelined mode: `next_imgs` (a callable moving the next batch's images into
        d['imgs']) runs right before the next batch's encoder."""
        if graphs is None:
            step_eager()
        elif len(graphs) == 1:
            graphs[0].replay()
        elif pipelined:
            graphs[1].replay()                    # decoder forward / backward of this batch (reads lat_ws of this batch)
            allreduce(comm)                       # in flight on the comm stream ...
            if next_imgs is not None:
                next_imgs()
            graphs[0].replay()                    # ... while the encoder of the next batch runs
            torch.cuda.current_stream().wait_stream(comm)
            graphs[2].replay()
        else:
            graphs[0].replay()
            allreduce()
            graphs[1].replay()

    if launches_per_step is None:
        n0 = ops.LAUNCHES[0]
        run_step()
        launches_per_step = ops.LAUNCHES[0] - n0
    if pipelined:
        graphs[0].replay()                        # pipeline fill: encoder of the first batch
    for _ in range(env.warmup):
        run_step()
    sampler = ClockSampler(env.local)
    sampler.start()
    # (1) inputs resident in HBM
    ms_dev = env.timed(lambda i: run_step(), env.steps)

    # (2) end to end: H2D of every step's inputs (pinned host -> device) + D2H of the loss, every step, all inside the timed
    # region.  The loop is the usual prefetching input pipeline: while step i computes, the copy stream uploads the inputs of
    # step i+1 into a staging set; a step starts with a device-to-device move of the staged inputs into the buffers the CUDA
    # graph reads (20 MB, ~6 us).  Exactly `steps` uploads happen in the timed region; the first one is not hidden.
    copy_stream = torch.cuda.Stream()
    staging = [{k: torch.empty_like(v) for k, v in d.items()} for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    n_e2e = env.steps

    def stage_upload(i):
        hb, st = host[i % n_host], staging[i % 2]
        with torch.cuda.stream(copy_stream):
            for k in ("imgs", "boxes", "gt", "mask"):
                st[k].copy_(hb[k], non_blocking=True)
            ev_up[i % 2].record(copy_stream)

    def e2e_step(i):
        main = torch.cuda.current_stream()
        st = staging[i % 2]
        if i == 0:
            stage_upload(0)
        main.wait_event(ev_up[i % 2])
        if i == 0 and pipelined:                   # pipeline fill inside the timed region: encoder of batch 0
            d["imgs"].copy_(st["imgs"], non_blocking=True)
            graphs[0].replay()
        keys = ("boxes", "gt", "mask") if pipelined else ("imgs", "boxes", "gt", "mask")
        for k in keys:
            d[k].copy_(st[k], non_blocking=True)
        ev_free[i % 2].record(main)
        if i + 1 < n_e2e:
            copy_stream.wait_event(ev_free[(i + 1) % 2])   # (recorded two steps ago) the set being overwritten was consumed
            stage_upload(i + 1)
        if pipelined:
            def next_imgs():
                if i + 1 < n_e2e:
                    main.wait_event(ev_up[(i + 1) % 2])
                    d["imgs"].copy_(staging[(i + 1) % 2]["imgs"], non_blocking=True)
            run_step(next_imgs)
        else:
            run_step()
        h_loss.copy_(d_loss, non_blocking=True)
        main.synchronize()     # the script reads loss.item() every step (FSC_finetune_cross.py:306)
    ms_e2e = env.timed(e2e_step, n_e2e)
    clocks = sampler.finish()
    final_loss = float(h_loss)
    tuner_metrics = tuner.metrics() if tuner is not None else None
    if pipelined:
        graphs[0].replay()                        # restore the pipeline invariant (d['imgs'] encoder output is current)

    sustained = None
    if not args.no_extras:
        sustained = sustained_window(env, run_step, ms_dev / env.steps, B)
    roof = dominant_kernel_roofline(env, B)

    if rank != 0:
        env.done()
        return
    imgs_per_step = B * world
    value = imgs_per_step * env.steps / (ms_dev * 1e-3)
    e2e = imgs_per_step * n_e2e / (ms_e2e * 1e-3)
    roof["whole_step"] = whole_step(value, world, "finetune", ms_dev, sustained["value"] if sustained else None)
    line = {
        "metric": METRIC["finetune"], "value": round(value, 2), "unit": "images/s", "n_gpus": world,
        "steps": env.steps, "warmup": env.warmup, "ms_per_step": round(ms_dev / env.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME["finetune"],
                   "global_batch": imgs_per_step,
                   "step": "encoder fwd (frozen) + decoder fwd/bwd + masked-MSE loss + counts + inf check / grad norm / unscale + AdamW + loss-scale update"
                   + (" + NCCL grad all-reduce (avg)" if world > 1 else ""),
                   "cuda_graph": graphs is not None, "schedule": mode,
                   "step_api": "countr_b200.train.FineTuner" if use_tuner else "script loop (autograd + torch.optim.AdamW)",
                   "loss_scale": loss_scale, "sm_budget": env.sm_budget,
                   "l2": "per-step working set (~1.5 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / n_e2e, 4), "h2d": "prefetched on a copy stream during the previous step"},
        "gpu_launches": launches_per_step * env.steps,
        "clocks": clocks,
        "roofline": roof,
        "final_loss": final_loss,
    }
    if parity is not None:
        line["parity"] = parity
    if tuner_metrics is not None:
        line["step_metrics"] = tuner_metrics
    if sustained is not None:
        line["sustained"] = sustained
    if world == 1 and not args.no_extras:
        line["eager_b200"] = eager_b200_leg(env)
        line["script_mode"] = script_mode_leg()
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, sec, sample = cpu_steps("finetune", 2, 2, 1)
        line["cpu_baseline"] = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    env.done()


# ------------------------------------------------------------------------------------------------ zero-shot inference
def run_infer0(env, args):
    torch = env.torch
    import models_mae_cross
    from countr_b200 import ops
    dev, world, rank = env.dev, env.world, env.rank
    B = PER_GPU_BATCH["infer0"]
    torch.manual_seed(0)
    model = models_mae_cross.mae_vit_base_patch16(norm_pix_loss=False).to(dev).eval()
    g = torch.Generator().manual_seed(77 + rank)
    host = [torch.rand(B, 3, 384, 384, generator=g).pin_memory() for _ in range(2)]
    d_imgs = torch.empty(B, 3, 384, 384, device=dev)
    d_imgs.copy_(host[0])
    empty = torch.empty(B, 0, device=dev)
    d_counts = torch.zeros(B, device=dev)
    h_counts = torch.zeros(B, pin_memory=True)

    def step():
        with torch.no_grad():
            out = model(d_imgs, empty, 0)
            torch.sum(out.view(B, -1), dim=1, out=d_counts)      # the script's count: density.sum() / 60 (demo_zero.py:51)

    eager_warmup(torch, step, 3)
    graph = None
    n0 = ops.LAUNCHES[0]
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
        except Exception as e:  # pragma: no cover
            print(f"[bench] graph capture failed ({type(e).__name__}: {e})", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if not env.agree(graph is not None):
        graph = None
    run_step = graph.replay if graph is not None else step
    if graph is None:
        n0 = ops.LAUNCHES[0]
        step()
    launches = ops.LAUNCHES[0] - n0
    for _ in range(env.warmup):
        run_step()
    sampler = ClockSampler(env.local)
    sampler.start()
    ms_dev = env.timed(lambda i: run_step(), env.steps)
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(d_imgs) for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    n = env.steps

    def upload(i):
        with torch.cuda.stream(copy_stream):
            staging[i % 2].copy_(host[i % 2], non_blocking=True)
            ev_up[i % 2].record(copy_stream)

    def e2e_step(i):
        main = torch.cuda.current_stream()
        if i == 0:
            upload(0)
        main.wait_event(ev_up[i % 2])
        d_imgs.copy_(staging[i % 2], non_blocking=True)
        ev_free[i % 2].record(main)
        if i + 1 < n:
            copy_stream.wait_event(ev_free[(i + 1) % 2])
            upload(i + 1)
        run_step()
        h_counts.copy_(d_counts, non_blocking=True)
        main.synchronize()
    ms_e2e = env.timed(e2e_step, n)
    clocks = sampler.finish()
    sustained = None if args.no_extras else sustained_window(env, run_step, ms_dev / env.steps, B)
    roof = dominant_kernel_roofline(env, B)
    if rank != 0:
        env.done()
        return
    value = B * world * env.steps / (ms_dev * 1e-3)
    e2e = B * world * n / (ms_e2e * 1e-3)
    roof["whole_step"] = whole_step(value, world, "infer0", ms_dev, sustained["value"] if sustained else None)
    line = {"metric": METRIC["infer0"], "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": env.steps,
            "warmup": env.warmup, "ms_per_step": round(ms_dev / env.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME["infer0"], "global_batch": B * world, "step": "SupervisedMAE.forward(imgs, empty boxes, 0) "
                       "under no_grad + per-image count", "cuda_graph": graph is not None, "parallelism": "replicas only (no collective)",
                       "l2": "activations of one step (> 2 GB) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": d_imgs.numel() * 4, "d2h_bytes_per_step": 4 * B,
                    "ms_per_step": round(ms_e2e / n, 4), "h2d": "fp32 images prefetched on a copy stream during the previous step"},
            "gpu_launches": launches * env.steps, "clocks": clocks, "roofline": roof}
    if sustained is not None:
        line["sustained"] = sustained
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, sec, sample = cpu_steps("infer0", 4, 2, 1)
        line["cpu_baseline"] = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    env.done()


# ------------------------------------------------------------------------------------------------ MAE pre-training
def run_pretrain(env, args):
    torch, dist = env.torch, env.dist
    import models_mae_noct
    from countr_b200 import ops
    from countr_b200.engine import engine
    dev, world, rank = env.dev, env.world, env.rank
    B = PER_GPU_BATCH["pretrain"]
    torch.manual_seed(0)
    model = models_mae_noct.mae_vit_base_patch16(norm_pix_loss=True).to(dev).train()     # FSC_pretrain.py --norm_pix_loss
    from countr_b200.train import ArenaAdamW
    scale = 1024.0
    names, params = model._trainable()
    # unscale + inf check + grad norm + AdamW (timm add_weight_decay grouping) + GradScaler.update as three kernels over the flat
    # gradient arena the backward leaves behind (util/misc.py:260-301 / FSC_pretrain.py:226-228, 296-300)
    opt = ArenaAdamW(names, params, lr=1e-5, weight_decay=0.05, betas=(0.9, 0.95), loss_scale=scale, dynamic_scale=False)
    eng = engine()
    eng.grad_allreduce = None
    g = torch.Generator().manual_seed(99 + rank)
    host = [torch.rand(B, 3, 384, 384, generator=g).pin_memory() for _ in range(2)]
    d_imgs = torch.empty(B, 3, 384, 384, device=dev)
    d_imgs.copy_(host[0])
    d_loss = torch.zeros((), device=dev)
    h_loss = torch.zeros((), pin_memory=True)

    def fwd_bwd():
        loss, _, _ = model(d_imgs, mask_ratio=0.5)
        (loss * scale).backward()
        d_loss.copy_(loss.detach())

    def zero_grad():
        for p in params:
            p.grad = None

    def update():
        opt.step(eng.last_arena)

    def allreduce():
        if world > 1:
            dist.all_reduce(eng.last_arena, op=dist.ReduceOp.AVG)          # ONE collective over the flat 446.6 MB arena

    # N > 1, overlapped: the backward hands the arena to `slice_hook` in two pieces — [decoder_embed .. end] as soon as the decoder
    # backward is done, the encoder part at the end — and each piece is all-reduced on the comm stream while the main stream
    # keeps computing; the update waits for both.  The whole step (NCCL kernels included) is captured in ONE CUDA graph.
    overlapped = world > 1 and args.overlap_pretrain
    comm = torch.cuda.Stream() if world > 1 else None

    def slice_hook(sl):
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            dist.all_reduce(sl, op=dist.ReduceOp.AVG)

    def step_overlapped():
        eng.grad_slice_hook = slice_hook
        try:
            fwd_bwd()
        finally:
            eng.grad_slice_hook = None
        torch.cuda.current_stream().wait_stream(comm)
        update()

    def step_eager():
        zero_grad()
        if overlapped:
            step_overlapped()
            return
        fwd_bwd()
        allreduce()
        update()

    eager_warmup(torch, step_eager, 3)
    graphs = None
    n0 = ops.LAUNCHES[0]
    if not args.no_graph:
        try:
            zero_grad()
            if world == 1:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    fwd_bwd()
                    update()
                graphs = (g1,)
            elif overlapped:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    step_overlapped()
                graphs = (g1,)
            else:
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    fwd_bwd()
                with torch.cuda.graph(g2, pool=g1.pool()):
                    update()
                graphs = (g1, g2)
        except Exception as e:  # pragma: no cover
            print(f"[bench] rank {rank}: graph capture failed ({type(e).__name__}: {e})", file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()
    if not env.agree(graphs is not None):
        graphs = None

    def run_step():
        if graphs is None:
            step_eager()
        elif len(graphs) == 1:
            graphs[0].replay()
        else:
            graphs[0].replay()
            allreduce()
            graphs[1].replay()

    if graphs is None:
        n0 = ops.LAUNCHES[0]
        run_step()
    launches = ops.LAUNCHES[0] - n0
    for _ in range(env.warmup):
        run_step()
    sampler = ClockSampler(env.local)
    sampler.start()
    ms_dev = env.timed(lambda i: run_step(), env.steps)
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(d_imgs) for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    n = env.steps

    def upload(i):
        with torch.cuda.stream(copy_stream):
            staging[i % 2].copy_(host[i % 2], non_blocking=True)
            ev_up[i % 2].record(copy_stream)

    def e2e_step(i):
        main = torch.cuda.current_stream()
        if i == 0:
            upload(0)
        main.wait_event(ev_up[i % 2])
        d_imgs.copy_(staging[i % 2], non_blocking=True)
        ev_free[i % 2].record(main)
        if i + 1 < n:
            copy_stream.wait_event(ev_free[(i + 1) % 2])
            upload(i + 1)
        run_step()
        h_loss.copy_(d_loss, non_blocking=True)
        main.synchronize()
    ms_e2e = env.timed(e2e_step, n)
    clocks = sampler.finish()
    sustained = None if args.no_extras else sustained_window(env, run_step, ms_dev / env.steps, B)
    roof = None
    if rank != 0:
        env.done()
        return
    value = B * world * env.steps / (ms_dev * 1e-3)
    e2e = B * world * n / (ms_e2e * 1e-3)
    peaks, peak_src = load_peaks()
    roof = {"bound": "tensor", "kernel": "whole step (every contraction is a tcgen05 GEMM / attention kernel)", "unit": "TFLOP/s",
            "achieved": round(value / world * GFLOP_PER_IMG["pretrain"] / 1e3, 1), "peak": peaks.get("bf16_tflops"),
            "frac": round(value / world * GFLOP_PER_IMG["pretrain"] / 1e3 / peaks.get("bf16_tflops"), 4), "peak_source": peak_src, "traffic": None,
            "whole_step": whole_step(value, world, "pretrain", ms_dev, sustained["value"] if sustained else None)}
    line = {"metric": METRIC["pretrain"], "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": env.steps,
            "warmup": env.warmup, "ms_per_step": round(ms_dev / env.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME["pretrain"], "global_batch": B * world,
                       "step": "forward + full backward (encoder trained) + inf check / grad norm / unscale + AdamW (countr_b200.train.ArenaAdamW)" +
                               (" + NCCL all-reduce (avg) of the flat 446.6 MB gradient arena" if world > 1 else "") +
                               (" in two pieces overlapped with the backward" if overlapped else ""),
                       "cuda_graph": graphs is not None, "loss_scale": scale, "sm_budget": env.sm_budget,
                       "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": d_imgs.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / n, 4), "h2d": "fp32 images prefetched on a copy stream during the previous step"},
            "gpu_launches": launches * env.steps, "clocks": clocks, "roofline": roof, "final_loss": float(h_loss)}
    if sustained is not None:
        line["sustained"] = sustained
    if world == 1 and not args.no_cpu_baseline:
        ips, cores, sec, sample = cpu_steps("pretrain", 2, 1, 1)
        line["cpu_baseline"] = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    env.done()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="finetune", choices=["finetune", "infer0", "pretrain"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / sustained-window / eager_b200 / script_mode legs")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: do not overlap the all-reduce with the next batch's encoder")
    ap.add_argument("--overlap-pretrain", action="store_true",
                    help="N > 1, pretrain: all-reduce the gradient arena in slices during the backward (experimental; slower at N = 2)")
    ap.add_argument("--script-loop", action="store_true",
                    help="drive the fine-tune step with the reference script's own loop (model() -> loss.backward() -> torch.optim.AdamW) "
                         "instead of countr_b200.train.FineTuner; both run the same forward / backward kernels")
    ap.add_argument("--fused-step", action="store_true", help="(default since round 2; kept for compatibility)")
    ap.add_argument("--script-mode-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.script_mode_child:
        return script_mode_child()
    if args.impl == "reference":
        return run_reference(args)
    env = Env(args)
    {"finetune": run_finetune, "infer0": run_infer0, "pretrain": run_pretrain}[args.workload](env, args)


if __name__ == "__main__":
    main()
