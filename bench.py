#!/usr/bin/env python
"""bench.py -- DETR-R50 800x1333 train-step throughput (images/sec), BASELINE.json configs[1] (B=8 per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3                  # our arm (sm_100a kernels)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps K --warmup W   # CPU restatement of the reference (oracle)

One JSON line on rank 0.  `value`: whole-job images/sec with the batch already resident in HBM (the full train step
-- forward with dropout, on-device Hungarian matching + set loss, backward, gradient all-reduce, Adam with per-variable
clipnorm, bf16 weight refresh -- replayed as one CUDA graph per rank).  `e2e`: the same step through the public API
(training.run_train_step + aggregate_grad_and_apply, i.e. the body of training.fit) with HOST (pinned) inputs copied
in and the loss read back every step.  TensorFlow is not installable here, so the reference arm times the PyTorch-CPU
oracle restatement (oracle/detr_oracle.py) on the host cores, labelled as such.
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

# train-step GFLOP per 800x1333 image = 3 x forward - 5.0 (no data gradient into the image)
TRAIN_GFLOP_PER_IMAGE = {"resnet50": 604.9, "resnet101": 1082.2}
METRIC = "images/sec DETR-R50 800x1333 train step"


def measured_peaks():
    """(sustained bf16 TF/s, burst bf16 TF/s, HBM GB/s, source).  MEASURED_PEAKS.json is driver-written; B200_PROFILING.md's
    fallback figures otherwise."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def tensor_peak(clocks):
    """Which measured tensor peak applies: a kernel timed while the SM clock sits at its maximum with no power cap active runs
    in the burst regime (MEASURED_PEAKS bf16_tflops); under the power cap / at reduced clocks the sustained figure applies."""
    sus, burst, _, how = measured_peaks()
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"] \
            and "sw_power_cap" not in (clocks.get("reasons") or []):
        return burst, f"MEASURED_PEAKS.json bf16_tflops (burst: SM clock {clocks['sm_mhz']:.0f} of {clocks['sm_max_mhz']:.0f} MHz, no power cap; {how})"
    return sus, f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(B, H, W, seed, n=20):
    """images ~ N(0,1); n=20 targets/image, cx,cy~U(.1,.9), w,h~U(.02,.5), class~randint(0,91)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, H, W, 3, generator=g)
    tb = torch.zeros(B, 100, 4)
    tc = torch.zeros(B, 100, 1, dtype=torch.int64)
    for b in range(B):
        tb[b, 0, 0] = n
        tb[b, 1:1 + n, :2] = torch.rand(n, 2, generator=g) * 0.8 + 0.1
        tb[b, 1:1 + n, 2:] = torch.rand(n, 2, generator=g) * 0.48 + 0.02
        tc[b, 1:1 + n, 0] = torch.randint(0, 91, (n,), generator=g)
    return images, tb, tc


def cpu_oracle_step_time(H, W, threads, steps=1, warmup=0):
    """seconds per image of the oracle's train step (fwd + matcher + set loss + bwd + Adam) on `threads` host cores."""
    import torch
    from oracle import detr_oracle as O
    torch.set_num_threads(threads)
    P = O.init_params(seed=0)
    images, tb, tc = synthetic_batch(1, H, W, seed=0)
    state = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in P.items() if O.param_group(n)}

    def one(step):
        _, total, _, grads = O.train_step(P, images, tb, tc, background_class=91, training=True)
        for n, g in grads.items():
            if g is not None:
                lr = 1e-5 if O.param_group(n) == "backbone" else 1e-4
                O.adam_clipnorm_step(P[n], g, state[n][0], state[n][1], step, lr, 0.1)
        return float(total)
    for i in range(warmup):
        one(i + 1)
    t0 = time.time()
    for i in range(steps):
        one(warmup + i + 1)
    return (time.time() - t0) / steps


def c5_inputs(B=256, Q=100, C=92, n=20, seed=1234, layers=1):
    """SURVEY 8(d) C5 (BASELINE configs[4]): logits ~ N(0,1) [B,100,92]; pred boxes cx,cy~U(.05,.95), w,h~U(.02,.5);
    n=20 targets per image; seed 1234.  `layers` stacks independent decoder-layer outputs ([L,B,Q,*])."""
    import torch
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(layers, B, Q, C, generator=g)
    boxes = torch.cat([torch.rand(layers, B, Q, 2, generator=g) * 0.9 + 0.05,
                       torch.rand(layers, B, Q, 2, generator=g) * 0.48 + 0.02], -1)
    tb, tc = synthetic_targets(B, seed, n)
    return logits, boxes, tb, tc


def synthetic_targets(B, seed, n=20):
    """T0 wire format (data/processing.py:35-55): row 0 = [n,0,0,0], rows 1..n = (cx,cy,w,h); classes in t_class rows 1..n"""
    import torch
    g = torch.Generator().manual_seed(seed + 7)
    tb = torch.zeros(B, 100, 4)
    tc = torch.zeros(B, 100, 1, dtype=torch.int64)
    for b in range(B):
        tb[b, 0, 0] = n
        tb[b, 1:1 + n, :2] = torch.rand(n, 2, generator=g) * 0.8 + 0.1
        tb[b, 1:1 + n, 2:] = torch.rand(n, 2, generator=g) * 0.48 + 0.02
        tc[b, 1:1 + n, 0] = torch.randint(0, 91, (n,), generator=g)
    return tb, tc


def matcher_cpu_us_per_image(logits, boxes, tb, tc, max_images=256):
    """The reference's matcher arithmetic on ONE host core, the way the reference runs it (python loop over images under the
    GIL, hungarian_matching.py:163-203): numpy cost build + the real scipy.optimize.linear_sum_assignment."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    lg, bx = logits[0].numpy(), boxes[0].numpy()
    tbn, tcn = tb.numpy(), tc.numpy()
    nimg = min(max_images, lg.shape[0])

    def xyxy(b):
        return np.clip(np.concatenate([b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2], -1), 0.0, 1.0)
    t0 = time.perf_counter()
    for i in range(nimg):
        n = int(tbn[i, 0, 0])
        t_b, t_c = tbn[i, 1:1 + n], tcn[i, 1:1 + n, 0]
        e = np.exp(lg[i] - lg[i].max(-1, keepdims=True))
        prob = e / e.sum(-1, keepdims=True)
        cost_class = -prob[:, t_c]
        cost_bbox = np.abs(bx[i][:, None, :] - t_b[None, :, :]).sum(-1)
        p, t = xyxy(bx[i]), xyxy(t_b)
        area_p, area_t = (p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1]), (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
        lt, rb = np.maximum(p[:, None, :2], t[None, :, :2]), np.minimum(p[:, None, 2:], t[None, :, 2:])
        wh = np.clip(rb - lt, 0, None)
        inter = wh[..., 0] * wh[..., 1]
        union = area_p[:, None] + area_t[None, :] - inter
        iou = inter / union
        elt, erb = np.minimum(p[:, None, :2], t[None, :, :2]), np.maximum(p[:, None, 2:], t[None, :, 2:])
        ewh = np.clip(erb - elt, 0, None)
        earea = ewh[..., 0] * ewh[..., 1]
        giou = iou - (earea - union) / earea
        cost = 5.0 * cost_bbox + 1.0 * cost_class + 2.0 * (-giou)
        linear_sum_assignment(cost.astype(np.float32))
    return (time.perf_counter() - t0) / nimg * 1e6


def matcher_microbench(D, iters=20):
    """BASELINE metric part 2, 'matcher us/image' (configs[4]): 100 queries x 20 targets x batch 256.
    (a) cost build + exact assignment of one decoder layer (256 problems, one launch of matcher_kernel);
    (b) the whole set loss through the public API get_losses on 6 layers (1536 problems: matcher + loss kernels);
    (c) the same with HOST inputs (h2d copies inside);  (d) the reference arithmetic on one host core."""
    import torch
    from detr_tensorflow_b200 import ops
    B, Q, C = 256, 100, 92
    logits, boxes, tb, tc = c5_inputs(B, Q, C, 20, 1234, layers=6)
    dl, db, dtb, dtc = (x.cuda() for x in (logits, boxes, tb, tc))
    out = dict(p=torch.empty(B, Q, dtype=torch.int64, device="cuda"), t=torch.empty(B, Q, dtype=torch.int64, device="cuda"),
               s=torch.empty(B, Q, dtype=torch.uint8, device="cuda"), m=torch.empty(B, Q, dtype=torch.int32, device="cuda"),
               st=torch.empty(B, dtype=torch.int32, device="cuda"))
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def one_layer():
        ops.matcher(dl[0], C, db[0], dtb, dtc, B, B, Q, C, out["p"], out["t"], out["s"], out["m"], None, out["st"])
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    m_dev = {"pred_logits": dl[5], "pred_boxes": db[5], "aux": [{"pred_logits": dl[i], "pred_boxes": db[i]} for i in range(5)]}
    m_host = {"pred_logits": logits[5].pin_memory(), "pred_boxes": boxes[5].pin_memory(),
              "aux": [{"pred_logits": logits[i].pin_memory(), "pred_boxes": boxes[i].pin_memory()} for i in range(5)]}
    tbh, tch = tb.pin_memory(), tc.pin_memory()

    def timed(fn, sync_result=False):
        ts = []
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            r = fn()
            e1.record()
            if sync_result:
                float(r[0])
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if i >= 3:
                ts.append((t1 - t0) * 1e3 if sync_result else e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    ms_a = timed(one_layer)
    assert int(out["st"].abs().sum()) == 0 and int((out["m"] >= 0).sum()) == B * 20
    ms_b = timed(lambda: D.get_losses(m_dev, dtb, dtc, cfg))
    ms_c = timed(lambda: D.get_losses({"pred_logits": m_host["pred_logits"].cuda(non_blocking=True),
                                       "pred_boxes": m_host["pred_boxes"].cuda(non_blocking=True),
                                       "aux": [{k: v.cuda(non_blocking=True) for k, v in a.items()} for a in m_host["aux"]]},
                                      tbh, tch, cfg), sync_result=True)
    cpu_us = matcher_cpu_us_per_image(logits, boxes, tb, tc)
    alg_bytes = Q * C * 4 + Q * 4 * 4 + 20 * 24 + 420                              # ~39.3 KB / problem
    _, _, peak_hbm, _ = measured_peaks()
    return {"workload": "BASELINE configs[4]: 100 queries x 20 targets x batch 256, seed 1234; L2 flushed between iterations",
            "us_per_image": ms_a * 1e3 / B, "problems": B, "kernel": "matcher_kernel (cost build + shortest-augmenting-path LSAP, 1 CTA/problem)",
            "achieved_gbs": alg_bytes * B / (ms_a * 1e-3) / 1e9, "peak_gbs": peak_hbm, "bound": "latency (sequential augmentations)",
            "set_loss_6layers_us_per_image": ms_b * 1e3 / B, "set_loss_6layers_problems": 6 * B,
            "e2e_us_per_image": ms_c * 1e3 / B,
            "e2e_path": "get_losses(m_outputs, t_bbox, t_class, config) with pinned HOST inputs (36.9 MB h2d) and the total loss read back",
            "cpu_us_per_image": cpu_us, "cpu_cores": 1,
            "cpu_kind": "reference arithmetic: numpy cost build + scipy.optimize.linear_sum_assignment, python loop over 256 images"}


def bind_near_gpu(local):
    """Pin this process to the CPUs NVML reports as local to its GPU (the driver's NUMA affinity mask) BEFORE any pinned host
    buffer is allocated: the end-to-end leg moves a 102 MB batch host->device every step and syncs on the loss, so a process
    that happens to run on the remote socket pays for it in `e2e` (the device-timed `value` does not depend on the host).
    Best effort: any failure leaves the affinity untouched.  Returns (original mask, bound mask or None)."""
    orig = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & set(orig)
        if len(cpus) >= 4 and cpus != set(orig):
            os.sched_setaffinity(0, cpus)
            return orig, sorted(cpus)
    except Exception:
        pass
    return orig, None


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_first = cpu_oracle_step_time(args.height, args.width, cores, steps=1, warmup=0)     # also the warm-up
    k = max(1, min(args.steps, int(150.0 / max(t_first, 1e-3))))
    t = cpu_oracle_step_time(args.height, args.width, cores, steps=k, warmup=0) if k > 1 else t_first
    v = 1.0 / t
    sample = f"{k} step(s) of 1 synthetic 800x1333 image each (full train step: fwd+matcher+set loss+bwd+Adam), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus, "steps": k,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DETR-R50 6enc/6dec, 100 queries, 800x1333 synthetic, full train step", "global_batch": 1,
                   "note": "TensorFlow is not installable offline: this is the PyTorch-CPU oracle restatement of the reference, not TF"},
        "cpu_baseline": {"value": v, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def dbg(msg):
    if os.environ.get("DETRB_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', 0)} t={time.time():.1f}] {msg}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU (BASELINE configs[1]: 8)")
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-matcher-bench", action="store_true")
    ap.add_argument("--backbone", default="resnet50", choices=["resnet50", "resnet101"],
                    help="resnet101 + --batch 4 = BASELINE configs[3] per GPU (DETR-R101, 32 images on 8 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import detr_tensorflow_b200 as D
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    orig_affinity, bound = bind_near_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, H, W, K, Wm = args.batch, args.height, args.width, args.steps, max(args.warmup, 3)

    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, B, None
    cfg.train_backbone, cfg.train_transformers = True, True
    model = D.get_detr_model(cfg, include_top=True, seed=0, backbone=args.backbone)    # identical replicas on every rank
    opt = D.setup_optimizers(model, cfg)
    eng = model.engine
    images, tb, tc = synthetic_batch(B, H, W, seed=rank)
    images_h, tb_h, tc_h = images.pin_memory(), tb.pin_memory(), tc.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dbg("model built")
    # ------------------------------------------------------------------ device-resident leg (`value`)
    eng.forward(images_h, training=True)                                # plans buffers, copies the batch into HBM
    eng.set_targets(tb_h, tc_h)
    eng.set_global_normalisers(tb_h)
    eng.set_lrs(cfg.backbone_lr, cfg.transformers_lr, cfg.nlayers_lr)
    eng.set_enabled(True, True)
    torch.cuda.synchronize()
    if args.no_graph:
        n0 = eng.launches
        eng.train_step(91, cfg.gradient_norm_clipping)
        eng.launches_per_step = eng.launches - n0
        step = lambda: eng.train_step(91, cfg.gradient_norm_clipping)
    else:
        step = eng.capture_train_step(91, cfg.gradient_norm_clipping)
    dbg("captured")
    for _ in range(Wm):
        step()
    barrier()
    dbg("warm")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step()
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    clocks = sampler.stop() if rank == 0 else None
    loss_after = float(eng.a["total"][0])
    value = world * B * K / (ms / 1e3)

    dbg(f"timed leg done {ms:.1f} ms")
    # ------------------------------------------------------------------ end-to-end leg through the public API
    # the call a user makes: training.fit over an iterable of HOST batches (pinned fp32 images + padded targets); every step
    # copies its batch host->device inside the timed region and reads the step's loss back (on_step hook -> float())
    Ke = K                                                              # same number of steps as the device-timed leg
    We = max(Wm, 3)                                                     # steps at the head of every fit() call that are not timed
    host_losses, stamps = [], []

    def on_step(step, total_loss, log):
        host_losses.append(float(total_loss))                           # device -> host read of the step's loss
        stamps.append(time.perf_counter())                              # = "step `step` is finished" (the read-back waited for it)

    def batches(n):
        for _ in range(n):
            yield images_h, tb_h, tc_h
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):                     # fit prints a progress line every 100 steps
        D.training.fit(model, batches(3), opt, cfg, 0, None, on_step=on_step)          # warm-up (captures the step graph)
        passes, calls = [], []
        # Three fit() calls of We + Ke + 1 steps.  fit calls the hook of step i after step i+1 has been enqueued, and the hook's
        # read-back (a synchronous copy on the compute stream) returns when everything enqueued so far has finished: hook We-1 returns
        # at the end of step We, the last hook (called after the loop) at the end of step We+Ke -> the region between them holds exactly
        # Ke full steps (copy in, compute, optimizer, loss out).  (Were the read-back to wait for its own step only, the same region
        # would hold Ke + 1 steps: the figure can only err on the low side.)  The first batch's un-overlapped copy and the step-0
        # progress print are warm-up; `whole_call_img_per_s` is the cross-check with everything included.
        for _ in range(3):
            barrier()
            del stamps[:]
            t_call = time.perf_counter()
            D.training.fit(model, batches(We + Ke + 1), opt, cfg, 0, None, on_step=on_step)
            t_pass = torch.tensor([stamps[We + Ke] - stamps[We - 1]], device="cuda")
            barrier()
            calls.append(world * B * (We + Ke + 1) / (time.perf_counter() - t_call))
            if world > 1:
                dist.all_reduce(t_pass, op=dist.ReduceOp.MAX)
            passes.append(float(t_pass))
    e2e_value = world * B * Ke / sorted(passes)[1]                      # the median pass
    h2d = images_h.numel() * 4 + tb_h.numel() * 4 + tc_h.numel() * 8
    dbg("e2e leg done")

    # ------------------------------------------------------------------ rooflines, timed live (CUDA events on the launching stream)
    # One object per kernel family of the step; `roofline` is the family with the largest share of the step BY TIME in the committed
    # ncu launch list (profiles/family_shares.json: the tcgen05 weight-gradient kernel, 22 %), the others follow in `rooflines`
    # (streaming GEMM 17 %, persistent GEMM / conv 16 %, attention 13 %, halo conv).  Every rank runs the probe steps (they contain
    # the gradient all-reduce), rank 0 reports.
    peak_sus, peak_burst, peak_hbm, how = measured_peaks()
    peak_tf, peak_tf_src = tensor_peak(clocks)
    if world > 1:                                                       # every rank needs the same probe set; clocks only exist on rank 0
        peak_tf, peak_tf_src = peak_sus, f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})"
    p_t, p_h = "backbone/layer3/1/conv2", "backbone/layer1/1/conv3"     # 3x3 256->256 @50x84 | 1x1 64->256 + residual + ReLU @200x334
    p_c = "backbone/layer1/1/conv2"                                     # 3x3 64->64 @200x334 (halo-reusing row kernel)
    eng.probe_names = (p_t, p_h, p_c, p_c + "#dgrad", p_t + "#wgrad", p_h + "#wgrad", p_h + "#dgrad", "e2_attn#fwd", "e2_attn#bwd")
    eng.probe_events = {}
    for _ in range(3):
        eng.train_step(91, cfg.gradient_norm_clipping)
    torch.cuda.synchronize()
    times = {k: sorted(a.elapsed_time(b) for a, b in v) for k, v in eng.probe_events.items()}
    # the weight gradients once more with the side stream off (same step, kernels serialised on the main stream): a kernel's own
    # duration, which is what a roofline fraction describes -- beside the data-gradient chain the same launch shares HBM / SMs with
    # another kernel and takes 1.5-1.7x as long (both figures are reported)
    eng.probe_names = (p_t + "#wgrad", p_h + "#wgrad", p_h + "#dgrad", p_c + "#dgrad")
    eng.probe_events = {}
    saved_overlap, eng.overlap_wgrad = eng.overlap_wgrad, False
    for _ in range(3):
        eng.train_step(91, cfg.gradient_norm_clipping)
    torch.cuda.synchronize()
    eng.overlap_wgrad = saved_overlap
    times_alone = {k: sorted(a.elapsed_time(b) for a, b in v) for k, v in eng.probe_events.items()}
    eng.probe_names = None
    roof, roofs = None, None
    if rank == 0:
        med = lambda k: times[k][len(times[k]) // 2] * 1e-3             # seconds
        med_alone = lambda k: times_alone[k][len(times_alone[k]) // 2] * 1e-3
        gf = TRAIN_GFLOP_PER_IMAGE[args.backbone]

        def traffic(fname):
            prof = os.path.join(ROOT, "profiles", fname)
            return json.load(open(prof)).get("dram_bytes_per_launch") if os.path.exists(prof) else None

        def tensor_obj(kernel, flops, t, tr=None):
            return {"bound": "tensor", "kernel": kernel, "achieved": flops / t / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": flops / t / 1e12 / peak_tf, "traffic": tr, "peak_source": peak_tf_src, "flops_per_launch": flops,
                    "us_per_launch": t * 1e6}

        def hbm_obj(kernel, byts, t, tr=None):
            return {"bound": "hbm", "kernel": kernel, "achieved": byts / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
                    "frac": byts / t / 1e9 / peak_hbm, "traffic": tr, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({how})",
                    "bytes_per_launch": byts, "us_per_launch": t * 1e6}
        st, sh = eng.slots[p_t], eng.slots[p_h]
        bt = [b for b in eng.blocks if b["c2"] is st][0]
        bh = [b for b in eng.blocks if b["c3"] is sh][0]
        Mt, Mh = B * bt["out_hw"][0] * bt["out_hw"][1], B * bh["out_hw"][0] * bh["out_hw"][1]
        S = eng.S
        # algorithmic bytes (bf16): forward = input + weights + residual + output; data gradient = dY + weights
        # + ReLU mask + output; weight gradient = input + dY (+ the fp32 gradient tile, negligible)
        # (+ M*N/8 bytes of 1-bit ReLU mask written by the forward, read by the data gradient instead of the bf16 activation)
        stream_obj = hbm_obj(f"gemm_stream_kernel<256,residual,bits-out> streaming tcgen05 GEMM (weights resident in smem, in-place chunk slots): "
                             f"1x1 conv 64->256 + residual + ReLU, layer1, M={Mh} N=256 K=64",
                             2.0 * (Mh * sh.K + sh.N * sh.K + 2 * Mh * sh.N) + Mh * sh.N / 8, med(p_h), traffic("hbm_kernel_traffic.json"))
        # `roofline` = the kernel with the largest share of the step BY TIME in the committed launch list (profiles/family_shares.json):
        # wgrad_tc_kernel, 22 % of the serialised step -- represented by its largest HBM-bound launch, timed where it runs: on the side
        # stream, sharing HBM with the data-gradient chain of the main stream (alone it runs at 0.85 of the copy bandwidth: `isolated`)
        roof = hbm_obj(f"wgrad_tc_kernel<plain> tcgen05 weight gradient (MN-major operands) of the layer1 1x1 conv 64->256, M={Mh} N=256 K=64 "
                       f"(timed inside a train step with the weight-gradient side stream off: the kernel alone on the GPU)",
                       2.0 * (Mh * sh.K + Mh * sh.N), med_alone(p_h + "#wgrad"), traffic("wgrad1x1_kernel_traffic.json"))
        roof["beside_data_gradient_chain"] = {"us_per_launch": med(p_h + "#wgrad") * 1e6,
                                              "note": "the same launch as it runs in the timed step: on the side stream, concurrently with the data-gradient "
                                                      "kernel of the main stream that reads the same dY (together they move ~690 MB in this time)"}
        roof["family"] = "wgrad_tc_kernel (tcgen05 weight gradients): the largest share of the serialised step by time, 22 % (profiles/family_shares.json); it runs on the side stream, overlapped with the data-gradient chain -- the largest family of the critical path is the streaming GEMM kernel (rooflines.hbm_conv1x1_layer1_stream)"
        roof["whole_step"] = {"achieved": gf * 1e9 * B * world * K / (ms / 1e3) / 1e12, "unit": "TFLOP/s (all GPUs)",
                              "frac_of_tensor_peak": gf * 1e9 * B * K / (ms / 1e3) / 1e12 / peak_tf, "peak": peak_tf, "peak_source": peak_tf_src}
        roofs = {
            "hbm_conv1x1_layer1_stream": stream_obj,
            "persistent_conv3x3": tensor_obj(f"gemm_tcp_kernel<256,4,im2col> persistent tcgen05 implicit-GEMM conv 3x3 256->256, layer3, M={Mt} N=256 K=2304",
                                             2.0 * Mt * st.N * st.K, med(p_t), traffic("top_kernel_traffic.json")),
            "wgrad_conv3x3": dict(tensor_obj(f"wgrad_tc_kernel<im2col> tcgen05 weight gradient of the same 3x3 conv, M={Mt} (side stream off: the kernel alone)",
                                             2.0 * Mt * st.N * st.K, med_alone(p_t + "#wgrad"), traffic("wgrad_kernel_traffic.json")),
                                  us_beside_data_gradient_chain=med(p_t + "#wgrad") * 1e6),
            "dgrad_conv1x1_layer1": hbm_obj(f"gemm_stream_kernel<64,bits-in>: data gradient of the layer1 1x1 conv 256->64 + 1-bit ReLU mask, M={Mh}",
                                            2.0 * (Mh * sh.N + sh.N * sh.K + Mh * sh.K) + Mh * sh.K / 8, med(p_h + "#dgrad"), traffic("stream_dgrad_kernel_traffic.json")),
            "conv3x3_layer1_halo": dict(tensor_obj(f"conv3x3_halo_kernel: 3x3 conv 64->64 + ReLU, layer1 (every input row staged once, nine taps = nine "
                                                   f"descriptor views), M={Mh} N=64 K=576", 2.0 * Mh * 64 * 576, med(p_c), traffic("halo_kernel_traffic.json")),
                                        hbm_gbs=(2.0 * 2 * Mh * 64 + Mh * 8) / med(p_c) / 1e9),
            "dgrad_conv3x3_layer1_halo": tensor_obj(f"conv3x3_halo_kernel: data gradient of the same conv (+ 1-bit ReLU mask), M={Mh}",
                                                    2.0 * Mh * 64 * 576, med(p_c + "#dgrad"), traffic("halo_dgrad_kernel_traffic.json")),
            "attention_fwd": tensor_obj(f"encoder self-attention forward, B={B} H=8 S={S} dh=32 (QK^T + PV flops; {B * 8 * S * S / 1e6:.1f} M exponentials)",
                                        4.0 * B * 8 * S * S * 32, med("e2_attn#fwd"), traffic("attn_fwd_kernel_traffic.json")),
            "attention_bwd": tensor_obj(f"encoder self-attention backward (delta + dK/dV + dQ kernels), B={B} H=8 S={S} dh=32",
                                        14.0 * B * 8 * S * S * 32, med("e2_attn#bwd"), traffic("attn_bwd_kernel_traffic.json")),
        }
        # the two data-gradient launches above are timed where they run: beside the weight-gradient kernels of the side stream, which
        # read the same dY (HBM and SMs are shared: 1.5-2x the kernel's own duration).  The kernel's own figure -- the same step with
        # the side stream off, what ncu sees for the isolated launch (profiles/r02_ncu_dgrad_kernels.txt) -- is reported next to it
        for key, probe, byts in (("dgrad_conv1x1_layer1", p_h + "#dgrad", 2.0 * (Mh * sh.N + sh.N * sh.K + Mh * sh.K) + Mh * sh.K / 8),
                                 ("dgrad_conv3x3_layer1_halo", p_c + "#dgrad", None)):
            if probe in times_alone and times_alone[probe]:
                ta = med_alone(probe)
                roofs[key]["alone"] = {"us_per_launch": ta * 1e6, "note": "side stream off: the kernel alone on the GPU"}
                if byts is not None:
                    roofs[key]["alone"].update(achieved=byts / ta / 1e9, frac=byts / ta / 1e9 / peak_hbm)
                else:
                    roofs[key]["alone"].update(achieved=roofs[key]["flops_per_launch"] / ta / 1e12, frac=roofs[key]["flops_per_launch"] / ta / 1e12 / peak_tf)
        shares = os.path.join(ROOT, "profiles", "family_shares.json")
        if os.path.exists(shares):
            roof["family_shares_of_step"] = json.load(open(shares))

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if bound and orig_affinity:
            os.sched_setaffinity(0, orig_affinity)                      # the CPU baseline may use every host core
        cores = os.cpu_count() or 1
        t_cpu = cpu_oracle_step_time(H, W, cores, steps=3, warmup=1)    # warm, like the reference arm (--impl reference)
        cpu = {"value": 1.0 / t_cpu, "unit": "images/sec", "cores": cores, "kind": "port",
               "sample": "3 train steps after 1 warm-up step (fwd+matcher+set loss+bwd+Adam) on 1 synthetic 800x1333 image, R50, "
                         "PyTorch-CPU oracle restatement of the reference (TensorFlow not installable offline)"}

    # ------------------------------------------------------------------ BASELINE configs[4]: matcher us/image (rank 0, N=1)
    matcher = None
    if rank == 0 and world == 1 and not args.no_matcher_bench:
        matcher = matcher_microbench(D)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC if args.backbone == "resnet50" else METRIC.replace("R50", "R101"),
            "value": value, "unit": "images/sec", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"DETR-{'R50' if args.backbone == 'resnet50' else 'R101'} 6enc/6dec, 100 queries, batch {B} per GPU, fixed {H}x{W} synthetic, full train step "
                                   "(fwd w/ dropout, on-device Hungarian + set loss, bwd, grad all-reduce, Adam+clipnorm)"
                                   + ("" if args.backbone == "resnet50" and B == 8 else " [BASELINE configs[3] per GPU]" if args.backbone == "resnet101" and B == 4 else ""),
                       "precision": "bf16 operands and activations, fp32 accumulation / master weights / optimizer / loss (the fp32-tolerance "
                                    "parity tests run the same kernels in precision='parity': bf16 pairs, three passes)",
                       "global_batch": world * B, "parallelism": f"dp{world}", "l2": "working set (>3 GB activations/step) exceeds the 126 MB L2",
                       "cuda_graph": not args.no_graph, "targets_per_image": 20},
            "clocks": clocks, "gpu_launches": int(eng.launches_per_step * K),
            "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "steps": Ke, "path": "training.fit(model, host_batches, optimizers, config, ...) with a per-step loss read-back (on_step hook -> float(); fit calls the hook of step i after step i+1 is enqueued); timed: `steps` full steps inside one fit() call of warm + steps + 1 steps (between two loss read-backs), median of three calls",
                    "passes_img_per_s": [world * B * Ke / t for t in passes], "whole_call_img_per_s": calls, "reported": "median of the three passes",
                    "cpu_affinity": ("bound to the GPU-local CPUs (NVML affinity): " + str(len(bound)) + " cpus") if bound else "unchanged"},
            "roofline": roof, "rooflines": roofs, "cpu_baseline": cpu, "matcher": matcher, "loss_after": loss_after,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
