#!/usr/bin/env python
"""Benchmark of the ERFNet hot path on B200 (BASELINE.json metric: ERFNet 512x1024 fwd & train images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  = training images/s (config[1]: ERFNet(20) encoder+decoder, batch 6/GPU at 512x1024, weighted
           CrossEntropy2d, Adam; data-parallel, weak scaling), inputs resident in HBM, CUDA-event timed,
           max over ranks.  `fwd` carries the inference (eval-mode, BN folded) images/s of the same net.
`e2e`    = the same training step through the public API with HOST buffers (FlatAdamTrainer.train_stream(..., host_losses=True)): pinned H2D
           copy of every step's images and labels and a D2H read of every step's loss inside the timed region (each loss is handed to the
           host loop one step behind the launches, so reading it does not stall the next launch).
`fwd_b1` / `fwd_b32` / `eval_fullres` (ERF_BENCH_EXTRA=0 skips them) = inference at batch 1 (BASELINE configs[0], the eval_forwardTime.py
           shape: latency), at batch 32 (enough work to fill the 148 SMs) and configs[4] (batch 8 x 1024x2048 forward + argmax + iouEval.addBatch).
`--impl reference` times the reference's CPU path (the oracle port of its torch arithmetic; the reference is pure
Python and /root/reference does not exist on the GPU box) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, BATCH, NC = 512, 1024, 6, 20
GFLOP_FWD_PER_IMG = 53.225
GFLOP_TRAIN_PER_IMG = 159.6
NCU_TRAFFIC_BYTES = 19.3e6   # DRAM read+write per conv1d_tc launch, launch-weighted over a step's mix (56 x 25.2 MB at C=64, 64 x 12.75 MB at C=128,
                             # 16 x ~25 MB at C=16): ncu --set full, profiles/r02_ncu_full_tables.md


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=p["hbm_gbs"], tc=p["bf16_tflops"], tc_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tc=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (20 ms period, no start-up delay), falling
    back to an `nvidia-smi -lms` child process when NVML is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.samples, self.proc, self.stop, self.t = index, [], None, False, None
        self.max_mhz, self.reasons = 0, set()

    def _nvml_loop(self, nv, h):
        while not self.stop:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            self.t = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            s = [v.strip() for v in line.split(",")]
            if s and s[0].isdigit():
                self.samples.append(int(s[0]))
                if len(s) > 1 and s[1].isdigit():
                    self.max_mhz = max(self.max_mhz, int(s[1]))
                for n, v in zip(self.NAMES, s[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        elif self.t is not None:
            self.t.join(timeout=1)

    def summary(self):
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz or None, "reasons": sorted(self.reasons), "samples": len(sm)}


def synth_batch(torch, seed, batch=BATCH, h=H, w=W):
    """Synthetic Cityscapes-shaped data: piece-wise constant 32x32 label blocks, ~5 % ignored,
    image = per-class colour + noise."""
    g = torch.Generator().manual_seed(seed)
    blocks = torch.randint(0, NC - 1, (batch, h // 32, w // 32), generator=g)
    ign = torch.rand(batch, h // 32, w // 32, generator=g) < 0.05
    blocks = torch.where(ign, torch.full_like(blocks, -1), blocks)
    labels = blocks.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    palette = torch.rand(NC + 1, 3, generator=g)
    img = palette[labels.clamp_min(0)].permute(0, 3, 1, 2) * 0.8 + 0.1 * torch.rand(batch, 3, h, w, generator=g)
    return img.contiguous(), labels


def cpu_reference_run(steps, warmup, batch=1, fwd_only=False):
    """The reference's CPU path (oracle port): train steps (fwd + CE + bwd + Adam) at `batch` x 512x1024."""
    import torch
    from oracle import erfnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.make_state_dict(O.erfnet_param_shapes(NC), 0)
    x, y = synth_batch(torch, 0, batch)
    wv = torch.linspace(0.5, 1.5, NC)
    times = []
    if fwd_only:
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.erfnet_forward(x, sd)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return batch * len(times) / sum(times)
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v2 = {k: torch.zeros_like(v) for k, v in params.items()}
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params.values():
            p.grad = None
        loss = O.cross_entropy_2d(O.erfnet_forward(x, sd, training=True), y, weight=wv)
        loss.backward()
        with torch.no_grad():
            for k, p in params.items():
                if p.grad is not None:
                    O.adam_step(p, p.grad, m[k], v2[k], i + 1)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times)


def bench_config(batch, world, use_graph=True):
    """The `config` object of both arms (the reference arm measures the SAME workload; its bounded per-step sample is described in its
    cpu_baseline.sample, not here)."""
    return {"workload": f"ERFNet({NC}) encoder+decoder training, batch {batch}/GPU at {H}x{W}, weighted CrossEntropy2d, Adam (BASELINE configs[1])",
            "global_batch": batch * world, "parallelism": f"dp{world}", "l2": "inputs rotate over a 4-batch pool (252 MB) > L2; per-step working set >> L2",
            "random_init": "deterministic synthetic weights (pretrained blob missing)", "cuda_graph": bool(use_graph)}


def run_reference(args):
    """The reference's CPU path on the host cores: exactly --warmup untimed and --steps timed steps, each step a BOUNDED SAMPLE of the
    workload (one 512x1024 image instead of the batch of 6: images/s is per image, so the metric is the same); ~0.25 s per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    ips = cpu_reference_run(steps, warmup, batch=1)
    cores = os.cpu_count()
    sample = (f"{steps} timed train steps (fwd + weighted CE + bwd + Adam) after {warmup} warm-up, each a bounded sample of the workload: batch 1 x 3x{H}x{W} fp32 "
              f"(oracle port of the reference's torch ops, {cores} threads)")
    line = {"impl": "reference", "metric": "train_images_per_s", "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1000.0 / ips, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.batch, args.gpus, not args.no_graph),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from erfnet_pytorch_b200 import erfnet as E
    from erfnet_pytorch_b200 import ops
    from erfnet_pytorch_b200.parallel import FlatAdamTrainer
    from erfnet_pytorch_b200.synthetic import synthetic_state_dict   # deterministic "trained-like" init (the GPU arm never touches oracle/)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, max(3, args.warmup)
    pk0 = peaks()
    B = args.batch

    model = E.ERFNet(NC)
    model.load_state_dict(synthetic_state_dict(model, 0))
    model.precision = args.precision
    model = model.to(dev)
    trainer = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))
    use_graph = not args.no_graph and (world == 1 or os.environ.get("ERF_GRAPH_DDP", "1") == "1")

    # a pool of distinct batches larger than L2 (126 MB) so inputs are never L2-resident between steps
    pool = [synth_batch(torch, 100 * rank + i, B) for i in range(4)]
    pool_dev = [(x.to(dev), y.to(dev)) for x, y in pool]
    pool_pin = [(x.pin_memory(), y.pin_memory()) for x, y in pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as cs:
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record()
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, cs.summary()

    # ---- training, inputs resident in HBM
    def train_resident(i):
        x, y = pool_dev[i % len(pool_dev)]
        trainer.train_step(x, y)

    l0 = ops.LAUNCHES
    train_resident(0)  # one eager step: counts this library's kernel launches per step (and warms every code path)
    launches_per_step = ops.LAUNCHES - l0
    if use_graph:
        trainer.capture_cuda_graph((B, 3, H, W), (B, H, W))
    train_ms, clocks = timed(train_resident, K, Wm)
    launches = launches_per_step * K

    # ---- training end to end: pinned host buffers -> device, loss back to the host, every step
    last = {}

    # The public call is FlatAdamTrainer.train_stream(host batches): every step copies ITS batch from pinned host memory (on a side
    # stream, overlapping the previous step's compute) and the loop reads every step's loss back to the host.
    def e2e_run(steps):
        for loss in trainer.train_stream((pool_pin[i % len(pool_pin)] for i in range(steps)), host_losses=True):
            last["loss"] = loss       # a Python float: every step's loss is read back on the host (one step behind the launches)

    e2e_run(2)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    e2e_run(K)
    ev1.record()
    barrier()
    e2e_t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t) / K

    # ---- inference forward (eval mode, BN folded), resident inputs, NCHW fp32 logits out (reference API)
    model.eval()
    if use_graph:
        from erfnet_pytorch_b200.graphs import GraphedInference
        gi = GraphedInference(model, pool_dev[0][0])

    def fwd(i):
        if use_graph:
            gi(pool_dev[i % len(pool_dev)][0])
        else:
            with torch.no_grad():
                model(pool_dev[i % len(pool_dev)][0])

    fwd_ms, _ = timed(fwd, K, Wm)

    # ---- inference at a batch that fills the 148 SMs (the 128-channel layers of a batch of 6 are only 2.6 waves of 128-pixel tiles), and
    # BASELINE configs[4]: batch 8 at 1024 x 2048 -> fused argmax -> iouEval confusion histogram (what eval_iou.py does per batch)
    extra = {}
    if os.environ.get("ERF_BENCH_EXTRA", "1") == "1":
        from erfnet_pytorch_b200.iouEval import iouEval
        gi = None
        torch.cuda.empty_cache()
        BB = 32
        xb = torch.cat([pool_dev[i % len(pool_dev)][0] for i in range((BB + B - 1) // B)])[:BB].contiguous()
        xs = [xb, xb.flip(0).contiguous()]          # 2 x 403 MB of input: never L2-resident
        gb = GraphedInference(model, xs[0]) if use_graph else None

        def fwd_big(i):
            if gb is not None:
                gb(xs[i % 2])
            else:
                with torch.no_grad():
                    model(xs[i % 2])

        big_ms, _ = timed(fwd_big, max(3, K // 2), 3)
        extra["fwd_b32"] = {"images_per_s": BB * world / (big_ms * 1e-3), "ms_per_step": big_ms, "batch": BB,
                            "tc_frac": BB / (big_ms * 1e-3) * GFLOP_FWD_PER_IMG / 1e3 / pk0["tc_sustained"]}
        del gb, xs, xb
        torch.cuda.empty_cache()
        # BASELINE configs[0] / eval/eval_forwardTime.py:31-55: batch 1 x 3 x 512 x 1024 forward latency (graph replay; inputs alternate)
        x1 = [pool_dev[i][0][:1].contiguous() for i in range(2)]
        g1 = GraphedInference(model, x1[0]) if use_graph else None

        def fwd_one(i):
            if g1 is not None:
                g1(x1[i % 2])
            else:
                with torch.no_grad():
                    model(x1[i % 2])

        one_ms, _ = timed(fwd_one, max(10, K), 3)
        extra["fwd_b1"] = {"workload": "configs[0] shape: 1 x 3 x 512 x 1024 (eval_forwardTime.py), bf16", "ms_per_image": one_ms, "images_per_s": world / (one_ms * 1e-3)}
        del g1, x1
        torch.cuda.empty_cache()
        HF, WF, BF = 1024, 2048, 8
        gen = torch.Generator(device=dev).manual_seed(7 + rank)
        xf = [torch.rand(BF, 3, HF, WF, device=dev, generator=gen) for _ in range(2)]
        lab = torch.randint(0, NC, (BF, HF, WF), device=dev, generator=gen, dtype=torch.int64).to(torch.uint8)   # 19 = ignore (eval_iou.py:37)
        ev = iouEval(NC)

        def eval_full(i):
            with torch.no_grad():
                logits = model.forward_nhwc(xf[i % 2])
                pred = ops.argmax(logits, NC, False, torch.uint8)
                ev.addBatch(pred.unsqueeze(1), lab.unsqueeze(1))

        full_ms, _ = timed(eval_full, max(3, K // 2), 3)
        miou = float(ev.getIoU()[0])
        extra["eval_fullres"] = {"workload": f"configs[4]: batch {BF} x 3x{HF}x{WF} forward + argmax + iouEval.addBatch, eager launches",
                                 "frames_per_s": BF * world / (full_ms * 1e-3), "ms_per_batch": full_ms, "mean_iou_random_labels": miou,
                                 "tc_frac": BF / (full_ms * 1e-3) * 4 * GFLOP_FWD_PER_IMG / 1e3 / pk0["tc_sustained"]}
        del xf, lab
        torch.cuda.empty_cache()

    # ---- dominant kernel: per-launch CUDA events on one extra (untimed, eager) training step.  The GPU is first parked on
    # a 40 ms spin so that the whole step is already queued when it starts: event intervals then measure kernels, not the
    # CPU launch path.
    model.train()
    saved_graph, trainer._graph = trainer._graph, None
    x0, y0 = pool_dev[0]
    trainer.train_step(x0, y0)  # untimed eager pass: lets the caching allocator grow outside the measured one
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.04 * 1.9e9))
    ops.start_kernel_timing()
    ops.CONV1D_CALL_LOG = []
    trainer.train_step(x0, y0)
    torch.cuda.synchronize()
    ktimes = ops.stop_kernel_timing()
    calls, ops.CONV1D_CALL_LOG = ops.CONV1D_CALL_LOG, None
    trainer._graph = saved_graph
    # The event pairs of the pass above break the programmatic-dependent-launch chain (every interval also holds ~2-3 us of launch latency that
    # the replayed step hides).  The dominant kernel's launches of that very step -- same tensors, shapes, epilogue options, in order -- are
    # therefore re-issued back to back under ONE event pair: their average duration as they run inside the step.
    b2b_ms, conv_bytes = None, 0.0
    if calls:
        # algorithmic bytes of a launch: input + output (+ residual / mask operand), bf16
        conv_bytes = float(sum(c[0].numel() * 2 * (2 + int(c[6] is not None or c[7] is not None)) for c in calls)) / len(calls)
        reps = 3
        scratch = {}
        def reissue():
            for (cx, cw, cb, ch, cd, cr, cres, cmask, cstats, ccol, cst) in calls:
                if cstats is not None:
                    cstats = scratch.setdefault(("s", cstats.numel()), torch.zeros_like(cstats))
                if ccol is not None:
                    ccol = scratch.setdefault(("c", ccol.numel()), torch.zeros_like(ccol))
                ops.conv1d_tc(cx, cw, cb, ch, cd, relu=cr, res=cres, mask=cmask, stats=cstats, colsum=ccol, static_w=cst)
        reissue()
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.04 * 1.9e9))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            reissue()
        e1.record()
        torch.cuda.synchronize()
        b2b_ms = e0.elapsed_time(e1) / (reps * len(calls))
    del calls
    pk = peaks()
    roof = None
    if ktimes:
        tot = sum(r["ms"] for r in ktimes.values())
        name = "conv1d_tc_kernel" if "conv1d_tc_kernel" in ktimes else max(ktimes.items(), key=lambda kv: kv[1]["ms"])[0]
        rec = ktimes[name]
        avg_ms = b2b_ms if (b2b_ms and name == "conv1d_tc_kernel") else rec["ms"] / rec["n"]
        ach = rec["flops"] / rec["n"] / (avg_ms * 1e-3) / 1e12 if rec["flops"] else None
        # Which roofline bounds this kernel: an un-fused 3-tap conv moves 4 bytes per 6*C flops (96 / 192 flop/B at C = 64 / 128, 24 at C = 16),
        # below the ridge of this GPU (1427 TFLOP/s / 6.4 TB/s = 221 flop/B): as designed it is HBM-bound, so `frac` is the HBM fraction; the
        # tensor-pipe view of the same launches is kept under "tensor".
        gbs = conv_bytes / (avg_ms * 1e-3) / 1e9 if conv_bytes else None
        tensor_view = {"achieved": ach, "peak": pk["tc_sustained"], "unit": "TFLOP/s", "frac": (ach / pk["tc_sustained"]) if ach else None,
                       "peak_source": pk["src"] + " (sustained bf16 dense)"}
        roof = {"kernel": name, "bound": "hbm" if gbs else "tensor", "achieved": gbs if gbs else ach, "peak": pk["hbm"] if gbs else pk["tc_sustained"],
                "unit": "GB/s" if gbs else "TFLOP/s", "frac": (gbs / pk["hbm"]) if gbs else ((ach / pk["tc_sustained"]) if ach else None),
                "algorithmic_bytes_per_launch": conv_bytes or None, "tensor": tensor_view,
                "traffic": NCU_TRAFFIC_BYTES, "traffic_note": "dram read+write per launch, weighted over the launch mix of a step, ncu --set full (profiles/r02_ncu_full_tables.md): the input is read once from DRAM, taps and the freshly written output are L2 traffic (126 MB L2), so DRAM traffic stays BELOW the algorithmic bytes", "peak_source": pk["src"] + (" (copy bandwidth)" if gbs else " (sustained bf16 dense)"),
                "launches": rec["n"], "share_of_step": rec["ms"] / tot if tot else None, "avg_launch_ms": avg_ms,
                "avg_launch_ms_how": "the step's launches of this kernel re-issued back to back (PDL chain intact) under one CUDA-event pair" if b2b_ms else "per-launch event pairs",
                "avg_launch_ms_event_pairs": rec["ms"] / rec["n"],
                "flops_per_launch": rec["flops"] / rec["n"] if rec["flops"] else None,
                "step_kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1]["ms"])[:8]}}

    if rank == 0:
        ips = B * world / (train_ms * 1e-3)
        line = {"metric": "train_images_per_s", "value": ips, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": train_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": bench_config(B, world, use_graph),
                "clocks": clocks,
                "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": world * B * (3 * H * W * 4 + H * W * 8), "d2h_bytes_per_step": 4 * world,
                        "ms_per_step": e2e_ms, "loss": last.get("loss")},
                "gpu_launches": launches,
                "fwd": {"images_per_s": B * world / (fwd_ms * 1e-3), "ms_per_step": fwd_ms, "batch": B,
                        "tc_frac": B * world / (fwd_ms * 1e-3) * GFLOP_FWD_PER_IMG / 1e3 / (pk["tc_sustained"] * world)},
                "train_tc_frac": ips * GFLOP_TRAIN_PER_IMG / 1e3 / (pk["tc_sustained"] * world),
                "roofline": roof}
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:
            t0 = time.time()
            cpu_ips = cpu_reference_run(40, 1, batch=1)                    # ~8-10 s on 16 cores
            cpu_fwd = cpu_reference_run(60, 1, batch=1, fwd_only=True)     # ~3-4 s
            line["cpu_baseline"] = {"value": cpu_ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"40 train steps + 60 forwards of batch 1 x 3x{H}x{W} fp32 (oracle port of the reference's torch ops), {time.time() - t0:.0f} s of CPU work",
                                    "fwd_images_per_s": cpu_fwd}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
