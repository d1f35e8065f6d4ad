#!/usr/bin/env python
"""Benchmark of the volumetric-triangulation hot path (BASELINE.json metric: volumetric samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--mode tc|tc1|simt]

One "step" = one eval forward of a batch of synthetic multi-view samples (4 views of 3x384x384 -> 17 joints,
ResNet-152 backbone, 64^3 grid, softmax aggregation: BASELINE config #2, batch 8 per GPU) through the native
sm_100a path.  Prints ONE JSON line (rank 0).  See the module docstrings / DESIGN.md for the field definitions.

  value     device-timed throughput, inputs resident in HBM (CUDA events around K graph replays, max over ranks)
  e2e       same metric through the public nn.Module call with pinned-host images copied H2D and the keypoints
            read back D2H inside the timed region
  roofline  dominant kernel (tcgen05 conv): algorithmic conv FLOPs / summed per-launch CUDA-event time, vs the
            measured dense bf16 peak (MEASURED_PEAKS.json); unprojection / soft-argmax HBM rooflines alongside
  cpu_baseline  the CPU oracle port of the reference path timed on this box's host cores (rank 0, bounded sample)

--impl reference times the reference's CPU implementation of the path (the oracle port: the reference is pure
Python and is not present on the GPU box) with all host threads, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "volumetric samples/sec (4-view 384x384, 64^3 grid)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "torch_gpu"],
                    help="native: this repo's kernels; reference: the CPU oracle port of the reference path (contract arm); torch_gpu: the "
                         "same torch formulation as the reference on cuda:0 through ATen/cuDNN (secondary bar of SURVEY 8d, not a contract arm)")
    ap.add_argument("--tf32", action="store_true", help="torch_gpu only: allow TF32 in cuDNN / matmul (default: true fp32)")
    ap.add_argument("--mode", default=os.environ.get("LT_B200_CONV", "tc"), choices=["tc", "tc1", "simt"])
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per step")
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--image", type=int, default=384)
    ap.add_argument("--volume", type=int, default=64)
    ap.add_argument("--layers", type=int, default=152)
    ap.add_argument("--collective", default="features", choices=["all_reduce", "reduce_scatter", "p2p", "features"],
                    help="view-group exchange: one NCCL all-reduce (contract), reduce-scatter, or the fused P2P-store kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the ATen/cuDNN secondary bars (fp32 and TF32) of the native arm")
    ap.add_argument("--no-calibrate", action="store_true", help="keep PyTorch default init (faster start, degenerate signal)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_rate(args, steps, sd=None, images=None, batch=None):
    """samples/sec of the CPU oracle port of the reference path on this box's host cores.

    Bounded sample of the SAME workload: one untimed warm-up forward of one sample, then `steps` timed forwards of a
    whole batch (args.batch samples, ~20 s each for config #2 on 32 threads).  With `sd/images/batch` given (the
    native arm's own weights and inputs) the last forward's outputs are returned as well, for the parity object.
    """
    from oracle import parity
    import lt_b200
    from lt_b200 import testing
    # all host threads the reference can use productively: intra-op scaling of fp32 convs flattens (and on shared,
    # oversubscribed hosts reverses) beyond ~32 threads; LT_BENCH_CPU_THREADS overrides
    torch.set_num_threads(int(os.environ.get("LT_BENCH_CPU_THREADS", min(os.cpu_count(), 32))))
    if sd is None:
        cfg = testing.make_config(num_layers=args.layers, volume_size=args.volume)
        sd = lt_b200.VolumetricTriangulationNet(cfg, device="cpu", backend="torch").state_dict()   # parameter holder only
    if images is None:
        images, batch = testing.make_batch(args.batch, args.views, image_size=args.image, seed=0)
    B = images.shape[0]
    one = {k: ([c[:1] for c in v] if k == "cameras" else v[:1]) for k, v in batch.items()}
    parity.oracle_forward(sd, images[:1], one, args.volume)       # warm-up (thread pool, allocator)
    total, out = 0.0, None
    for _ in range(steps):
        out, secs = parity.oracle_forward(sd, images, batch, args.volume)
        total += secs
    dt = total / steps
    return B / dt, dt, torch.get_num_threads(), out


def main_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 2)), 1
    rate, dt, threads, _ = cpu_oracle_rate(args, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Volumetric(softmax) ResNet-%d, %d views %dx%d, %d^3 grid, batch %d per GPU"
                   % (args.layers, args.views, args.image, args.image, args.volume, args.batch),
                   "note": "CPU arm: %d timed forward(s) of one batch (bounded sample of the same workload), warm-up = one sample" % steps},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": "%d timed forward(s) of a %d-sample batch of %d-view inputs, CPU oracle port (torch fp32 + numpy), %d of %d host threads"
                                   % (steps, args.batch, args.views, threads, os.cpu_count())},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def torch_gpu_rate(args, dev, sd, images_dev, batch, tf32, steps=None):
    """The reference formulation (torch ops through ATen/cuDNN, eval, no autograd) on `dev`: samples/s, CUDA-event timed."""
    import lt_b200
    from lt_b200 import testing
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.benchmark = True
    try:
        cfg = testing.make_config(num_layers=args.layers, volume_size=args.volume)
        model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend="torch")
        model.load_state_dict(sd)
        model = model.to(dev).eval()
        steps = steps or max(2, min(args.steps, 5))
        with torch.no_grad():
            for _ in range(3):
                model(images_dev, None, batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(images_dev, None, batch)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del model
        torch.cuda.empty_cache()
        return {"value": images_dev.shape[0] / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "steps": steps,
                "dtype": "tf32" if tf32 else "f32", "what": "same module with backend='torch' (ATen/cuDNN, cudnn.benchmark) on the same GPU, weights and inputs"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


def main_torch_gpu(args, rank):
    """Secondary bar: the reference's own formulation -- torch ops through ATen/cuDNN -- on one B200, eval mode."""
    if rank != 0:
        return
    import lt_b200
    from lt_b200 import testing
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.benchmark = True
    B, V, S, n = args.batch, args.views, args.image, args.volume
    cfg = testing.make_config(num_layers=args.layers, volume_size=n)
    torch.manual_seed(0)
    model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend="torch")
    if not args.no_calibrate:
        testing.randomize_weights(model, seed=0, calib_size=S, calib_views=1)
    model = model.to(dev).eval()
    images, batch = testing.make_batch(B, V, image_size=S, seed=0)
    images_dev = images.to(dev)
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            model(images_dev, None, batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            model(images_dev, None, batch)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "torch_gpu", "metric": METRIC, "value": B * args.steps / (ms / 1e3), "unit": "samples/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "dtype": "tf32" if args.tf32 else "f32", "data": "synthetic",
                      "config": {"workload": "Volumetric(softmax) ResNet-%d, %d views %dx%d, %d^3 grid, batch %d, torch ops (ATen/cuDNN) on cuda:0"
                                 % (args.layers, V, S, S, n, B)}}), flush=True)


def main_native(args, rank, world, local_rank):
    import lt_b200
    from lt_b200 import testing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (the native path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, V, S, n = args.batch, args.views, args.image, args.volume
    cfg = testing.make_config(num_layers=args.layers, volume_size=n)
    torch.manual_seed(0)
    np.random.seed(0)
    sharded = world > 1
    model = lt_b200.VolumetricTriangulationNet(cfg, device=dev, backend="native", conv_mode=args.mode, use_cuda_graph=not sharded)
    if not args.no_calibrate:
        testing.randomize_weights(model, seed=0, calib_size=S, calib_views=1)
    model = model.to(dev).eval()
    eng = model.engine()
    model.clone_outputs = False
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    if not sharded:
        images, batch = testing.make_batch(B, V, image_size=S, seed=rank)
        pinned = images.pin_memory()
        images_dev = images.to(dev)
        parallelism = "dp1"

        def step(img):
            return model(img, None, batch)[0]
    else:
        # view-sharded: G ranks share a group batch of B*G samples and split its views; W/G groups are replicas
        from lt_b200 import dist as lt_dist
        plan = lt_dist.make_plan(world, rank, V)
        pg = lt_dist.new_view_groups(plan)
        Bg = B * plan.group_size
        images_g, batch = testing.make_batch(Bg, V, image_size=S, seed=100 + plan.group_index)
        views = plan.views
        images = images_g[:, views].contiguous()
        pinned = images.pin_memory()
        images_dev = images.to(dev)
        proj, base, position, stepv, rots, _ = model._host_geometry(batch, Bg, (S, S), (S // 4, S // 4))
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        geo = (up(proj[:, views]), up(position), up(base), up(stepv), up(rots.reshape(Bg, 9)), up(proj))
        parallelism = "view-sharded: %d group(s) x %d ranks, %d view(s)/rank, packed num/den %s, V2V batch-sharded" % (
            plan.n_groups, plan.group_size, len(views),
            "stored into the owner rank by the unprojection kernel over NVLink peer memory" if args.collective == "p2p"
            else ("replaced by a feature-map exchange over NVLink peer memory (unprojection on the owner)" if args.collective == "features"
                  else args.collective + " over NCCL"))

        def step(img):
            return eng.forward_view_sharded(img, geo[0], geo[1], geo[2], geo[3], geo[4], plan, pg, args.collective, proj_all=geo[5])[0]
    h2d = images.numel() * 4

    with torch.no_grad():
        if sharded and args.collective != "all_reduce":
            # the peer-memory exchanges need CUDA IPC / symmetric memory between all ranks of a view group; if any rank
            # cannot set them up, every rank falls back to the NCCL all-reduce contract path (reported in config)
            failed = 0
            try:
                step(images_dev)
                torch.cuda.synchronize()
            except Exception as exc:   # noqa: BLE001 - reported, then the contract path is used
                failed = 1
                print("rank %d: collective %r unavailable (%s: %s); using all_reduce" % (rank, args.collective, type(exc).__name__, exc),
                      file=sys.stderr, flush=True)
            flag = torch.tensor([failed], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()):
                parallelism += " [requested %s exchange unavailable on this box -> all_reduce over NCCL]" % args.collective
                args.collective = "all_reduce"
        # ---- device-resident arm: CUDA events, L2 flushed between iterations ----
        for _ in range(max(args.warmup, 3)):
            step(images_dev)
        barrier()
        launches = eng.launches
        sampler = ClockSampler(local_rank)
        sampler.start()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(images_dev)
            e1.record()
            evs.append((e0, e1))
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        clocks = sampler.summary()

        # ---- end-to-end arm: pinned host images -> device, forward, keypoints -> host ----
        for _ in range(2):
            kp = step(pinned.to(dev, non_blocking=True)).cpu()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            kp = step(pinned.to(dev, non_blocking=True)).cpu()
        barrier()
        e2e_s = time.perf_counter() - t0
        d2h = kp.numel() * 4
        e2e_sync_s = e2e_s

        # ---- end-to-end, pipelined (single-GPU path): the collated HWC batch in pinned memory goes through
        # lt_b200.pipeline.InferenceStream -- every step still uploads its own images and reads its own keypoints back,
        # but the upload of batch i+1 overlaps the forward of batch i ----
        if not sharded:
            from lt_b200 import pipeline
            hwc = pipeline.pinned_empty((B, V, S, S, 3), np.float32)
            hwc[...] = images.permute(0, 1, 3, 4, 2).numpy()
            stream = pipeline.InferenceStream(model)

            def feed(k):
                for _ in range(k):
                    b = dict(batch)
                    b["images"] = hwc
                    yield b
            for _ in stream.run(feed(3)):
                pass
            barrier()
            stream.h2d_bytes = stream.d2h_bytes = 0
            t0 = time.perf_counter()
            for kp_host in stream.run(feed(args.steps)):
                pass
            barrier()
            e2e_s = time.perf_counter() - t0
            assert stream.h2d_bytes == h2d * args.steps and stream.d2h_bytes == d2h * args.steps

        # ---- per-kernel timing for the roofline: one eager (non-graph) forward with event pairs per launch ----
        eng.use_graph = False
        model.use_cuda_graph = False
        step(images_dev)
        eng.timeline = []
        step(images_dev)
        torch.cuda.synchronize()
        agg = {}
        per_launch = []
        for label, flops, nbytes, a, b, desc in eng.timeline:
            r = agg.setdefault(label, [0.0, 0.0, 0.0, 0])
            ms = a.elapsed_time(b)
            r[0] += ms; r[1] += flops; r[2] += nbytes; r[3] += 1
            per_launch.append({"kernel": label, "desc": desc, "ms": round(ms, 4), "gflop": round(flops / 1e9, 3), "mb": round(nbytes / 1e6, 3)})
        if os.environ.get("LT_BENCH_TIMELINE") and rank == 0:
            with open(os.environ["LT_BENCH_TIMELINE"], "w") as f:
                json.dump(per_launch, f)
        eng.timeline = None

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])
    total_samples = B * world * args.steps
    value = total_samples / (dev_ms / 1e3)
    e2e = total_samples / e2e_s

    pk = peaks()
    roof = None
    extra = {}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # DRAM bytes per launch from the committed ncu launch list
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    tensor_kernels = [k for k in ("conv_tc", "conv_fold", "conv_ffma") if k in agg]
    conv_key = max(tensor_kernels, key=lambda k: agg[k][0]) if tensor_kernels else None

    def tensor_roof(key):
        ms, fl, _, cnt = agg[key]
        ach = fl / (ms / 1e3) / 1e12
        tc = key != "conv_ffma"
        peak = pk["bf16_tflops"] if tc else 75.0
        r = {"kernel": {"conv_tc": "conv_tc_kernel", "conv_fold": "conv_fold_kernel", "conv_ffma": "conv_simt_kernel"}[key],
             "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
             "traffic": traffic.get({"conv_tc": "conv_tc_kernel", "conv_fold": "conv_fold_kernel"}.get(key)),
             "launches": cnt, "ms_per_step": ms,
             "peak_source": pk["src"] + (" sustained dense bf16/fp16 (cuBLAS)" if tc else " nominal fp32 FFMA")}
        if tc and args.mode == "tc":
            # fp32-grade results cost three fp16 products per term (hi*hi, hi*lo, lo*hi): tensor-pipe work actually issued
            r["issued_mma_tflops"] = 3.0 * ach
            r["issued_frac"] = 3.0 * ach / peak
        return r

    if conv_key:
        roof = tensor_roof(conv_key)
    for key in tensor_kernels:
        if key != conv_key:
            extra["roofline_" + key] = tensor_roof(key)
    for key in ("unproject", "softargmax"):
        if key in agg:
            ms, fl, nb, cnt = agg[key]
            ach = nb / (ms / 1e3) / 1e9
            extra["roofline_" + key] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                        "frac": ach / pk["hbm_gbs"], "ms_per_step": ms, "launches": cnt, "traffic": traffic.get(key),
                                        "peak_source": pk["src"] + " copy bandwidth"}
    extra["step_breakdown_ms"] = {k: round(v[0], 3) for k, v in agg.items()}

    cpu = None
    parity_obj = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU arm runs the bench's own weights and inputs, so its outputs double as the parity check of this very run:
        # full-size config #2 batch through the native path (graph-free replay of the timed step) vs the oracle
        from oracle import parity
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            native_out = model(images_dev, None, batch)
        torch.cuda.synchronize()
        rate, dt, threads, oracle_out = cpu_oracle_rate(args, 1, sd_cpu, images, batch)
        cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "1 warm-up forward of one sample + 1 timed forward of the bench's %d-sample batch (same config, same weights and "
                         "inputs as the GPU arm), CPU oracle port, %d of %d host threads" % (B, threads, os.cpu_count())}
        parity_obj = parity.compare_outputs(native_out, oracle_out)
        parity_obj["against"] = "oracle/vol_oracle.volumetric_forward (CPU restatement of reference triangulation.py:245-355) on the bench batch"
        del native_out, oracle_out

    torch_bars = {}
    if rank == 0 and world == 1 and not args.no_torch_gpu:
        # secondary bar: the reference's own torch formulation through ATen/cuDNN on the same B200, same weights/inputs
        sd_dev = model.state_dict()
        for key, tf32 in (("torch_gpu", False), ("torch_gpu_tf32", True)):
            try:
                torch_bars[key] = torch_gpu_rate(args, dev, sd_dev, images_dev, batch, tf32)
                torch_bars[key]["native_over_this"] = value / torch_bars[key]["value"]
            except Exception as exc:   # noqa: BLE001 - reported in the line
                torch_bars[key] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tc": "fp16x3 (split-fp16 operands, 3 tcgen05 products per term, fp32 accumulate)", "tc1": "fp16", "simt": "f32"}[args.mode],
            "data": "synthetic",
            "config": {"workload": "Volumetric(softmax) ResNet-%d, %d views %dx%d, %d^3 grid, batch %d per GPU"
                                   % (args.layers, V, S, S, n, B),
                       "global_batch": B * world, "parallelism": parallelism,
                       "conv_mode": args.mode, "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "weights": "random (seeded recipe, BN calibrated)" if not args.no_calibrate else "random (default init)"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": ("lt_b200.pipeline.InferenceStream(model).run(batches): pinned HWC batch -> H2D on a copy stream -> layout kernel "
                            "-> forward -> keypoints D2H, upload of batch i+1 overlapping forward i") if not sharded
                           else "engine.forward_view_sharded per step, synchronous",
                    "sync_value": total_samples / e2e_sync_s,
                    "sync_api": "model(images_pinned.to(device, non_blocking=True), None, batch)[0].cpu() per step, no overlap"},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity_obj,
        }
        line.update(torch_bars)
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank)
    elif args.impl == "torch_gpu":
        main_torch_gpu(args, rank)
    else:
        main_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
