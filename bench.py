#!/usr/bin/env python
"""Headline benchmark: gelan-c 640x640 detection inference, images/sec INCLUDING NMS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

One "step" = one pass of the hot path (forward + DFL decode + batched NMS, conf 0.25 / iou 0.45) over a
batch of 64 synthetic 640x640 images per GPU (BASELINE.json configs[1]; at N=8 the global batch is the
512 of configs[2]).  Images shard across ranks with no collective on the data path;
NCCL only gathers the timing.  Prints ONE JSON line (rank 0).

  value      images/s, device-timed (CUDA events), inputs already resident in HBM
  e2e        same metric through the public API from pinned HOST memory: H2D of the batch and D2H of
             the detections inside the timed region, every step
  roofline   the conv kernels (dominant, tensor-bound): algorithmic FLOPs of the folded graph /
             their summed device time, measured per launch with CUDA events
  cpu_baseline  the oracle port of the reference forward+NMS on the box's host cores (bounded sample)

`--impl reference` times the reference algorithm's CPU port (oracle/, the reference itself is pure
Python/PyTorch and is not present on the GPU box) on the same config/metric.

Use of oracle/ here: (1) the cpu_baseline leg and the reference arm execute it (that is what they measure);
(2) both arms take their SYNTHETIC DATA from it -- `calibrated_state_dict` (the calibrated random-init weight recipe of
SURVEY.md 8d, which needs one CPU train-mode forward to set the BatchNorm statistics) and the graph description it is
built from -- so that the two arms run the same weights.  No number of the timed GPU path is computed by oracle/: the
product (yolo_b200 + libyre.so) never imports it and fails loudly without the CUDA library.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "yolo-re_b200"))

import torch  # noqa: E402

CFG = ROOT / "configs" / "models" / "gelan-c.yaml"
IMG = 640
PER_GPU_BATCH = 64
CONF, IOU, MAX_DET = 0.25, 0.45, 300
GF_PER_IMAGE_FOLDED = 102.136e9      # SURVEY.md section 8d, gelan-c @640 folded graph


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1374.1), d.get("hbm_gbs", 6549.8), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(batch: int, seed: int = 7) -> torch.Tensor:
    """Seeded multi-octave images in [0,1] (oracle.gelan_ref.fractal; matches /255 preprocessing).  With the
    calibrated weights ~1-5 % of the anchors pass conf 0.25, so NMS does real work."""
    from oracle import gelan_ref as G
    g = torch.Generator().manual_seed(seed)
    base = G.fractal(min(batch, 8), IMG, g)
    reps = -(-batch // base.shape[0])
    x = base.repeat(reps, 1, 1, 1)[:batch].clone()
    # de-duplicate the repeats with a per-image brightness ramp (keeps values in [0,1])
    x *= torch.linspace(0.85, 1.0, batch).view(-1, 1, 1, 1)
    return x


def cpu_reference_step(nodes, nc, sd, x):
    from oracle import gelan_ref as G
    from oracle import nms_ref as N
    y, _ = G.forward(nodes, nc, sd, x)
    return N.non_max_suppression(y.permute(0, 2, 1).contiguous(), CONF, IOU, MAX_DET)


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import gelan_ref as G
    torch.set_num_threads(os.cpu_count() or 1)
    nodes, nc = G.load_graph(CFG)
    sd = G.calibrated_state_dict(nodes, nc)
    sample = 4
    x = make_inputs(sample)
    for _ in range(args.warmup):
        cpu_reference_step(nodes, nc, sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(nodes, nc, sd, x)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    print(json.dumps({
        "impl": "reference", "metric": "gelan-c 640x640 images/sec incl. NMS", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "gelan-c inference 640x640 + DFL decode + NMS(conf=0.25, iou=0.45), calibrated random-init weights",
                   "per_gpu_batch": PER_GPU_BATCH, "sample_batch": sample},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} images/step x {args.steps} steps, oracle port of the reference forward + NMS (CPU fp32)"},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (default 64 = the metric's config)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-op", default="", help="write a per-launch CSV (kernel, shape, ms, TFLOP/s) to this path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus, all_cpus = None, os.sched_getaffinity(0)
    if os.environ.get("YRE_BENCH_NUMA", "1") != "0":
        from yolo_b200.shard import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)       # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from oracle import gelan_ref as G
    from yolo_b200 import YOLO, nms_raw, non_max_suppression

    torch.set_num_threads(max(1, (os.cpu_count() or 8) // max(world, 1)))
    nodes, nc = G.load_graph(CFG)
    sd = G.calibrated_state_dict(nodes, nc)          # deterministic: every rank builds identical weights
    model = YOLO.from_yaml(CFG)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval().set_precision(args.precision)
    model.check_weights = False
    model.fresh_outputs = False
    model.use_cuda_graph = os.environ.get("YRE_BENCH_GRAPH", "1") != "0"     # static buffers -> the forward replays as one CUDA graph

    Bn = args.batch
    x_host = make_inputs(Bn, seed=7 + rank).pin_memory()
    x_dev = x_host.to(dev)

    def step_resident():
        y, _ = model(x_dev)
        return nms_raw(y.permute(0, 2, 1), CONF, IOU, MAX_DET)

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up (also compiles the plan) ----
    for _ in range(args.warmup):
        out, counts, keep = step_resident()
    torch.cuda.synchronize(dev)
    n_cand_frac = None
    plan = next(iter(model._plans.values()))
    launches_per_step = plan.num_launches + 3

    # ---- device-timed throughput, inputs resident ----
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, counts, keep = step_resident()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * Bn * args.steps / (ms_max / 1e3)
    dets_per_img = float(counts.float().mean().item())

    # ---- end-to-end through the public API from pinned host memory ----
    copy_stream = torch.cuda.Stream(dev)
    bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    out_host = torch.empty((Bn, MAX_DET, 6), dtype=torch.float32).pin_memory()
    cnt_host = torch.empty((Bn,), dtype=torch.int32).pin_memory()

    def e2e_run(steps):
        main_s = torch.cuda.current_stream(dev)
        for b in range(2):
            done[b].record(main_s)
        with torch.cuda.stream(copy_stream):                 # prefetch step 0
            copy_stream.wait_event(done[0])
            bufs[0].copy_(x_host, non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(steps):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < steps:
                with torch.cuda.stream(copy_stream):         # H2D of step i+1 overlaps the compute of step i
                    copy_stream.wait_event(done[nxt])
                    bufs[nxt].copy_(x_host, non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_s.wait_event(ready[cur])
            y, _ = model(bufs[cur])
            o, c, _k = nms_raw(y.permute(0, 2, 1), CONF, IOU, MAX_DET)
            done[cur].record(main_s)
            out_host.copy_(o, non_blocking=True)             # D2H of the step's result
            cnt_host.copy_(c, non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_run(2)
    sync_all()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * Bn * args.steps / float(t.item())
    h2d = x_host.numel() * 4
    d2h = out_host.numel() * 4 + cnt_host.numel() * 4

    # ---- extra (informational): the same step fed from uint8 HWC camera frames through K8 (yolo_b200.preprocess) ----
    # What scripts/detect.py:223-227 does on the host -- BGR->RGB, HWC->CHW, /255 -- happens on the device, so the H2D
    # is 4x smaller.  Not the headline `e2e` (whose host buffer is the fp32 tensor the reference's forward takes).
    e2e_u8 = None
    try:
        from yolo_b200 import preprocess
        u8_host = (x_host.permute(0, 2, 3, 1).flip(-1) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
        u8_bufs = [torch.empty(u8_host.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
        x_u8 = [torch.empty_like(x_dev), torch.empty_like(x_dev)]

        def u8_run(steps):
            main_s = torch.cuda.current_stream(dev)
            for b in range(2):
                done[b].record(main_s)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[0])
                u8_bufs[0].copy_(u8_host, non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(steps):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(done[nxt])
                        u8_bufs[nxt].copy_(u8_host, non_blocking=True)
                        ready[nxt].record(copy_stream)
                main_s.wait_event(ready[cur])
                xin, _, _ = preprocess(list(u8_bufs[cur]), IMG, out=x_u8[cur])
                y, _ = model(xin)
                o, c, _k = nms_raw(y.permute(0, 2, 1), CONF, IOU, MAX_DET)
                done[cur].record(main_s)
                out_host.copy_(o, non_blocking=True)
                cnt_host.copy_(c, non_blocking=True)
            torch.cuda.synchronize(dev)

        u8_run(2)
        torch.cuda.synchronize(dev)
        u8_ok = True
    except Exception as e:      # informational only; the headline numbers above do not depend on it
        u8_ok, u8_err = False, f"{type(e).__name__}: {e}"[:200]
    # every rank reaches the collectives below whatever happened above
    sync_all()
    t0 = time.perf_counter()
    if u8_ok:
        try:
            u8_run(args.steps)
        except Exception as e:
            u8_ok, u8_err = False, f"{type(e).__name__}: {e}"[:200]
    tu = torch.tensor([time.perf_counter() - t0 if u8_ok else float("inf")], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tu, op=dist.ReduceOp.MAX)
    if math.isfinite(float(tu.item())):
        e2e_u8 = {"value": world * Bn * args.steps / float(tu.item()), "unit": "images/s", "h2d_bytes_per_step": int(u8_host.numel()),
                  "how": "pinned host uint8 HWC BGR frames -> H2D -> yolo_b200.preprocess (K8) -> YOLO.forward -> nms -> D2H"}
    else:
        e2e_u8 = {"error": u8_err if not u8_ok else "failed on another rank"}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- NMS device time (3 launches: init, filter, select) ----
    y_static, _ = model(x_dev)
    pred_static = y_static.permute(0, 2, 1)
    torch.cuda.synchronize(dev)
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(5):
        nms_raw(pred_static, CONF, IOU, MAX_DET)
    n1.record()
    torch.cuda.synchronize(dev)
    nms_ms = n0.elapsed_time(n1) / 5

    # ---- per-op device timing: roofline of the conv kernels (rank 0) ----
    table = plan.op_table()
    n_ops = len(table)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ops + 1)]
    per_op = [0.0] * n_ops
    reps = 3
    for _ in range(reps):
        torch.cuda.synchronize(dev)
        evs[0].record()
        for i in range(n_ops):
            plan.run_op(i)
            evs[i + 1].record()
        torch.cuda.synchronize(dev)
        for i in range(n_ops):
            per_op[i] += evs[i].elapsed_time(evs[i + 1]) / reps
    # An event record between two kernels costs a few microseconds of idle GPU per interval, so the raw
    # intervals sum to more than one un-instrumented pass over the same ops.  Measure that pass and remove the
    # average excess from every interval: the corrected per-op times sum to the real plan time.
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    p0.record()
    for _ in range(reps):
        for i in range(n_ops):
            plan.run_op(i)
    p1.record()
    torch.cuda.synchronize(dev)
    plain_ms = p0.elapsed_time(p1) / reps
    raw_sum = sum(per_op)
    gap_ms = max(0.0, (raw_sum - plain_ms) / n_ops)
    per_op_raw = list(per_op)
    per_op = [max(t - gap_ms, 0.25 * t) for t in per_op]
    if args.per_op:
        with open(args.per_op, "w") as f:
            f.write("op,kernel,shape,ms,gflop,tflops,ms_raw\n")
            for i, ((name, fl), tms, desc) in enumerate(zip(table, per_op, plan.op_descriptions())):
                f.write(f"{i},{name},{desc},{tms:.5f},{fl / 1e9:.3f},{(fl / (tms / 1e3) / 1e12) if tms > 0 else 0:.1f},{per_op_raw[i]:.5f}\n")
    peak_tf, peak_gbs, peak_src = peaks()
    fam = {}
    for (name, fl), tms in zip(table, per_op):
        f = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "launches": 0})
        f["ms"] += tms; f["flops"] += fl; f["launches"] += 1
    conv_name = "conv_tc" if "conv_tc" in fam else "conv_ffma"
    cf = fam[conv_name]
    achieved = cf["flops"] / (cf["ms"] / 1e3) / 1e12 if cf["ms"] > 0 else 0.0
    all_conv_flops = sum(f["flops"] for n, f in fam.items() if n.startswith("conv") or n == "stem")
    # DRAM bytes of the conv launches of one step, from the committed ncu launch list of this same command
    # (profiles/*_ncu_launch_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum per kernel family)
    traffic, traffic_src = None, None
    try:
        import csv
        summ = sorted((ROOT / "profiles").glob("r*_ncu_launch_summary.csv"))
        if summ and Bn == PER_GPU_BATCH:
            rows = [r for r in csv.DictReader(open(summ[-1])) if r["kernel"].startswith("conv")]
            traffic = sum(float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in rows) * 1e6
            traffic_src = f"{summ[-1].name}: {sum(int(r['launches_per_step']) for r in rows)} conv launches of one step"
    except Exception:
        traffic, traffic_src = None, None
    roofline = {"bound": "tensor", "kernel": conv_name, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": traffic,
                "traffic_unit": "bytes per step (all conv launches; algorithmic unfused bf16 bytes = 2(|in|+|out|+|w|) = 28.9e9)",
                "traffic_source": traffic_src,
                "launches_per_step": cf["launches"], "ms_per_step": cf["ms"], "flops_per_step": cf["flops"],
                "timing": f"per-launch CUDA-event intervals minus the measured event-record gap ({gap_ms * 1e3:.1f} us/interval); "
                          f"all ops: raw {raw_sum:.3f} ms, un-instrumented pass {plain_ms:.3f} ms",
                "flops_per_image_folded_graph": all_conv_flops / Bn}
    stage_ms = {n: round(f["ms"], 4) for n, f in fam.items()}
    stage_ms["nms"] = round(nms_ms, 4)
    # HBM-bound stages against the measured copy bandwidth (algorithmic bytes)
    A = sum((IMG // s_) ** 2 for s_ in (8, 16, 32))
    dec_bytes = Bn * A * (144 * 4 + 84 * 4)
    hbm = {"dfl_decode_score": {"bytes": dec_bytes, "GBps": dec_bytes / (fam["dfl_decode_score"]["ms"] / 1e3) / 1e9,
                                "frac_of_measured_hbm": dec_bytes / (fam["dfl_decode_score"]["ms"] / 1e3) / 1e9 / peak_gbs}}

    # ---- CPU baseline: oracle port on the host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                  # the CPU arm gets every host core again
        torch.set_num_threads(os.cpu_count() or 1)
        xs = x_host[:2].clone()
        cpu_reference_step(nodes, nc, sd, xs)
        t0 = time.perf_counter()
        n_runs = 3
        for _ in range(n_runs):
            ref_dets = cpu_reference_step(nodes, nc, sd, xs)
        cdt = (time.perf_counter() - t0) / n_runs
        cpu = {"value": 2 / cdt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"2 images x {n_runs} runs of the oracle port (reference forward + NMS, CPU fp32)"}

    print(json.dumps({
        "metric": "gelan-c 640x640 images/sec incl. NMS", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "gelan-c inference 640x640 + DFL decode + NMS(conf=0.25, iou=0.45), calibrated random-init weights",
                   "per_gpu_batch": Bn, "global_batch": Bn * world, "l2": "input batch (315 MB) and activations exceed the 126 MB L2",
                   "detections_per_image": dets_per_img, "parallelism": f"image-sharded x{world}, no data-path collective"},
        "e2e_u8_input": e2e_u8,
        "host_cores_bound": (len(numa_cpus) if numa_cpus else None),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "pinned host fp32 batch -> H2D (copy stream, double-buffered) -> YOLO.forward -> nms -> D2H detections"},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "tcgen05_convs_per_step": plan.num_tcgen05,
        "roofline": roofline, "stage_ms_per_step": stage_ms, "hbm_stages": hbm, "cpu_baseline": cpu, "clocks": clocks,
    }))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
