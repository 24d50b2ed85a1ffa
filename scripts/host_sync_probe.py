"""Probe (GPU box): public-API step loop (counts D2H + list slicing, one batch in flight) against the no-host-sync loop,\nalternating, to separate host cost from clock drift.  Result: profiles/r02_notes.md section 3."""
import sys, os, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO, nms_raw, non_max_suppression_async
dev = torch.device("cuda", 0)
cfg = ROOT / "configs" / "models" / "gelan-c.yaml"
nodes, nc = G.load_graph(cfg); sd = G.calibrated_state_dict(nodes, nc)
model = YOLO.from_yaml(cfg); model.load_state_dict(sd, strict=True)
model = model.to(dev).eval().set_precision("bf16")
model.check_weights = False; model.fresh_outputs = False; model.use_cuda_graph = True
x = make_inputs(64, 640, seed=7).to(dev)
def pred():
    y, _ = model(x); return y.permute(0, 2, 1)
def public(steps):
    pend = None
    for _ in range(steps):
        cur = non_max_suppression_async(pred(), 0.25, 0.45, 300)
        if pend is not None: pend.result()
        pend = cur
    return pend.result()
def nosync(steps):
    for _ in range(steps): nms_raw(pred(), 0.25, 0.45, 300)
def fwd_only(steps):
    for _ in range(steps): pred()
def public_nocopy(steps):   # NMS launches + event, no D2H, no result
    evs = []
    for _ in range(steps):
        out = nms_raw(pred(), 0.25, 0.45, 300)
        e = torch.cuda.Event(); e.record(); evs.append(e)
        if len(evs) > 1: evs[-2].synchronize()
def timeit(f, steps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); f(steps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) / steps * 1e3
public(5); nosync(2)
for name, f in [("public", public), ("nosync", nosync), ("public", public), ("nosync", nosync), ("fwd_only", fwd_only), ("public_nocopy", public_nocopy), ("public", public), ("nosync", nosync)]:
    print(name, "%.3f ms/step (events)  %.3f (wall)" % timeit(f), flush=True)
