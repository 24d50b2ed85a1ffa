"""Probe (GPU box): NMS of batch i on a second stream beside the forward of batch i+1 (two plans = double-buffered outputs)\nagainst the plain single-stream loop.  Result: profiles/r02_notes.md section 3 (1 %, not adopted)."""
import sys, os, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO, nms_raw
dev = torch.device("cuda", 0)
cfg = ROOT / "configs" / "models" / "gelan-c.yaml"
nodes, nc = G.load_graph(cfg); sd = G.calibrated_state_dict(nodes, nc)
def mk():
    m = YOLO.from_yaml(cfg); m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    m.check_weights = False; m.fresh_outputs = False; m.use_cuda_graph = True
    return m
models = [mk(), mk()]
x = make_inputs(64, 640, seed=7).to(dev)
side = torch.cuda.Stream(dev)
def base(steps):
    for i in range(steps):
        y, _ = models[0](x)
        nms_raw(y.permute(0, 2, 1), 0.25, 0.45, 300)
def overlap(steps):
    cur = torch.cuda.current_stream(dev)
    done = [None, None]
    for i in range(steps):
        s = i & 1
        if done[s] is not None: cur.wait_event(done[s])     # the slot's previous NMS has read its y
        y, _ = models[s](x)
        ev = torch.cuda.Event(); ev.record(cur)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            nms_raw(y.permute(0, 2, 1), 0.25, 0.45, 300)
            d = torch.cuda.Event(); d.record(side); done[s] = d
    cur.wait_stream(side)
def timeit(f, steps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(steps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
base(4); overlap(4)
for name, f in [("base", base), ("overlap", overlap)] * 3:
    print(name, "%.3f ms/step" % timeit(f), flush=True)
