"""Probe: does running the batch as k independent sub-batches on k streams (one CUDA graph with k branches) hide the
per-kernel ramp/tail?  Forward only (no NMS).  python scripts/streams_probe.py [B]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO
from yolo_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
cfg = ROOT / "configs" / "models" / "gelan-c.yaml"
nodes, nc = G.load_graph(cfg)
sd = G.calibrated_state_dict(nodes, nc)
model = YOLO.from_yaml(cfg); model.load_state_dict(sd, strict=True)
model = model.to(dev).eval().set_precision("bf16")
x = make_inputs(B, 640, seed=7).to(dev)

def bench(k, serial=False):
    parts = [x[i * (B // k):(i + 1) * (B // k)].contiguous() for i in range(k)]
    plans = [E.compile_model(model, p) for p in parts]
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    def run():
        cur = torch.cuda.current_stream(dev)
        if serial or k == 1:
            for p in plans: p.run()
            return
        for s, p in zip(streams, plans):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                p.run()
        for s in streams:
            cur.wait_stream(s)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"k={k} serial={serial}: {ms:.3f} ms per {B} images -> {B / ms * 1e3:.0f} img/s (forward only)", flush=True)
    return plans[0].result[1]

y1 = bench(1)
bench(2, serial=True)
bench(2)
bench(4)
bench(8)
