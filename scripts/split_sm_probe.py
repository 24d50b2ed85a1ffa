"""Probe (GPU box, tuning build): two half batches on two streams, every persistent conv grid sized for HALF of the SMs
(YRE_TC_SMS=74), the second stream started half a forward later -- so that one half batch is in the HBM-bound early
layers while the other is in the tensor-bound neck / head.  Compared with the whole batch on one stream.
python scripts/split_sm_probe.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO
from yolo_b200 import engine as E

B = 64
dev = torch.device("cuda", 0)
cfg = ROOT / "configs" / "models" / "gelan-c.yaml"
nodes, nc = G.load_graph(cfg); sd = G.calibrated_state_dict(nodes, nc)
model = YOLO.from_yaml(cfg); model.load_state_dict(sd, strict=True)
model = model.to(dev).eval().set_precision("bf16")
x = make_inputs(B, 640, seed=7).to(dev)

def timed(fn, reps=10):
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side): fn()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

os.environ.pop("YRE_TC_SMS", None)
full = E.compile_model(model, x)
t_full = timed(lambda: full.run())
print(f"whole batch, one stream: {t_full:.3f} ms per {B} images", flush=True)

for sms in (74, 96, 148):
    os.environ["YRE_TC_SMS"] = str(sms)
    halves = [x[: B // 2].contiguous(), x[B // 2:].contiguous()]
    # each stream runs TWO consecutive half batches per graph replay (= 2 x B/2 per stream = 2B images per replay in total),
    # stream 1 delayed by running its first half-forward's worth of work later: emulate the steady state of a pipeline
    plans = [[E.compile_model(model, h) for h in halves] for _ in range(2)]
    n_ops = plans[0][0].num_launches
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    def run_pipe(offset_ops):
        cur = torch.cuda.current_stream(dev)
        ev0 = torch.cuda.Event(); ev0.record(cur)
        names = [n for n, _ in plans[0][0].op_table()]
        nops = len(names)
        with torch.cuda.stream(streams[0]):
            streams[0].wait_event(ev0)
            marks = []
            for rep in range(2):
                for i in range(nops):
                    plans[0][rep].run_op(i)
                    if rep == 0 and i == offset_ops:
                        m = torch.cuda.Event(); m.record(streams[0]); marks.append(m)
        with torch.cuda.stream(streams[1]):
            streams[1].wait_event(marks[0])                  # start when stream 0 is `offset_ops` ops into its first half batch
            for rep in range(2):
                for i in range(nops):
                    plans[1][rep].run_op(i)
        cur.wait_stream(streams[0]); cur.wait_stream(streams[1])
    for off in (0, 45, 75):
        t = timed(lambda: run_pipe(off), reps=5)
        # one replay = 4 half batches = 2B images
        print(f"sms={sms} offset={off} ops: {t:.3f} ms per {2 * B} images -> {t / 2:.3f} ms per {B}", flush=True)
