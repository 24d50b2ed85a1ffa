"""Probe (GPU box): run the launch plan as a DAG on several streams (dependencies derived from the buffers each op reads
and writes) instead of one stream.  Interesting for yolov9-c at 16 images per GPU: the auxiliary branch is independent of
the main neck / head and most of its launches have fewer CTAs than the GPU has SMs.
python scripts/dag_streams_probe.py [config: 4|2|5] [streams]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO
from yolo_b200 import engine as E

cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name, img, B = {2: ("gelan-c", 640, 64), 4: ("yolov9-c", 640, 16), 5: ("gelan-c", 1280, 16)}[cfgno]
dev = torch.device("cuda", 0)
cfg = ROOT / "configs" / "models" / f"{name}.yaml"
nodes, nc = G.load_graph(cfg); sd = G.calibrated_state_dict(nodes, nc)
model = YOLO.from_yaml(cfg); model.load_state_dict(sd, strict=True)
model = model.to(dev).eval().set_precision("bf16")
x = make_inputs(B, img, seed=7).to(dev)
p = E.compile_model(model, x)
n = len(p.trace)

def region(v):      # (buffer id, first channel, last channel)
    return (v.t.data_ptr(), v.c_off, v.c_off + v.C)
def overlaps(a, b):
    return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]
reads, writes = [], []
for kind, a in p.trace:
    r, w = [], []
    if kind == "conv":
        r.append(region(a["x"]))
        if a.get("xu") is not None: r.append(region(a["xu"]))
        if a["res"] is not None: r.append(region(a["res"]))
        w.append(region(a["y"]))
    elif kind == "stem":
        w.append(region(a["y"]))
    elif kind == "adown":
        r.append(region(a["x"])); w += [region(a["lo"]), region(a["hi"])]
    elif kind == "spp":
        r.append(region(a["x"])); w += [region(a[k]) for k in ("y5", "y9", "y13")]
    elif kind == "upsample":
        r.append(region(a["x"])); w.append(region(a["y"]))
    elif kind == "cbfuse":
        r += [region(s) for s in a["srcs"]] + [region(a["target"])]; w.append(region(a["y"]))
    elif kind == "decode":
        r += [region(v) for v in a["raws"]]; w.append((a["y"].data_ptr(), 0, 1 << 30))
    reads.append(r); writes.append(w)
deps = [set() for _ in range(n)]
for j in range(n):
    for i in range(j):
        if any(overlaps(wr, rd) for wr in writes[i] for rd in reads[j]) or \
           any(overlaps(wr, w2) for wr in writes[i] for w2 in writes[j]) or \
           any(overlaps(rd, w2) for rd in reads[i] for w2 in writes[j]):
            deps[j].add(i)
# transitive reduction is not needed; assign streams greedily: an op goes to the stream of one of its deps if that stream's
# last op is that dep (chain continuation), else to the least recently used stream
stream_of, last_on = [0] * n, [-1] * K
for j in range(n):
    cand = [stream_of[i] for i in deps[j] if last_on[stream_of[i]] == i]
    s = cand[0] if cand else min(range(K), key=lambda q: last_on[q])
    stream_of[j] = s; last_on[s] = j
print("ops per stream:", [stream_of.count(s) for s in range(K)])
streams = [torch.cuda.Stream(dev) for _ in range(K)]

def run_serial():
    for i in range(n): p.run_op(i)
def run_dag():
    cur = torch.cuda.current_stream(dev)
    ev0 = torch.cuda.Event(); ev0.record(cur)
    for s in streams: s.wait_event(ev0)
    done = [None] * n
    for j in range(n):
        s = streams[stream_of[j]]
        for i in deps[j]:
            if stream_of[i] != stream_of[j]: s.wait_event(done[i])
        with torch.cuda.stream(s):
            p.run_op(j)
            e = torch.cuda.Event(); e.record(s); done[j] = e
    for s in streams: cur.wait_stream(s)

def bench(fn, label):
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side): fn()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / 20:.3f} ms per forward", flush=True)
    res = p.result[1]
    return [t.clone() for t in (res if isinstance(res, list) else [res])]
a = bench(run_serial, "serial")
b = bench(run_dag, f"dag on {K} streams")
print("identical:", all(torch.equal(u, v) for u, v in zip(a, b)))
bench(run_serial, "serial"); bench(run_dag, f"dag on {K} streams")
