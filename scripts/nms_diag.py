"""GPU-box diagnostic: where the NMS select time goes at config 5 (1280x1280, conf .001) and config 2.
Prints candidates per image, how deep into the sorted list the scan goes before max_det boxes are kept, and the
device time of nms_raw at max_det = 300 / 1 (== sort + one chunk)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "yolo-re_b200"))
import torch
from bench_data import make_inputs
from oracle import gelan_ref as G
from yolo_b200 import YOLO, nms_raw

dev = "cuda"
cfg = ROOT / "configs/models/gelan-c.yaml"
nodes, nc = G.load_graph(cfg)
sd = G.calibrated_state_dict(nodes, nc)
m = YOLO.from_yaml(cfg); m.load_state_dict(sd); m = m.to(dev).eval()


def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for img, Bn, conf, iou in ((1280, 16, 0.001, 0.6), (640, 64, 0.25, 0.45)):
    x = make_inputs(Bn, img).to(dev)
    y, _ = m(x)
    pred = y.permute(0, 2, 1).contiguous()
    sc, _ = pred[..., 4:].max(-1)
    ncand = (sc > conf).sum(1)
    out, cnt, keep = nms_raw(pred, conf, iou, 300)
    torch.cuda.synchronize()
    depth = []
    for b in range(Bn):
        k = int(cnt[b])
        if k == 0:
            depth.append(0); continue
        last = int(keep[b, k - 1])
        s_last = sc[b, last]
        depth.append(int(((sc[b] > s_last) & (sc[b] > conf)).sum()) + 1)
    print(f"{img}: candidates/img min {int(ncand.min())} mean {float(ncand.float().mean()):.0f} max {int(ncand.max())}; kept mean {float(cnt.float().mean()):.0f}; "
          f"scan depth (rank of the last kept box) min {min(depth)} mean {sum(depth) / len(depth):.0f} max {max(depth)}")
    for md in (300, 100, 1):
        print(f"   nms_raw max_det={md}: {t_ms(lambda: nms_raw(pred, conf, iou, md)):.3f} ms")
    import ctypes as C
    from yolo_b200 import _lib as L
    lib = L.lib()
    if hasattr(lib, "yre_debug_nms_prof"):
        nms_raw(pred, conf, iou, 300)
        buf = (C.c_longlong * 16)()
        lib.yre_debug_nms_prof(buf)
        t = list(buf)
        names = ["sort", "load chunk0", "vs kept", "mask", "resolve", "publish", "rest"]
        print("   image 0 (n=%d) select phases [SM cycles]: " % int(ncand[0]) + ", ".join(f"{nm} {t[i + 1] - t[i]}" for i, nm in enumerate(names)) + f"; resolve loop alone {t[8] - t[4]} cycles for {t[9]} keeps of {t[10]} candidates")
    print(f"   nms_raw conf=0.9 (no candidates): {t_ms(lambda: nms_raw(pred, 0.9999, iou, 300)):.3f} ms")
