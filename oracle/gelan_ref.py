"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the
yolo-re detection-inference forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file; the product package
(``yolo-re_b200/yolo_b200``) never does.

What it restates (paths relative to the reference checkout):
  * the YAML graph builder              src/yolo/model/parser.py:65-280
  * the named-DAG interpreter           src/yolo/model/model.py:87-107
  * stride discovery / head bias prior  src/yolo/model/model.py:109-163, src/yolo/heads/detect.py:111-127
  * every block on the path             src/yolo/blocks/{conv,bottleneck,csp,gelan,downsample,sppelan,common,auxiliary}.py
  * the DFL heads                       src/yolo/heads/{detect,dfl,anchor}.py

The reference's arithmetic lives in a third-party dependency that is not under
/root/reference: PyTorch (pinned 2.9.1 in uv.lock:2374; this image has 2.11.0).
The restatement therefore calls the same ATen CPU primitives the reference
dispatches (conv2d, batch_norm, silu, avg/max pool, softmax, sigmoid) through
``torch.nn.functional`` on a flat ``state_dict`` -- there is no nn.Module tree
here, so it is an independent formulation of the graph, not a copy of it.

Pinning: ``tests/test_oracle_vs_reference.py`` (runs only where /root/reference
is mounted) loads the oracle's calibrated state_dict into the real reference
model with ``strict=True`` and requires bit-equal outputs; the committed
fixtures under ``tests/golden/`` were produced by the real reference
(``tests/golden/make_golden.py``) and are checked on every machine.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from pathlib import Path

import torch
import torch.nn.functional as F
import yaml

BN_EPS = 1e-3  # blocks/conv.py:85
REG_MAX = 16   # heads/detect.py:42


# --------------------------------------------------------------------------- graph
@dataclass
class Node:
    name: str
    kind: str
    src: list[str]
    p: dict
    cin: list[int]
    cout: int
    single_src: bool = True


def _round_width(v: int, mult: float, div: int = 8) -> int:
    # parser.py:33-47
    if mult == 1.0:
        return v
    return max(div, int(v * mult + div / 2) // div * div)


def _round_depth(v: int, mult: float) -> int:
    # parser.py:50-62
    return v if mult == 1.0 else max(1, round(v * mult))


def load_graph(cfg_path: str | Path, in_ch: int = 3, num_classes: int | None = None):
    """parser.py:19-30 + 88-122: returns (nodes, num_classes)."""
    data = yaml.safe_load(open(cfg_path))
    m = data.get("model", {})
    nc = num_classes if num_classes is not None else m.get("num_classes", 80)
    wm, dm = m.get("width_multiplier", 1.0), m.get("depth_multiplier", 1.0)
    chan = {"input": in_ch}
    prev = "input"
    nodes: list[Node] = []
    for raw in data.get("layers", []):
        d = dict(raw)
        name, kind = d.pop("name"), d.pop("type")
        frm = d.pop("from", None) or prev
        single = isinstance(frm, str)
        src = [frm] if single else list(frm)
        cin = [chan[s] for s in src]
        if kind in ("DetectDFL", "DualDetectDFL"):
            cout = 0
        elif kind == "Concat":
            cout = sum(cin)
        elif kind in ("Silence", "Upsample"):
            cout = cin[0]
        elif kind == "CBLinear":
            d["out_channels_list"] = [_round_width(c, wm) for c in d["out_channels_list"]]
            cout = d["out_channels_list"][-1]
        elif kind == "CBFuse":
            cout = cin[-1]
        elif kind in ("Conv", "ADown", "RepNCSPELAN4", "SPPELAN"):
            for k in ("out_channels", "hidden_channels", "block_channels"):
                if k in d:
                    d[k] = _round_width(d[k], wm)
            if "num_repeats" in d:
                d["num_repeats"] = _round_depth(d["num_repeats"], dm)
            cout = d["out_channels"]
        else:
            raise ValueError(f"unknown block type {kind}")
        nodes.append(Node(name, kind, src, d, cin, cout, single))
        chan[name] = cout
        prev = name
    return nodes, nc


# --------------------------------------------------------------------------- parameter schema
def _conv_keys(out: OrderedDict, p: str, cin: int, cout: int, k: int, g: int = 1):
    out[f"{p}.conv.weight"] = ((cout, cin // g, k, k), torch.float32)
    out[f"{p}.bn.weight"] = ((cout,), torch.float32)
    out[f"{p}.bn.bias"] = ((cout,), torch.float32)
    out[f"{p}.bn.running_mean"] = ((cout,), torch.float32)
    out[f"{p}.bn.running_var"] = ((cout,), torch.float32)
    out[f"{p}.bn.num_batches_tracked"] = ((), torch.int64)


def _csp_keys(out, p, cin, cout, n):
    h = int(cout * 0.5)                      # csp.py:46
    _conv_keys(out, f"{p}.conv1", cin, h, 1)
    _conv_keys(out, f"{p}.conv2", cin, h, 1)
    _conv_keys(out, f"{p}.conv3", 2 * h, cout, 1)
    for i in range(n):                       # bottleneck.py:44-46 with expansion 1.0
        b = f"{p}.bottlenecks.{i}"
        _conv_keys(out, f"{b}.conv1.conv1", h, h, 3)
        _conv_keys(out, f"{b}.conv1.conv2", h, h, 1)
        _conv_keys(out, f"{b}.conv2", h, h, 3)


def head_widths(ch0: int, nc: int) -> tuple[int, int]:
    # detect.py:45-46
    c2 = math.ceil(max(ch0 // 4, REG_MAX * 4, 16) / 4) * 4
    c3 = max(ch0, min(nc * 2, 128))
    return c2, c3


def _tower_keys(out, pbox, pcls, chs, nc):
    c2, c3 = head_widths(chs[0], nc)
    for i, ch in enumerate(chs):
        _conv_keys(out, f"{pbox}.{i}.0", ch, c2, 3)
        _conv_keys(out, f"{pbox}.{i}.1", c2, c2, 3, 4)
        out[f"{pbox}.{i}.2.weight"] = ((4 * REG_MAX, c2 // 4, 1, 1), torch.float32)
        out[f"{pbox}.{i}.2.bias"] = ((4 * REG_MAX,), torch.float32)
    for i, ch in enumerate(chs):
        _conv_keys(out, f"{pcls}.{i}.0", ch, c3, 3)
        _conv_keys(out, f"{pcls}.{i}.1", c3, c3, 3)
        out[f"{pcls}.{i}.2.weight"] = ((nc, c3, 1, 1), torch.float32)
        out[f"{pcls}.{i}.2.bias"] = ((nc,), torch.float32)


def param_schema(nodes: list[Node], nc: int) -> "OrderedDict[str, tuple]":
    """Every state_dict key the reference model owns, with shape and dtype."""
    out: OrderedDict = OrderedDict()
    for n in nodes:
        p, q = f"layers.{n.name}", n.p
        if n.kind == "Conv":
            _conv_keys(out, p, n.cin[0], n.cout, q.get("kernel_size", 1), q.get("groups", 1))
        elif n.kind == "RepNCSPELAN4":       # gelan.py:46-56
            h, b, r = q["hidden_channels"], q["block_channels"], q.get("num_repeats", 1)
            _conv_keys(out, f"{p}.conv_in", n.cin[0], h, 1)
            _csp_keys(out, f"{p}.block1.0", h // 2, b, r)
            _conv_keys(out, f"{p}.block1.1", b, b, 3)
            _csp_keys(out, f"{p}.block2.0", b, b, r)
            _conv_keys(out, f"{p}.block2.1", b, b, 3)
            _conv_keys(out, f"{p}.conv_out", h + 2 * b, n.cout, 1)
        elif n.kind == "ADown":              # downsample.py:34-38
            _conv_keys(out, f"{p}.conv_stride", n.cin[0] // 2, n.cout // 2, 3)
            _conv_keys(out, f"{p}.conv_pool", n.cin[0] // 2, n.cout // 2, 1)
        elif n.kind == "SPPELAN":            # sppelan.py:35-41
            h = q["hidden_channels"]
            _conv_keys(out, f"{p}.conv_in", n.cin[0], h, 1)
            _conv_keys(out, f"{p}.conv_out", 4 * h, n.cout, 1)
        elif n.kind == "CBLinear":           # auxiliary.py:51-59
            tot = sum(q["out_channels_list"])
            out[f"{p}.conv.weight"] = ((tot, n.cin[0], 1, 1), torch.float32)
            out[f"{p}.conv.bias"] = ((tot,), torch.float32)
        elif n.kind == "DetectDFL":          # detect.py:48-66
            _tower_keys(out, f"{p}.box_convs", f"{p}.cls_convs", n.cin, nc)
            out[f"{p}.dfl.conv.weight"] = ((1, REG_MAX, 1, 1), torch.float32)
        elif n.kind == "DualDetectDFL":      # detect.py:149-190
            L = len(n.cin) // 2
            _tower_keys(out, f"{p}.aux_box_convs", f"{p}.aux_cls_convs", n.cin[:L], nc)
            _tower_keys(out, f"{p}.main_box_convs", f"{p}.main_cls_convs", n.cin[L:], nc)
            out[f"{p}.dfl.conv.weight"] = ((1, REG_MAX, 1, 1), torch.float32)
            out[f"{p}.dfl2.conv.weight"] = ((1, REG_MAX, 1, 1), torch.float32)
    return out


# --------------------------------------------------------------------------- blocks (eval mode)
class _Ctx:
    """Carries the state_dict plus optional train-mode BN calibration."""

    def __init__(self, sd, calibrate=False):
        self.sd = sd
        self.calibrate = calibrate

    def cba(self, p: str, x, stride=1, pad=None, groups=1, act=True):
        """Conv.forward = act(bn(conv(x)))  -- blocks/conv.py:88-89; pad = k//2 (conv.py:12-21)."""
        w = self.sd[f"{p}.conv.weight"]
        k = w.shape[-1]
        y = F.conv2d(x, w, None, stride, k // 2 if pad is None else pad, 1, groups)
        g, b = self.sd[f"{p}.bn.weight"], self.sd[f"{p}.bn.bias"]
        if self.calibrate:
            # BatchNorm2d in train mode with momentum=1.0: running := batch stats
            # (unbiased variance is what torch stores), output normalised with the biased one.
            mean = y.mean(dim=(0, 2, 3))
            n = y.numel() // y.shape[1]
            var_b = y.var(dim=(0, 2, 3), unbiased=False)
            self.sd[f"{p}.bn.running_mean"] = mean.clone()
            self.sd[f"{p}.bn.running_var"] = var_b * (n / max(n - 1, 1))
            self.sd[f"{p}.bn.num_batches_tracked"] = self.sd[f"{p}.bn.num_batches_tracked"] + 1
            y = F.batch_norm(y, None, None, g, b, True, 0.0, BN_EPS)
        else:
            y = F.batch_norm(y, self.sd[f"{p}.bn.running_mean"], self.sd[f"{p}.bn.running_var"],
                             g, b, False, 0.0, BN_EPS)
        return F.silu(y) if act else y

    def repconv(self, p, x):
        # conv.py:140-141: act(conv3x3_bn(x) + conv1x1_bn(x)); the branches are never fused upstream
        return F.silu(self.cba(f"{p}.conv1", x, act=False) + self.cba(f"{p}.conv2", x, pad=0, act=False))

    def bottleneck(self, p, x):
        # bottleneck.py:47-51 (shortcut=True and in==out inside RepNCSP, csp.py:52-54)
        return x + self.cba(f"{p}.conv2", self.repconv(f"{p}.conv1", x))

    def csp(self, p, x):
        # csp.py:59-60
        t = self.cba(f"{p}.conv1", x)
        i = 0
        while f"{p}.bottlenecks.{i}.conv2.conv.weight" in self.sd:
            t = self.bottleneck(f"{p}.bottlenecks.{i}", t)
            i += 1
        return self.cba(f"{p}.conv3", torch.cat((t, self.cba(f"{p}.conv2", x)), 1))

    def elan(self, p, x):
        # gelan.py:58-62
        y = list(self.cba(f"{p}.conv_in", x).chunk(2, 1))
        y.append(self.cba(f"{p}.block1.1", self.csp(f"{p}.block1.0", y[-1])))
        y.append(self.cba(f"{p}.block2.1", self.csp(f"{p}.block2.0", y[-1])))
        return self.cba(f"{p}.conv_out", torch.cat(y, 1))

    def adown(self, p, x):
        # downsample.py:40-46
        x = F.avg_pool2d(x, 2, 1, 0, ceil_mode=True)
        a, b = x.chunk(2, 1)
        a = self.cba(f"{p}.conv_stride", a, stride=2, pad=1)
        b = self.cba(f"{p}.conv_pool", F.max_pool2d(b, 3, 2, 1), pad=0)
        return torch.cat((a, b), 1)

    def sppelan(self, p, x):
        # sppelan.py:43-48
        y = [self.cba(f"{p}.conv_in", x)]
        for _ in range(3):
            y.append(F.max_pool2d(y[-1], 5, 1, 2))
        return self.cba(f"{p}.conv_out", torch.cat(y, 1))

    def tower(self, pbox, pcls, i, x):
        # detect.py:87-88 with towers from detect.py:48-64
        b = self.cba(f"{pbox}.{i}.1", self.cba(f"{pbox}.{i}.0", x), groups=4)
        b = F.conv2d(b, self.sd[f"{pbox}.{i}.2.weight"], self.sd[f"{pbox}.{i}.2.bias"], groups=4)
        c = self.cba(f"{pcls}.{i}.1", self.cba(f"{pcls}.{i}.0", x))
        c = F.conv2d(c, self.sd[f"{pcls}.{i}.2.weight"], self.sd[f"{pcls}.{i}.2.bias"])
        return torch.cat((b, c), 1)


def anchors_and_strides(shapes: list[tuple[int, int]], strides, dtype):
    """make_anchors -- heads/anchor.py:26-40 (offset 0.5, ij meshgrid, (x, y) order)."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=dtype) + 0.5
        sy = torch.arange(h, dtype=dtype) + 0.5
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=dtype))
    return torch.cat(pts), torch.cat(st)


def dfl_expectation(box: torch.Tensor, w_dfl: torch.Tensor) -> torch.Tensor:
    """DFL.forward -- heads/dfl.py:46-50: [B,64,A] -> softmax over the 16 bins -> 1x1 conv -> [B,4,A]."""
    b, _, a = box.shape
    prob = box.view(b, 4, REG_MAX, a).transpose(2, 1).softmax(1)
    return F.conv2d(prob, w_dfl).view(b, 4, a)


def decode(raws: list[torch.Tensor], strides, nc: int, w_dfl: torch.Tensor) -> torch.Tensor:
    """DetectDFL.forward tail -- heads/detect.py:93-108 and dist2bbox heads/anchor.py:57-64."""
    b = raws[0].shape[0]
    flat = torch.cat([r.reshape(b, 4 * REG_MAX + nc, -1) for r in raws], 2)
    box, cls = flat.split((4 * REG_MAX, nc), 1)
    pts, st = anchors_and_strides([tuple(r.shape[2:]) for r in raws], strides, raws[0].dtype)
    pts, st = pts.transpose(0, 1).unsqueeze(0), st.transpose(0, 1)
    lt, rb = torch.split(dfl_expectation(box, w_dfl), 2, 1)
    x1y1, x2y2 = pts - lt, pts + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


def graph_strides(nodes: list[Node]) -> dict[str, int]:
    """Down-sampling factor of every node output relative to the input; equals what
    init_stride measures with its 256x256 dummy forward (model/model.py:109-163)."""
    s = {"input": 1}
    for n in nodes:
        base = s[n.src[-1]] if n.kind == "CBFuse" else s[n.src[0]]
        if n.kind == "Conv":
            base *= n.p.get("stride", 1)
        elif n.kind == "ADown":
            base *= 2
        elif n.kind == "Upsample":
            base //= n.p.get("scale_factor", 2)
        s[n.name] = base
    return s


def detect_strides(nodes: list[Node]) -> list[float]:
    s = graph_strides(nodes)
    det = nodes[-1]
    src = det.src[len(det.src) // 2:] if det.kind == "DualDetectDFL" else det.src
    return [float(s[x]) for x in src]


@torch.no_grad()
def forward(nodes, nc, sd, x, *, calibrate=False, capture: dict | None = None, train: bool = False):
    """YOLO.forward -- model/model.py:87-107.  Returns the detect layer's eval output:
    ``(y, raws)`` for DetectDFL, ``([y_aux, y_main], [raws_aux, raws_main])`` for DualDetectDFL;
    with ``train=True`` (or ``calibrate``) just the raw per-level maps."""
    cx = _Ctx(sd, calibrate)
    outs = {"input": x}
    strides = detect_strides(nodes)
    result = None
    for n in nodes:
        p = f"layers.{n.name}"
        ins = [outs[s] for s in n.src]
        a = ins[0]
        if n.kind == "Conv":
            o = cx.cba(p, a, stride=n.p.get("stride", 1), pad=n.p.get("padding"), groups=n.p.get("groups", 1),
                       act=n.p.get("activation", "silu") == "silu")
        elif n.kind == "RepNCSPELAN4":
            o = cx.elan(p, a)
        elif n.kind == "ADown":
            o = cx.adown(p, a)
        elif n.kind == "SPPELAN":
            o = cx.sppelan(p, a)
        elif n.kind == "Upsample":             # parser.py:159-171 (nearest)
            o = F.interpolate(a, scale_factor=float(n.p.get("scale_factor", 2)), mode=n.p.get("mode", "nearest"))
        elif n.kind == "Concat":               # common.py:32-33
            o = torch.cat(ins, n.p.get("dimension", 1))
        elif n.kind == "Silence":              # common.py:49-50
            o = a
        elif n.kind == "CBLinear":             # auxiliary.py:61-62
            o = F.conv2d(a, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"]).split(n.p["out_channels_list"], 1)
        elif n.kind == "CBFuse":               # auxiliary.py:100-110
            tgt = ins[-1]
            parts = [F.interpolate(c[n.p["idx"][i]], size=tgt.shape[2:], mode="nearest")
                     for i, c in enumerate(ins[:-1])]
            o = torch.sum(torch.stack([*parts, tgt]), 0)
        elif n.kind == "DetectDFL":
            raws = [cx.tower(f"{p}.box_convs", f"{p}.cls_convs", i, f) for i, f in enumerate(ins)]
            o = raws if (train or calibrate) else (decode(raws, strides, nc, sd[f"{p}.dfl.conv.weight"]), raws)
        elif n.kind == "DualDetectDFL":
            L = len(ins) // 2
            ra = [cx.tower(f"{p}.aux_box_convs", f"{p}.aux_cls_convs", i, ins[i]) for i in range(L)]
            rm = [cx.tower(f"{p}.main_box_convs", f"{p}.main_cls_convs", i, ins[L + i]) for i in range(L)]
            if train or calibrate:
                o = [ra, rm]
            else:
                o = ([decode(ra, strides, nc, sd[f"{p}.dfl.conv.weight"]),
                      decode(rm, strides, nc, sd[f"{p}.dfl2.conv.weight"])], [ra, rm])
        else:
            raise ValueError(n.kind)
        outs[n.name] = o
        if capture is not None and isinstance(o, torch.Tensor):
            capture[n.name] = o
        result = o
    return result


# --------------------------------------------------------------------------- weights
def default_state_dict(nodes, nc, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """A reference-compatible state_dict with torch's default-style init: conv weights
    U(+-1/sqrt(fan_in)), identity BN, head priors of detect.py:111-127, DFL weight arange(16)."""
    g = torch.Generator().manual_seed(seed)
    sd: OrderedDict = OrderedDict()
    strides = detect_strides(nodes)
    for key, (shape, dt) in param_schema(nodes, nc).items():
        leaf = key.rsplit(".", 1)[1]
        if key.endswith("dfl.conv.weight") or key.endswith("dfl2.conv.weight"):
            t = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
        elif dt == torch.int64:
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "weight" and len(shape) == 4:
            bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif leaf in ("weight", "running_var"):
            t = torch.ones(shape)
        elif leaf == "bias" and "_convs." in key:
            lvl = int(key.split("_convs.")[1].split(".")[0])
            if "box_convs" in key:
                t = torch.ones(shape)
            else:
                t = torch.full(shape, math.log(5 / nc / (640 / strides[lvl]) ** 2))
        elif leaf == "bias" and key.endswith("conv.bias"):      # CBLinear
            bound = 1.0 / math.sqrt(sd[key[:-4] + "weight"].shape[1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = torch.zeros(shape)
        sd[key] = t
    return sd


def fractal(b: int, s: int, g: torch.Generator, w: int | None = None) -> torch.Tensor:
    """Multi-octave noise image batch in [0,1] -- the parity input."""
    w = w or s
    x = torch.zeros(b, 3, s, w)
    amp, tot, r = 1.0, 0.0, min(s, w)
    while r >= 5:
        x += amp * F.interpolate(torch.rand(b, 3, r, r, generator=g), size=(s, w), mode="bilinear",
                                 align_corners=False)
        tot += amp
        r //= 2
        amp *= 1.6
    x /= tot
    lo, hi = x.amin(dim=(1, 2, 3), keepdim=True), x.amax(dim=(1, 2, 3), keepdim=True)
    return (x - lo) / (hi - lo)


def calibrated_state_dict(nodes, nc, *, seed: int = 1234, cal_size: int = 640, cal_batch: int = 4,
                          cls_bias: float = -2.2, cls_gain: float = 1.0, box_gain: float = 1.0):
    """Calibrated random init: literal default init is
    numerically degenerate, so BN affine terms are randomised, running stats are taken from
    one train-mode pass over a calibration batch, and the head-final convs are re-drawn."""
    g = torch.Generator().manual_seed(seed)
    sd = default_state_dict(nodes, nc, seed=seed + 1)
    for k in sd:
        if k.endswith(".bn.weight"):
            sd[k] = torch.empty_like(sd[k]).uniform_(0.8, 1.2, generator=g)
        elif k.endswith(".bn.bias"):
            sd[k] = torch.empty_like(sd[k]).normal_(0, 0.2, generator=g)
    xcal = fractal(cal_batch, cal_size, torch.Generator().manual_seed(99))
    forward(nodes, nc, sd, xcal, calibrate=True)
    for k in list(sd):
        if "_convs." in k and k.endswith(".2.weight"):
            w = sd[k]
            gain = box_gain if "box_convs" in k else cls_gain
            sd[k] = torch.empty_like(w).normal_(0, gain / math.sqrt(w[0].numel()), generator=g)
            bk = k[:-6] + "bias"
            if "box_convs" in k:
                sd[bk] = torch.empty_like(sd[bk]).normal_(0, 0.5, generator=g)
            else:
                sd[bk] = torch.full_like(sd[bk], cls_bias)
    return sd
