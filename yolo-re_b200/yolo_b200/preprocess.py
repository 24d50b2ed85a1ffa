"""Device-side mirror of the steps either side of the detection path.

    letterbox(img, new_shape=640, color=(114, 114, 114))        reference scripts/detect.py:40-71
    preprocess(imgs, new_shape=640)                             reference scripts/detect.py:223-227 (letterbox,
                                                                BGR->RGB, HWC->CHW, .float()/255, batched)
    scale_boxes(boxes, img_shape, orig_shape, ratio_pad=None)   reference scripts/detect.py:74-109

Same names, argument meaning and return values as the reference; the arithmetic runs in libyre's K8/K9 kernels
(bit-exact with cv2.resize(INTER_LINEAR) on uint8 and with torch's fp32 box arithmetic).  Images are uint8 HWC BGR,
either numpy arrays (copied to the current CUDA device) or CUDA uint8 tensors.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _to_device_u8(img) -> torch.Tensor:
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8:
            raise TypeError(f"letterbox: expected a uint8 image, got {img.dtype}")
        img = torch.from_numpy(np.ascontiguousarray(img)).cuda(non_blocking=True)
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8):
        raise L.YreError("letterbox: image must be a uint8 numpy array or a CUDA uint8 tensor (no CPU path)")
    if img.dim() != 3 or img.shape[2] != 3:
        raise ValueError(f"letterbox: expected an HxWx3 image, got {tuple(img.shape)}")
    if img.stride(2) != 1 or img.stride(1) != 3:
        img = img.contiguous()
    return img


def _desc(img: torch.Tensor, new_shape: int, color) -> tuple[L.LetterboxDesc, float, tuple[int, int]]:
    d = L.LetterboxDesc()
    d.src, d.h, d.w, d.row_pitch, d.new_shape = img.data_ptr(), img.shape[0], img.shape[1], img.stride(0), int(new_shape)
    d.color[0], d.color[1], d.color[2] = int(color[0]), int(color[1]), int(color[2])
    r, pw, ph = C.c_double(), C.c_int32(), C.c_int32()
    L.check(L.lib().yre_letterbox_geometry(C.byref(d), C.byref(r), C.byref(pw), C.byref(ph)), "letterbox_geometry")
    return d, r.value, (pw.value, ph.value)


def letterbox(img, new_shape: int = 640, color: tuple[int, int, int] = (114, 114, 114)):
    """Resize and pad to a square: returns (uint8 HWC image, (r, r), (pad_w, pad_h)) like the reference.  A numpy input
    gives a numpy result, a CUDA tensor stays on the device."""
    was_numpy = isinstance(img, np.ndarray)
    src = _to_device_u8(img)
    d, r, pad = _desc(src, new_shape, color)
    out = torch.empty((new_shape, new_shape, 3), dtype=torch.uint8, device=src.device)
    d.out_mode, d.dst = L.LB_U8_HWC, out.data_ptr()
    L.check(L.lib().yre_letterbox_u8(C.byref(d), _stream()), "letterbox_u8")
    return (out.cpu().numpy() if was_numpy else out), (r, r), pad


def preprocess(imgs, new_shape: int = 640, color: tuple[int, int, int] = (114, 114, 114), out: torch.Tensor | None = None,
               dtype: torch.dtype = torch.float32):
    """Fused letterbox + BGR->RGB + HWC->CHW + /255 for a list of images (or one image).

    Returns (x [B, 3, S, S] fp32 on the device, ratios [(r, r)], pads [(pad_w, pad_h)]) -- x is the tensor the
    reference builds at scripts/detect.py:223-227, ready for ``model(x)``.

    ``dtype=torch.uint8`` stops after the letterbox and returns the frames as one uint8 [B, S, S, 3] BGR batch;
    ``model(x_u8)`` then fuses BGR->RGB / HWC->CHW / /255 into its first convolution, so the fp32 image tensor is
    never written (4x less traffic on both sides of it).  Detections are identical either way."""
    if isinstance(imgs, (np.ndarray, torch.Tensor)) and imgs.ndim == 3:
        imgs = [imgs]
    srcs = [_to_device_u8(i) for i in imgs]
    dev = srcs[0].device
    if dtype not in (torch.float32, torch.uint8):
        raise ValueError("preprocess: dtype must be torch.float32 or torch.uint8")
    shape = (len(srcs), 3, new_shape, new_shape) if dtype == torch.float32 else (len(srcs), new_shape, new_shape, 3)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=dev)
    if tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f"preprocess: `out` must be a contiguous CUDA {dtype} tensor of shape {shape}")
    ratios, pads = [], []
    descs = (L.LetterboxDesc * len(srcs))()
    for i, src in enumerate(srcs):
        d, r, pad = _desc(src, new_shape, color)
        d.out_mode, d.dst = (L.LB_F32_CHW if dtype == torch.float32 else L.LB_U8_HWC), out[i].data_ptr()
        descs[i] = d
        ratios.append((r, r)); pads.append(pad)
    L.check(L.lib().yre_letterbox_u8_batch(descs, len(srcs), _stream()), "letterbox_u8_batch")
    return out, ratios, pads


def scale_rows(ratios, pads, orig_shapes, device=None) -> torch.Tensor:
    """fp32 [B, 5] rows (pad_w, pad_h, gain, orig_w, orig_h) for ``nms_raw(..., scale=...)`` /
    ``non_max_suppression_async(..., scale=...)``: the ``ratio_pad`` form of scale_boxes (scripts/detect.py:100-101)
    for every image of a batch, so the box rescale/clip runs inside the NMS output pass."""
    rows = [[float(p[0]), float(p[1]), float(r[0]), float(o[1]), float(o[0])] for r, p, o in zip(ratios, pads, orig_shapes)]
    t = torch.tensor(rows, dtype=torch.float32)
    return t.to(device, non_blocking=True) if device is not None else t


def scale_boxes(boxes: torch.Tensor, img_shape, orig_shape, ratio_pad=None) -> torch.Tensor:
    """Boxes (xyxy, model-input pixels) -> original-image pixels, clipped; IN PLACE like the reference, any row
    stride (``detections[:, :4]`` works)."""
    if ratio_pad is None:
        gain = min(img_shape[0] / orig_shape[0], img_shape[1] / orig_shape[1])
        pad = ((img_shape[1] - orig_shape[1] * gain) / 2, (img_shape[0] - orig_shape[0] * gain) / 2)
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    if not (boxes.is_cuda and boxes.dtype == torch.float32):
        raise L.YreError("scale_boxes: boxes must be a CUDA fp32 tensor (no CPU path)")
    if boxes.dim() != 2 or boxes.shape[1] < 4 or (boxes.shape[0] > 1 and boxes.stride(1) != 1):
        raise ValueError(f"scale_boxes: expected [n, >=4] rows with unit column stride, got {tuple(boxes.shape)}")
    n = boxes.shape[0]
    stride = boxes.stride(0) if n > 1 else max(4, boxes.shape[1])
    L.check(L.lib().yre_scale_boxes(boxes.data_ptr(), n, stride, float(pad[0]), float(pad[1]), float(gain),
                                    float(orig_shape[1]), float(orig_shape[0]), _stream()), "scale_boxes")
    return boxes
