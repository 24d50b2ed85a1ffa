"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the steps either side of the hot path.

  letterbox + BGR->RGB + HWC->CHW + /255     reference scripts/detect.py:40-71, 223-227
  scale_boxes                                reference scripts/detect.py:74-109

The arithmetic of the resize lives in a third-party dependency that is not under /root/reference:
OpenCV (`cv2.resize(..., interpolation=cv2.INTER_LINEAR)` on uint8, opencv-python 4.13.0 in the build image).
Its published algorithm is restated here (8-bit fixed point, 11-bit coefficients):

  fx = float((dx + 0.5) * (src_w / dst_w) - 0.5); sx = floor(fx); fx -= sx
  horizontally the weight is clamped with the index (sx < 0 -> sx = 0, fx = 0; sx >= w-1 -> sx = w-1, fx = 0),
  vertically only the row indices are clipped, the weights are not;
  a = saturate_cast<short>(w * 2048) (round half to even);  row[dx] = S[sx]*a0 + S[sx+1]*a1   (int32)
  dst = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2
  exact 2x down-scaling in both directions is routed to INTER_AREA: (a + b + c + d + 2) >> 2.

Pinned bit-exactly against cv2 itself in tests/test_oracle.py (when cv2 is importable) and against the fixtures
tests/golden/preproc_cases.npz, which were produced by the REFERENCE's own letterbox / scale_boxes.
"""
from __future__ import annotations

import numpy as np


def _coeffs(dn: int, sn: int, clamp_weight: bool):
    scale = sn / dn
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weight:
        lo, hi = s < 0, s >= sn - 1
        f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        s = np.where(lo, 0, np.where(hi, sn - 1, s))
    i0 = np.clip(s, 0, sn - 1)
    i1 = np.clip(s + 1, 0, sn - 1)
    a0 = np.rint((np.float32(1.0) - f).astype(np.float32) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return i0, i1, a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC."""
    sh, sw = src.shape[:2]
    s = src.astype(np.int32)
    if sw == 2 * dw and sh == 2 * dh:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = _coeffs(dw, sw, True)
    y0, y1, b0, b1 = _coeffs(dh, sh, False)
    rows = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(h: int, w: int, new_shape: int = 640):
    """(new_w, new_h, top, left, ratio, (pad_w, pad_h)) exactly as scripts/detect.py:57-71 computes them."""
    r = min(new_shape / h, new_shape / w)
    nw, nh = int(round(w * r)), int(round(h * r))
    dw, dh = (new_shape - nw) / 2, (new_shape - nh) / 2
    top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
    bottom, right = int(round(dh + 0.1)), int(round(dw + 0.1))
    return nw, nh, top, left, bottom, right, r, (int(dw), int(dh))


def letterbox(img: np.ndarray, new_shape: int = 640, color=(114, 114, 114)):
    """scripts/detect.py:40-71 -> (padded uint8 HWC image, (r, r), (pad_w, pad_h))."""
    h, w = img.shape[:2]
    nw, nh, top, left, bottom, right, r, pad = letterbox_geometry(h, w, new_shape)
    if (w, h) != (nw, nh):
        img = resize_linear_u8(img, nw, nh)
    out = np.empty((nh + top + bottom, nw + left + right, img.shape[2]), np.uint8)
    out[...] = np.asarray(color, np.uint8)
    out[top:top + nh, left:left + nw] = img
    return out, (r, r), pad


def preprocess(img_bgr: np.ndarray, new_shape: int = 640):
    """scripts/detect.py:223-227 -> (float32 CHW RGB in [0,1], ratio, pad)."""
    lb, ratio, pad = letterbox(img_bgr, new_shape)
    chw = np.ascontiguousarray(lb[:, :, ::-1].transpose(2, 0, 1))
    return chw.astype(np.float32) / np.float32(255.0), ratio, pad


def scale_boxes(boxes: np.ndarray, img_shape, orig_shape, ratio_pad=None) -> np.ndarray:
    """scripts/detect.py:74-109 on a float32 [n, 4] xyxy array (returns a new array)."""
    if ratio_pad is None:
        gain = min(img_shape[0] / orig_shape[0], img_shape[1] / orig_shape[1])
        pad = ((img_shape[1] - orig_shape[1] * gain) / 2, (img_shape[0] - orig_shape[0] * gain) / 2)
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    b = boxes.astype(np.float32).copy()
    b[:, [0, 2]] -= np.float32(pad[0])
    b[:, [1, 3]] -= np.float32(pad[1])
    b[:, :4] /= np.float32(gain)
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], np.float32(0), np.float32(orig_shape[1]))
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], np.float32(0), np.float32(orig_shape[0]))
    return b
