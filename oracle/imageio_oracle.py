"""TEST INFRASTRUCTURE ONLY - CPU restatement of the image output path.

* ``quantize``: the pixels of ``I(tensor).pil()`` for a float tensor - ``TF.to_pil_image(t.float().clamp(0,1))``
  (reference ``_util/twodee_v1.py:184``), i.e. torchvision's ``pic.mul(255).byte()`` laid out (H,W,C).  torchvision and
  opencv are not installed in this image, so ``_util/twodee_v1.py`` cannot be imported to generate fixtures: **parity
  unpinned** for the quantisation rule (it follows torchvision 0.12's ``to_pil_image``, the version the reference's
  Dockerfile pins next to torch 1.11); the PNG container itself is pinned by decoding every file with PIL.
* ``xyza``: the 4-channel image of ``_scripts/eval/generate.py:141-144``.
* ``png_filter_rows``: the five PNG filters (PNG specification 1.2, section 6) + libpng's minimum-sum-of-absolute-
  differences choice, byte for byte what ``p3d_image_to_png_scanlines`` must produce.
* ``png_unfilter``: inverse, for round-trip tests without a decoder.

Only ``tests/`` may import this module; the product path is ``panic3d_b200.imageio``."""
from __future__ import annotations

import numpy as np
import torch


def quantize(img):
    """(C,H,W) or (N,C,H,W) float tensor -> (N,H,W,C) uint8."""
    if img.dim() == 3:
        img = img[None]
    return img.float().clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous().numpy()


def xyza(image_xyz, image_weights, bw):
    """generate.py:141-144."""
    return torch.cat([(image_xyz + bw / 2) / bw, image_weights], dim=1)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def png_filter_rows(pix):
    """(H,W,C) uint8 -> (H, 1 + W*C) uint8: filter type byte + filtered row, filter chosen per row."""
    H, W, C = pix.shape
    raw = pix.reshape(H, W * C).astype(np.int32)
    out = np.zeros((H, 1 + W * C), dtype=np.uint8)
    zero = np.zeros(W * C, dtype=np.int32)
    for y in range(H):
        x = raw[y]
        b = raw[y - 1] if y > 0 else zero
        a = np.concatenate([np.zeros(C, np.int32), x[:-C]])
        c = np.concatenate([np.zeros(C, np.int32), b[:-C]])
        cand = [x, x - a, x - b, x - ((a + b) >> 1), x - _paeth(a, b, c)]
        cand = [v & 255 for v in cand]
        sums = [int(np.where(v < 128, v, 256 - v).sum()) for v in cand]
        t = int(np.argmin(sums))                                              # first minimum
        out[y, 0] = t
        out[y, 1:] = cand[t]
    return out


def png_unfilter(scan, H, W, C):
    """(H, 1+W*C) filtered rows -> (H,W,C) uint8 (PNG specification 1.2, section 6.6)."""
    scan = np.asarray(scan, dtype=np.uint8).reshape(H, 1 + W * C)
    out = np.zeros((H, W * C), dtype=np.int32)
    for y in range(H):
        t, f = int(scan[y, 0]), scan[y, 1:].astype(np.int32)
        b = out[y - 1] if y > 0 else np.zeros(W * C, np.int32)
        row = out[y]
        for i in range(W * C):
            a = row[i - C] if i >= C else 0
            c = b[i - C] if i >= C else 0
            pred = 0 if t == 0 else a if t == 1 else b[i] if t == 2 else (a + b[i]) >> 1 if t == 3 else int(_paeth(np.int32(a), np.int32(b[i]), np.int32(c)))
            row[i] = (f[i] + pred) & 255
    return out.astype(np.uint8).reshape(H, W, C)
