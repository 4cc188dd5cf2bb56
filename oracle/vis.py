"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of row f3, the explanation output path.

``test.py:33-35`` of the reference re-opens each ``slot_{id}.png`` written by the vis branch
(``sloter/utils/slot_attention.py:68-83``) and up-samples it to the input image's size with
``Image.resize(image_raw.size, resample=Image.BILINEAR)``; ``test.py:40-44`` computes the attention ratio
``sum(map) / (h*w*255)`` on the un-resized map.

The resampling arithmetic lives in a third-party dependency that is not under /root/reference: **Pillow**
(pinned ``Pillow==7.2.0``, requirements.txt:14; ``src/libImaging/Resample.c``).  Its published algorithm for
8-bit single-band images, restated here in numpy:

* per axis, ``precompute_coeffs``: ``scale = in/out``; ``filterscale = max(scale, 1)``;
  ``support = 1.0 * filterscale`` (triangle filter); for output index ``xx``: ``center = (xx+0.5)*scale``,
  ``xmin = max(0, int(center - support + 0.5))``, ``xmax = min(in, int(center + support + 0.5)) - xmin`` taps with
  weights ``tri((x + xmin - center + 0.5) / filterscale)`` normalised to sum 1 (all in double);
* ``normalize_coeffs_8bpc``: fixed point with 22 fractional bits, ``int(0.5 + k * 2^22)`` (``-0.5`` for negatives);
* two separable passes, horizontal first, each ``clip8((2^21 + sum(pixel * k)) >> 22)`` -- the intermediate image is
  uint8, i.e. rounded once between the passes.

Pinned against the real Pillow (12.2.0 in the build container; the 8bpc resampler is unchanged since 3.4) by
``oracle/make_golden_vis.py`` -> ``tests/golden/vis_upsample.npz`` and by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size: int, out_size: int):
    """(bounds (out,2) int32 [xmin, count], coeffs (out, ksize) int32) of one axis -- Resample.c precompute_coeffs +
    normalize_coeffs_8bpc for the bilinear (triangle, support 1.0) filter."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)          # C (int) cast truncates toward zero; operands are >= -0.5
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], np.float64)
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Resample the LAST axis of a uint8 array."""
    out = np.empty(img.shape[:-1] + (bounds.shape[0],), np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + (src[..., xmin:xmin + cnt] * kk[xx, :cnt].astype(np.int64)).sum(-1)
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(maps: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """(..., h, w) uint8 -> (..., out_h, out_w) uint8, bit-identical to ``Image.fromarray(m, 'L').resize((out_w, out_h),
    Image.BILINEAR)`` per map (test.py:35)."""
    maps = np.ascontiguousarray(maps, dtype=np.uint8)
    h, w = maps.shape[-2:]
    t = maps
    if out_w != w:                                                    # ImagingResample: need_horizontal
        t = _pass(t, *bilinear_coeffs(w, out_w))
    if out_h != h:                                                    # need_vertical
        t = np.swapaxes(_pass(np.swapaxes(t, -1, -2), *bilinear_coeffs(h, out_h)), -1, -2)
    return np.ascontiguousarray(t)


def attention_ratio(map_u8: np.ndarray) -> float:
    """test.py:40-44: share of the (un-resized) map that is lit."""
    h, w = map_u8.shape
    return float(map_u8.astype(np.uint64).sum()) / float(h * w * 255)
