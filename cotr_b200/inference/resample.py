"""Coefficient tables of Pillow's 8-bit antialiased bilinear resampler (the resize the reference applies to every crop:
`PIL.Image.fromarray(patch).resize((256, 256), resample=PIL.Image.BILINEAR)`, refinement_task.py:117-118,
inference_helper.py:110-111).

Pillow resamples separably (horizontal pass, then vertical pass) with per-output-pixel integer weights
(`precompute_coeffs` + `normalize_coeffs_8bpc` in libImaging/Resample.c): the filter support is max(scale, 1) source
pixels wide, the weights are normalised in double precision and quantised to 22-bit fixed point, each pass rounds to
uint8.  The tables below are computed with the same double-precision expressions on the host (they depend only on the
crop size, so they are cached) and consumed by the CUDA kernels in csrc/preprocess.cu; `resize_u8` is a numpy
emulation used by the CPU tests to pin the tables bit-exactly against Pillow itself.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OUT_SIZE = 256


def bilinear_coeffs(in_size, out_size=OUT_SIZE):
    """(bounds (out,2) int32 [first source index, count], weights (out,ksize) int32 22-bit fixed point)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size      # box edges are C floats in Pillow
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    weights = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            w = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - w if w < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            k[:xmax] /= ww
        q = np.where(k < 0, (-0.5 + k * (1 << PRECISION_BITS)), (0.5 + k * (1 << PRECISION_BITS)))
        weights[xx] = np.trunc(q).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return bounds, weights


def resize_u8(img, out_size=OUT_SIZE):
    """numpy emulation of Pillow's resize of a square uint8 HWC image to out_size x out_size (bilinear, antialiased)."""
    size = img.shape[0]
    assert img.shape[0] == img.shape[1]
    if size == out_size:
        return img.copy()
    bounds, weights = bilinear_coeffs(size, out_size)
    half = 1 << (PRECISION_BITS - 1)
    src = img.astype(np.int64)
    # horizontal pass over the rows the vertical pass needs (all of them for a full-image box)
    tmp = np.empty((size, out_size, img.shape[2]), dtype=np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = half + np.tensordot(src[:, x0:x0 + n, :], weights[xx, :n].astype(np.int64), axes=([1], [0]))
        tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    tmp = tmp.astype(np.int64)
    out = np.empty((out_size, out_size, img.shape[2]), dtype=np.uint8)
    for yy in range(out_size):
        y0, n = bounds[yy]
        acc = half + np.tensordot(weights[yy, :n].astype(np.int64), tmp[y0:y0 + n], axes=([0], [0]))
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out
