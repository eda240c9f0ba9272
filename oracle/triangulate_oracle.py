"""CPU restatement of `triangulate_corr` (COTR/inference/inference_helper.py:293-308) - TEST ORACLE, not product code.

The reference normalises the correspondences, triangulates the source points with `scipy.spatial.Delaunay` and lets
OpenGL (vispy) rasterise the triangles with the target coordinates as vertex colours, i.e. every pixel centre inside a
triangle receives the barycentric interpolation of its three target points; pixels outside the hull stay 0.  vispy /
GL are not available offline, so the restatement evaluates exactly that definition with scipy's own point location
(`find_simplex` + the simplices' affine transforms) in float64.  "parity unpinned": the reference's tests hold no
vectors for this function and its GL path cannot run here; the definition above is what is checked, and
tests/test_triangulate_cpu.py holds this restatement against scipy's LinearNDInterpolator (an independent
implementation of the same definition) and against exact affine maps.
"""
import numpy as np
from scipy.spatial import Delaunay


def triangulate_corr(corr, from_shape, to_shape):
    corr = np.asarray(corr, dtype=np.float64)
    h, w = from_shape[:2]
    tri = Delaunay(corr[:, :2])
    ys, xs = np.mgrid[0:h, 0:w]
    pix = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], axis=1)
    simplex = tri.find_simplex(pix)
    out = np.zeros((h * w, 2), dtype=np.float32)
    ok = simplex >= 0
    T = tri.transform[simplex[ok]]
    bary2 = np.einsum('nij,nj->ni', T[:, :2], pix[ok] - T[:, 2])
    bary = np.concatenate([bary2, 1 - bary2.sum(axis=1, keepdims=True)], axis=1)
    verts = tri.simplices[simplex[ok]]
    out[ok] = np.einsum('nk,nkc->nc', bary, corr[verts, 2:4]).astype(np.float32)
    return out.reshape(h, w, 2), ok.reshape(h, w)
