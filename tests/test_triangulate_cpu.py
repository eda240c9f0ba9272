"""CPU: the test oracle of `triangulate_corr` (oracle/triangulate_oracle.py) against an independent implementation of the
same definition.  The reference renders with OpenGL (COTR/inference/inference_helper.py:293-308), which cannot run here and
for which it ships no vectors, so parity with the reference itself stays unpinned; what CAN be pinned is that the oracle
the CUDA rasteriser is tested against really is "barycentric interpolation of the target points over the Delaunay
triangles of the source points, zero outside the hull" - scipy's LinearNDInterpolator is exactly that, written by
somebody else."""
import numpy as np
import pytest
from scipy.interpolate import LinearNDInterpolator

from oracle import triangulate_oracle


@pytest.mark.parametrize("seed,n,shape", [(0, 40, (96, 128)), (1, 300, (120, 90)), (2, 3, (64, 64))])
def test_oracle_is_barycentric_interpolation_over_delaunay(seed, n, shape):
    rs = np.random.RandomState(seed)
    h, w = shape
    src = np.stack([rs.uniform(0, w, n), rs.uniform(0, h, n)], axis=1)
    dst = np.stack([rs.uniform(0, 200, n), rs.uniform(0, 300, n)], axis=1)
    corr = np.concatenate([src, dst], axis=1)
    dense, inside = triangulate_oracle.triangulate_corr(corr, (h, w), (300, 200))
    assert dense.shape == (h, w, 2) and dense.dtype == np.float32 and inside.shape == (h, w)
    ys, xs = np.mgrid[0:h, 0:w]
    pix = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], axis=1)
    want = LinearNDInterpolator(src, dst)(pix).reshape(h, w, 2)
    hull = ~np.isnan(want[..., 0])
    # the two point-location routines may disagree only for pixel centres lying on the hull boundary itself
    assert (hull != inside).mean() < 2e-3
    both = hull & inside
    assert both.sum() > (10 if n > 3 else 1)
    assert np.abs(dense[both] - want[both]).max() < 1e-3            # fp32 output of values up to 300
    assert (dense[~inside] == 0).all()                               # outside the hull: untouched, like the GL clear colour


def test_oracle_reproduces_vertices_and_affine_maps():
    """Interpolating an affine map of the source points gives that affine map at every covered pixel centre."""
    rs = np.random.RandomState(5)
    h, w = 80, 100
    src = np.concatenate([np.array([[0, 0], [w, 0], [0, h], [w, h]], dtype=np.float64), np.stack([rs.uniform(0, w, 50), rs.uniform(0, h, 50)], axis=1)])
    A = np.array([[0.7, -0.2], [0.1, 1.3]]); t = np.array([5.0, -3.0])
    corr = np.concatenate([src, src @ A.T + t], axis=1)
    dense, inside = triangulate_oracle.triangulate_corr(corr, (h, w), (200, 200))
    assert inside.all()                                              # the four corners span every pixel centre
    ys, xs = np.mgrid[0:h, 0:w]
    pix = np.stack([xs + 0.5, ys + 0.5], axis=-1)
    assert np.abs(dense - (pix @ A.T + t)).max() < 1e-4
