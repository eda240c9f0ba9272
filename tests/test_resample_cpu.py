"""CPU: the coefficient tables behind cotr_preprocess (cotr_b200/inference/resample.py, restated in csrc/preprocess.cu)
against Pillow itself - the resize the reference applies to every crop (refinement_task.py:117-118,
inference_helper.py:110-111).  Bit-exact on uint8: down-scaling (antialiased, wide support), up-scaling, the identity
size and the smallest crops the zoom loop can produce."""
import numpy as np
import PIL.Image
import pytest

from cotr_b200.inference.resample import OUT_SIZE, PRECISION_BITS, bilinear_coeffs, resize_u8

SIZES = [2, 3, 7, 48, 100, 162, 255, 256, 257, 276, 390, 511, 512, 600, 783, 1024]


@pytest.mark.parametrize("size", SIZES)
def test_resize_matches_pillow_bit_exactly(size):
    rs = np.random.RandomState(size)
    img = rs.randint(0, 256, size=(size, size, 3), dtype=np.uint8)
    ref = np.array(PIL.Image.fromarray(img).resize((OUT_SIZE, OUT_SIZE), resample=PIL.Image.BILINEAR))
    got = resize_u8(img)
    assert got.dtype == np.uint8 and got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_extreme_images_saturate_like_pillow():
    for value in (0, 255):
        img = np.full((333, 333, 3), value, dtype=np.uint8)
        ref = np.array(PIL.Image.fromarray(img).resize((OUT_SIZE, OUT_SIZE), resample=PIL.Image.BILINEAR))
        assert np.array_equal(resize_u8(img), ref)
    # a checkerboard stresses the rounding of the intermediate uint8 pass
    yy, xx = np.mgrid[0:301, 0:301]
    img = np.repeat((((yy + xx) & 1) * 255).astype(np.uint8)[..., None], 3, axis=2)
    ref = np.array(PIL.Image.fromarray(img).resize((OUT_SIZE, OUT_SIZE), resample=PIL.Image.BILINEAR))
    assert np.array_equal(resize_u8(img), ref)


@pytest.mark.parametrize("size", [2, 100, 256, 700])
def test_coefficient_table_invariants(size):
    bounds, weights = bilinear_coeffs(size)
    assert bounds.shape == (OUT_SIZE, 2) and weights.shape[0] == OUT_SIZE
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= size).all() and (bounds[:, 1] >= 1).all()
    # normalised to one in 22-bit fixed point, up to the per-tap rounding
    total = weights.sum(axis=1)
    assert np.abs(total - (1 << PRECISION_BITS)).max() <= weights.shape[1]
    # taps beyond the count are zero (the kernels still bound their loops by the count)
    for xx in range(OUT_SIZE):
        assert (weights[xx, bounds[xx, 1]:] == 0).all()
