"""Oracle of the input pipeline (oracle/data.py) against hand-checked cases (CPU)."""
import numpy as np

from oracle import data as OD


def test_preprocess_image_known_answers():
    img = np.arange(2 * 3 * 1, dtype=np.float32).reshape(2, 3, 1) + 1          # [[1,2,3],[4,5,6]]
    # crop at the pad offset, no flip: identity
    assert np.array_equal(OD.preprocess_image(img, False, 1, 1, pad=1), img)
    # flip reverses the width axis
    assert np.array_equal(OD.preprocess_image(img, True, 1, 1, pad=1)[..., 0], [[3, 2, 1], [6, 5, 4]])
    # crop at (0, 0): first row / column come from the zero padding
    assert np.array_equal(OD.preprocess_image(img, False, 0, 0, pad=1)[..., 0], [[0, 0, 0], [0, 1, 2]])
    # crop at (2, 2): last row / column are padding
    assert np.array_equal(OD.preprocess_image(img, False, 2, 2, pad=1)[..., 0], [[5, 6, 0], [0, 0, 0]])


def test_normalise_is_float64_then_one_rounding():
    X = np.array([[[[0], [255]]]], dtype=np.uint8)
    mean = np.array([[[1.0 / 3], [127.5]]])
    got = OD.normalise(X, mean)
    assert got.dtype == np.float32
    assert got[0, 0, 0, 0] == np.float32((0.0 - 1.0 / 3) / 128) and got[0, 0, 1, 0] == np.float32(127.5 / 128)


def test_philox_params_ranges_and_determinism():
    p = OD.philox_params(1000, 7, (3 << 32) | 0x7FFF0000)
    assert p.shape == (1000, 3) and set(np.unique(p[:, 0])) == {0, 1}
    assert p[:, 1:].min() == 0 and p[:, 1:].max() == 8
    assert np.array_equal(p, OD.philox_params(1000, 7, (3 << 32) | 0x7FFF0000))
    assert not np.array_equal(p, OD.philox_params(1000, 7, (4 << 32) | 0x7FFF0000))
