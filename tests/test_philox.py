"""The numpy Philox used by the oracle against the published Random123 known-answer vectors."""
import numpy as np

from oracle import philox


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    r = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = philox.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_noise_range_and_determinism():
    u = philox.noise(1001, seed=7, offset=philox.make_offset(3, 5))
    assert u.dtype == np.float32 and u.shape == (1001,)
    assert u.min() >= 0.0 and u.max() <= np.float32(1.0 - 2.0 ** -24)
    np.testing.assert_array_equal(u, philox.noise(1001, 7, philox.make_offset(3, 5)))
    assert not np.array_equal(u, philox.noise(1001, 7, philox.make_offset(3, 6)))
    assert abs(float(u.mean()) - 0.5) < 0.05
