"""Known-answer tests for the Philox4x32-10 restatement (Random123 kat_vectors)."""
import numpy as np

from tests.philox_ref import philox4x32_10, uniforms


def _kat(ctr, key):
    return [int(x) for x in philox4x32_10(*ctr, *key)]


def test_philox_known_answers():
    assert _kat((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_contract_properties():
    u = uniforms(seed=123, draw=5, rows=64, units=33)
    assert u.shape == (64, 33) and u.min() >= 0.0 and u.max() < 1.0
    assert np.array_equal(u, u.astype(np.float32).astype(np.float64))          # fp32-representable
    # row offset = same stream shifted (independence of the number of GPUs)
    assert np.array_equal(uniforms(123, 5, 32, 33, row0=32), u[32:])
    assert abs(u.mean() - 0.5) < 0.02
