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


def test_uniforms_are_uniform_and_uncorrelated():
    """Statistical sanity of the stream contract: chi-square uniformity, and no correlation between
    neighbouring units, rows (incl. the 4 rows that share one Philox call) or consecutive draw ids."""
    u = uniforms(seed=7, draw=11, rows=512, units=256)
    counts, _ = np.histogram(u, bins=64, range=(0.0, 1.0))
    expected = u.size / 64
    chi2 = float(np.sum((counts - expected) ** 2 / expected))
    assert chi2 < 130.0                                  # 63 dof: mean 63, 130 is far beyond the 99.99th percentile
    def corr(a, b):
        a, b = a.ravel() - a.mean(), b.ravel() - b.mean()
        return float(a @ b / np.sqrt((a @ a) * (b @ b)))
    lim = 5.0 / np.sqrt(u.size)
    assert abs(corr(u[:, :-1], u[:, 1:])) < lim          # neighbouring units
    assert abs(corr(u[:-1], u[1:])) < lim                # neighbouring rows (3 of 4 pairs share a Philox call)
    assert abs(corr(u, uniforms(7, 12, 512, 256))) < lim  # next draw id
    assert abs(corr(u, uniforms(8, 11, 512, 256))) < lim  # next seed


def test_normals_have_unit_moments():
    from tests.philox_ref import normals
    z = normals(seed=3, draw=4, rows=512, units=256)
    n = z.size
    assert abs(z.mean()) < 5.0 / np.sqrt(n) and abs(z.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(z ** 3)) < 5.0 * np.sqrt(15.0 / n) and abs(np.mean(z ** 4) - 3.0) < 5.0 * np.sqrt(96.0 / n)
    # the Box-Muller pair (rows 2r, 2r+1 come from one call) must be uncorrelated
    a, b = z[0::2].ravel(), z[1::2].ravel()
    assert abs(float(np.corrcoef(a, b)[0, 1])) < 5.0 / np.sqrt(a.size)
    assert np.array_equal(normals(3, 4, 256, 256, row0=256), z[256:])   # row offset = same stream shifted
