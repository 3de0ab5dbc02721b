"""Device-level building blocks of the C ABI (`bzap_dev_*`, include/bzap.h) against numpy on the
same inputs.  These are the steps the distributed single-block path is made of; here they are
checked alone, on digit distributions that take each ranking mode of the onesweep pass (peer masks
by MATCH.ANY for skewed digits, by votes for many distinct digits per warp) and on sizes beyond the
L2 prefetch distance of 148 tiles."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    from bwt_mtf_huffman_compressor_b200.distributed import GpuBackend
    return GpuBackend()


def _keys(kind, m, rng):
    if kind == "uniform":            # every digit uniform: votes on all eight passes
        return rng.integers(0, 1 << 63, m, dtype=np.int64)
    if kind == "skewed":             # two values per digit: MATCH.ANY
        return (rng.integers(0, 2, m, dtype=np.int64) * 0x0101010101010101) & 0x7FFFFFFFFFFFFFFF
    if kind == "constant":           # all passes trivial
        return np.full(m, 0x1234567890ABCDEF, dtype=np.int64)
    if kind == "rank_pairs":         # (r1 << 32 | r2) with 20-bit ranks: mixed modes, trivial top digits
        return (rng.integers(0, 1 << 20, m, dtype=np.int64) << 32) | rng.integers(0, 1 << 20, m, dtype=np.int64)
    if kind == "few":                # 5 distinct keys
        return rng.choice(np.array([3, 1 << 40, 77, (1 << 62) + 5, 1 << 17], dtype=np.int64), m)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "skewed", "constant", "rank_pairs", "few"])
@pytest.mark.parametrize("m", [1, 31, 4096, 4097, 70001, 1_500_003])
def test_sort_pairs_is_a_stable_sort(backend, kind, m):
    rng = np.random.default_rng(m * 7 + len(kind))
    k = _keys(kind, m, rng)
    v = rng.integers(0, 1 << 31, m, dtype=np.int32)
    ks, vs = backend.sort_pairs(torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
    order = np.argsort(k.view(np.uint64), kind="stable")
    assert np.array_equal(ks.cpu().numpy(), k[order])
    assert np.array_equal(vs.cpu().numpy(), v[order])


@pytest.mark.parametrize("m", [5, 4096, 300_001, 3_000_003])
def test_scatter_is_the_inverse_permutation(backend, m):
    """out[idx[j] - offset] = vals[j]"""
    rng = np.random.default_rng(m)
    perm = rng.permutation(m).astype(np.int32)
    vals = rng.integers(0, 1 << 31, m, dtype=np.int32)
    off = 12345
    out = torch.zeros(m, dtype=torch.int32, device="cuda")
    backend.scatter(torch.from_numpy(perm + off).cuda(), torch.from_numpy(vals).cuda(), off, out)
    want = np.empty(m, dtype=np.int32)
    want[perm] = vals
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("m,shift", [(1000, 2), (70001, 9), (1_500_003, 13)])
def test_bucket_by_index_is_a_stable_partition(backend, m, shift):
    rng = np.random.default_rng(m + shift)
    idx = rng.integers(0, 256 << shift, m, dtype=np.int32)
    vals = rng.integers(0, 1 << 31, m, dtype=np.int32)
    oi, ov, counts = backend.bucket_by_index(torch.from_numpy(idx).cuda(), torch.from_numpy(vals).cuda(), shift)
    order = np.argsort(idx >> shift, kind="stable")
    assert np.array_equal(oi.cpu().numpy(), idx[order])
    assert np.array_equal(ov.cpu().numpy(), vals[order])
    assert np.array_equal(counts, np.bincount(idx >> shift, minlength=256))


def test_stable_perm_by_byte(backend):
    rng = np.random.default_rng(5)
    for m in (1, 257, 100_003):
        d = rng.integers(0, 7, m, dtype=np.uint8)
        perm, cum = backend.stable_perm_by_byte(torch.from_numpy(d).cuda())
        assert np.array_equal(perm.cpu().numpy(), np.argsort(d, kind="stable").astype(np.int32))
        assert np.array_equal(cum[1:], np.cumsum(np.bincount(d, minlength=256)))
