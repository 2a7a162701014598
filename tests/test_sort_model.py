"""CPU model of the quantised per-tile sort (csrc/binning.cu: sort_tiles_q_kernel): sorting 32-bit composites
(depth_bits - min) >> shift | position, then ranking the runs of equal quantised depth with the exact 64-bit keys, must
give exactly the reference order -- ascending (depth bits, Gaussian index) -- for any depth distribution, including
massive ties and tiles whose depth range needs the shift."""
import numpy as np


def quantised_sort(keys: np.ndarray, cap: int) -> np.ndarray:
    """keys: uint64 (depth_bits << 32 | idx) in emission order, len <= cap (a power of two).  Returns the sorted keys."""
    n = len(keys)
    idxb = int(np.log2(cap))
    d = (keys >> np.uint64(32)).astype(np.uint64)
    dmin, dmax = int(d.min()), int(d.max())
    range_bits = (dmax - dmin).bit_length()
    shift = max(0, range_bits - (32 - idxb))
    comp = (((d - np.uint64(dmin)) >> np.uint64(shift)) << np.uint64(idxb)) | np.arange(n, dtype=np.uint64)
    assert int(comp.max()) < 2 ** 32
    order = np.sort(comp.astype(np.uint32))          # the register network: any correct sort of the composites
    qd = order >> np.uint32(idxb)
    src = (order & np.uint32(cap - 1)).astype(np.int64)
    out = np.empty(n, np.uint64)
    p = 0
    while p < n:                                      # runs of equal quantised depth: rank by the exact keys
        e = p + 1
        while e < n and qd[e] == qd[p]:
            e += 1
        run = keys[src[p:e]]
        for k in run:
            out[p + int((run < k).sum())] = k
        p = e
    return out


def _keys(depths: np.ndarray, rng) -> np.ndarray:
    idx = rng.permutation(5_000_000)[: len(depths)].astype(np.uint64)  # unique Gaussian indices, emission order is random
    bits = depths.astype(np.float32).view(np.uint32).astype(np.uint64)
    return (bits << np.uint64(32)) | idx


def test_quantised_sort_equals_the_reference_order():
    rng = np.random.default_rng(3)
    cases = []
    for n, cap in ((1, 256), (7, 256), (256, 256), (300, 512), (1763, 2048), (4096, 4096), (5000, 8192)):
        cases.append((rng.uniform(0.2, 50.0, n), cap))                      # wide range: the shift is active
        cases.append((4.0 + rng.uniform(0, 1e-4, n), cap))                  # narrow range: no shift, few collisions
        cases.append((rng.choice(np.float32([2.0, 3.5, 3.5000002, 6.0]), n), cap))  # massive ties, adjacent floats
        cases.append((np.full(n, 1.25), cap))                               # all equal: one run, order by index
    for depths, cap in cases:
        keys = _keys(np.asarray(depths), rng)
        got = quantised_sort(keys, cap)
        assert np.array_equal(got, np.sort(keys)), (len(keys), cap)
