"""CPU twins of two index algorithms inside the CUDA kernels, transcribed statement by statement, so that their logic is
exercised without a GPU (the kernels themselves are covered by tests/test_gpu_neighbors.py and the parity tests):

* `warp_bitonic_sort<NQ>` (apax_b200/csrc/nl.cu): element e = q * 32 + lane; exchanges at distance >= 32 are register
  swaps inside a lane, the others shuffles between lanes; padding keys INT_MAX sink to the end.
* the interpolated start of the transpose-map search (k_transpose_symmetric, apax_b200/csrc/gm.cu): one probe at the
  interpolated position, one at the far end of an 8-slot window, then a binary search on what is left -- for uniform AND
  for clustered neighbour indices it must find exactly what the plain binary search finds."""
import numpy as np
import pytest


def warp_bitonic_sort(k):
    """k: int array [NQ][32] (register q of lane l).  Returns the array after the network."""
    k = k.copy()
    nq = k.shape[0]
    lane = np.arange(32)
    size = 2
    while size <= nq * 32:
        stride = size >> 1
        while stride > 0:
            if stride >= 32:
                dq = stride >> 5
                for q in range(nq):
                    pq = q ^ dq
                    if q < pq:
                        up = ((q * 32) & size) == 0
                        lo, hi = np.minimum(k[q], k[pq]), np.maximum(k[q], k[pq])
                        k[q], k[pq] = (lo, hi) if up else (hi, lo)
            else:
                lower = (lane & stride) == 0
                for q in range(nq):
                    other = k[q][lane ^ stride]                       # __shfl_xor_sync
                    up = ((q * 32 + lane) & size) == 0
                    k[q] = np.where(lower == up, np.minimum(k[q], other), np.maximum(k[q], other))
            stride >>= 1
        size <<= 1
    return k


@pytest.mark.parametrize("nq", [4, 8])
def test_bitonic_network_sorts_every_row_length(nq):
    rng = np.random.default_rng(nq)
    for cnt in list(range(0, 40)) + [63, 64, 65, 100, 127, 128] + ([129, 200, 255, 256] if nq == 8 else []):
        if cnt > nq * 32:
            continue
        keys = rng.choice(2_000_000, size=cnt, replace=False).astype(np.int64)
        k = np.full((nq, 32), 0x7FFFFFFF, dtype=np.int64)
        for e in range(cnt):
            k[e // 32][e % 32] = keys[e]
        out = warp_bitonic_sort(k)
        flat = np.array([out[e // 32][e % 32] for e in range(nq * 32)])
        assert np.array_equal(flat[:cnt], np.sort(keys)), cnt
        assert np.all(flat[cnt:] == 0x7FFFFFFF)


def transpose_search(nbr, b, e, a, n_atoms):
    """Position of `a` in the sorted slice nbr[b..e] (inclusive) or -1; the kernel's probes in the kernel's order."""
    found = -1
    if e - b >= 16:
        g = b + ((e - b + 1) * a) // n_atoms
        gc = min(g, e)
        v = nbr[gc]
        if v == a:
            found, b = gc, e + 1
        elif v < a:
            b = gc + 1
            w = min(gc + 8, e)
            if b <= e:
                if nbr[w] >= a:
                    e = w
                else:
                    b = w + 1
        else:
            e = gc - 1
            w = max(gc - 8, b)
            if b <= e:
                if nbr[w] <= a:
                    b = w
                else:
                    e = w - 1
    while b <= e:
        mid = (b + e) >> 1
        v = nbr[mid]
        if v == a:
            return mid
        if v < a:
            b = mid + 1
        else:
            e = mid - 1
    return found


@pytest.mark.parametrize("clustered", [False, True])
def test_interpolated_transpose_search_equals_binary_search(clustered):
    rng = np.random.default_rng(5 + clustered)
    n_atoms = 100_000
    for trial in range(300):
        cnt = int(rng.integers(1, 200))
        if clustered:                                  # indices bunched in a narrow range: the interpolation guess is wrong
            lo = int(rng.integers(0, n_atoms - 400))
            row = np.sort(rng.choice(np.arange(lo, lo + 400), size=min(cnt, 400), replace=False))
        else:
            row = np.sort(rng.choice(n_atoms, size=cnt, replace=False))
        pad = int(rng.integers(0, 50))
        nbr = np.concatenate([rng.integers(0, n_atoms, pad), row, rng.integers(0, n_atoms, 7)])
        b, e = pad, pad + len(row) - 1
        for a in list(row[:: max(1, len(row) // 12)]) + [int(rng.integers(0, n_atoms)) for _ in range(6)] + [0, n_atoms - 1]:
            want = np.flatnonzero(row == a)
            want = pad + int(want[0]) if len(want) else -1
            assert transpose_search(nbr, b, e, int(a), n_atoms) == want


def exp_chain(x):
    """apax_b200/csrc/md.cu::exp_chain: Taylor polynomials in Estrin form below 2^-8 (degree 5) and 2^-4 (degree 9)."""
    ax = abs(x)
    if ax > 0.0625:
        return float(np.exp(x))
    x2 = x * x
    a01, a23, a45 = 1.0 + x, x * (1.0 / 6.0) + 0.5, x * (1.0 / 120.0) + 1.0 / 24.0
    x4 = x2 * x2
    lo = x2 * a23 + a01
    if ax < 0.00390625:
        return x4 * a45 + lo
    a67, a89 = x * (1.0 / 5040.0) + 1.0 / 720.0, x * (1.0 / 362880.0) + 1.0 / 40320.0
    hi = x2 * a67 + a45
    hi2 = x4 * a89 + hi
    return x4 * hi2 + lo


def test_chain_exponential_is_exact_to_an_ulp():
    xs = np.concatenate([np.linspace(-0.0625, 0.0625, 4001), np.linspace(-0.0039, 0.0039, 4001), [0.0, 1e-300, -1e-12, 0.3, -0.7]])
    err = max(abs(exp_chain(float(x)) - np.exp(x)) / np.exp(x) for x in xs)
    assert err <= 2.5e-16, err
