"""CPU ORACLE (test infrastructure, NOT product code) for the two neighbour-list semantics
on apax's GMNN path.  Brute force, NumPy fp64, with the floating-point operation order
written out so that the CUDA kernels can reproduce the pair set bit for bit.

(1) `jaxmd_sparse_list`  -- apax/utils/jax_md_reduced/partition.py:478-787 with
    disable_cell_list=True, format=Sparse, fractional_coordinates=True (the call sites
    apax/md/ase_calc.py:54-62 and apax/md/simulate.py:506-514).
(2) `vesin_full_list`    -- the third-party vesin 0.4.2 `NeighborList(cutoff, full_list=True)`
    called at apax/md/ase_calc.py:337-355 and apax/data/preprocessing.py:47-54.  vesin is not
    under /root/reference and not installed; its published contract (all (i,j,S) with
    |r_j - r_i + S.cell| < cutoff, excluding (i,i,0)) is restated by image enumeration.  The
    reference pins it only as a multiset of distances (tests/unit_tests/data/
    test_input_pipeline.py:171-212); pair ORDER is unpinned, so compare as sets.

Parity unpinned against vesin itself (absent); pinned against the reference's partition.py
source through tests/golden/make_golden.py for (1).
"""
from __future__ import annotations

import numpy as np


def wrap_unit(dR: np.ndarray) -> np.ndarray:
    """space.periodic_displacement(side=1) (space.py:146-157): mod(dR + 0.5, 1) - 0.5.

    np.mod == jnp.mod bit for bit: fmod, then +1 when the remainder is negative.
    """
    return np.mod(dR + 0.5, 1.0) - 0.5


def metric_sq_minimage(Ra: np.ndarray, Rb: np.ndarray, box: np.ndarray) -> np.ndarray:
    """square_distance(periodic_general.displacement_fn(Ra, Rb, box=box)) in fp64
    (space.py:260-282, :160-168).  Operation order (the CUDA predicate follows it, without
    FMA contraction): w = wrap(Ra-Rb); d_i = (box[i,0]*w0 + box[i,1]*w1) + box[i,2]*w2;
    d2 = (d0*d0 + d1*d1) + d2*d2.
    """
    w = wrap_unit(Ra - Rb)
    d = []
    for i in range(3):
        d.append((box[i, 0] * w[..., 0] + box[i, 1] * w[..., 1]) + box[i, 2] * w[..., 2])
    return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]


def jaxmd_sparse_pairs(frac: np.ndarray, box: np.ndarray, r_cutoff: float, dr_threshold: float = 0.0,
                       block: int = 1024):
    """Valid entries of the Sparse list, in the reference's order.

    prune_neighbor_list_sparse (partition.py:640-664): candidates are all (sender i, receiver j)
    in row-major order over the N x N matrix minus the diagonal (mask_self, :613-617); an entry
    survives iff metric_sq(R[i], R[j]) < (r_cutoff + dr_threshold)**2 (:574-576, :650-652); the
    output is stack((receiver, sender)) (:664), i.e. idx[0]=j (neighbour), idx[1]=i (central),
    i ascending and j ascending within i.
    """
    frac = np.asarray(frac, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    n = frac.shape[0]
    cutoff = r_cutoff + dr_threshold
    cutoff_sq = cutoff * cutoff  # integer_pow(x, 2) == x*x
    recv, send = [], []
    for s in range(0, n, block):
        e = min(n, s + block)
        d2 = metric_sq_minimage(frac[s:e, None, :], frac[None, :, :], box)
        mask = d2 < cutoff_sq
        mask[np.arange(e - s), np.arange(s, e)] = False
        ii, jj = np.nonzero(mask)  # row-major: i ascending, j ascending within i
        send.append(ii + s)
        recv.append(jj)
    return np.stack([np.concatenate(recv), np.concatenate(send)]).astype(np.int32)


def jaxmd_sparse_pairs_fast(frac: np.ndarray, box: np.ndarray, r_cutoff: float, dr_threshold: float = 0.0):
    """Same output as `jaxmd_sparse_pairs` for ORTHORHOMBIC boxes, without the N^2 candidate matrix:
    candidate pairs come from a periodic k-d tree with a slightly enlarged radius, the keep/drop
    decision is then made by exactly the same fp64 predicate, and the survivors are put in the
    reference's order.  Used to feed the CPU baseline at sizes where N^2 is too slow."""
    from scipy.spatial import cKDTree

    frac = np.asarray(frac, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    assert np.allclose(box, np.diag(np.diag(box))), "orthorhombic boxes only"
    lengths = np.diag(box)
    cutoff = r_cutoff + dr_threshold
    wrapped = (frac - np.floor(frac)) * lengths
    wrapped = np.minimum(wrapped, np.nextafter(lengths, 0.0))
    tree = cKDTree(wrapped, boxsize=lengths)
    cand = tree.query_pairs(cutoff * (1.0 + 1e-9) + 1e-9, output_type="ndarray")
    i = np.concatenate([cand[:, 0], cand[:, 1]])
    j = np.concatenate([cand[:, 1], cand[:, 0]])
    keep = metric_sq_minimage(frac[i], frac[j], box) < cutoff * cutoff
    i, j = i[keep], j[keep]
    order = np.lexsort((j, i))
    return np.stack([j[order], i[order]]).astype(np.int32)


def jaxmd_sparse_list(frac, box, r_cutoff, dr_threshold=0.0, capacity_multiplier=1.25, extra_capacity=0,
                      max_occupancy=None):
    """neighbor_list_fn/allocate (partition.py:666-760): idx [2, capacity] padded with N, and the
    overflow bit.  capacity = int(occupancy*1.25 + N*extra_capacity) clipped to N(N-1) (:714-729)."""
    n = np.asarray(frac).shape[0]
    pairs = jaxmd_sparse_pairs(frac, box, r_cutoff, dr_threshold)
    occupancy = pairs.shape[1]
    if max_occupancy is None:
        max_occupancy = int(occupancy * capacity_multiplier + n * extra_capacity)
        max_occupancy = min(max_occupancy, n * n, n * (n - 1))
    idx = np.full((2, max_occupancy), n, dtype=np.int32)
    k = min(occupancy, max_occupancy)
    idx[:, :k] = pairs[:, :k]
    overflow = occupancy > max_occupancy
    return idx, occupancy, bool(overflow)


def needs_rebuild(frac, reference_frac, box, dr_threshold: float) -> bool:
    """The skin test of neighbor_list_fn (partition.py:771-779):
    any(metric_sq(position, reference_position) > (dr_threshold/2)**2)."""
    thr = (dr_threshold / 2.0) ** 2
    return bool(np.any(metric_sq_minimage(np.asarray(frac), np.asarray(reference_frac), np.asarray(box)) > thr))


def vesin_full_list(positions: np.ndarray, cell: np.ndarray, cutoff: float, periodic: bool = True):
    """All (i, j, S) with |r_j - r_i + S @ cell| < cutoff, excluding (i, i, 0); any cell shape.

    positions Cartesian [N,3]; cell rows are lattice vectors (ASE convention, as passed at
    ase_calc.py:338-344).  Returns idx_i, idx_j (int32), S (int32 [P,3]), sorted by (i, j, S).
    Operation order for the distance (mirrored by the CUDA kernel, no FMA contraction):
      t_k = (S0*cell[0,k] + S1*cell[1,k]) + S2*cell[2,k];  D_k = (r_j[k] - r_i[k]) + t_k;
      d2 = (D0*D0 + D1*D1) + D2*D2;  keep iff d2 < cutoff*cutoff.
    """
    pos = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64)
    n = pos.shape[0]
    cutoff_sq = cutoff * cutoff
    if periodic:
        inv = np.linalg.inv(cell)
        # perpendicular heights of the cell: 1/|column k of inv|
        heights = 1.0 / np.linalg.norm(inv, axis=0)
        frac = pos @ inv
        span = frac.max(axis=0) - frac.min(axis=0)
        nmax = np.ceil(cutoff / heights + span).astype(int) + 1
    else:
        nmax = np.zeros(3, dtype=int)
    out_i, out_j, out_s = [], [], []
    diff = pos[None, :, :] - pos[:, None, :]  # [i, j] = r_j - r_i
    for s0 in range(-nmax[0], nmax[0] + 1):
        for s1 in range(-nmax[1], nmax[1] + 1):
            for s2 in range(-nmax[2], nmax[2] + 1):
                t = [(s0 * cell[0, k] + s1 * cell[1, k]) + s2 * cell[2, k] for k in range(3)]
                D = [diff[:, :, k] + t[k] for k in range(3)]
                d2 = (D[0] * D[0] + D[1] * D[1]) + D[2] * D[2]
                mask = d2 < cutoff_sq
                if s0 == 0 and s1 == 0 and s2 == 0:
                    mask[np.arange(n), np.arange(n)] = False
                ii, jj = np.nonzero(mask)
                if len(ii):
                    out_i.append(ii)
                    out_j.append(jj)
                    out_s.append(np.tile(np.array([[s0, s1, s2]]), (len(ii), 1)))
    if not out_i:
        z = np.zeros((0,), dtype=np.int32)
        return z, z.copy(), np.zeros((0, 3), dtype=np.int32)
    ii = np.concatenate(out_i)
    jj = np.concatenate(out_j)
    ss = np.concatenate(out_s)
    order = np.lexsort((ss[:, 2], ss[:, 1], ss[:, 0], jj, ii))
    return ii[order].astype(np.int32), jj[order].astype(np.int32), ss[order].astype(np.int32)


def vesin_idx_offsets(positions, cell, cutoff, periodic=True, padded_length=None):
    """What ASECalculator.set_neighbours_and_offsets hands to the model (ase_calc.py:337-355):
    idx int32 [2,P] = [idxs_i, idxs_j], offsets f64 [P,3] = S @ cell, zero-padded with (0,0) pairs."""
    ii, jj, ss = vesin_full_list(positions, cell, cutoff, periodic)
    idx = np.stack([ii, jj]).astype(np.int32)
    offsets = np.matmul(ss.astype(np.float64), np.asarray(cell, dtype=np.float64))
    if padded_length is not None:
        pad = padded_length - idx.shape[1]
        assert pad >= 0
        idx = np.pad(idx, ((0, 0), (0, pad)), "constant")
        offsets = np.pad(offsets, ((0, pad), (0, 0)), "constant")
    return idx, offsets


def neighbor_calculable_with_jax(box: np.ndarray, r_max: float) -> bool:
    """apax/md/ase_calc.py:501-522 — min cell height / 2 > r_max selects the min-image path."""
    box = np.asarray(box, dtype=np.float64)
    if np.all(box < 1e-6):
        return True
    a = [box[0], box[0], box[1]]
    b = [box[1], box[2], box[2]]
    c = [box[2], box[1], box[0]]
    heights = []
    for k in range(3):
        nv = np.cross(a[k], b[k])
        proj = c[k] - np.sum(nv * c[k]) / np.sum(nv ** 2) * nv
        heights.append(np.linalg.norm(c[k] - proj))
    return bool(np.min(heights) / 2 > r_max)
