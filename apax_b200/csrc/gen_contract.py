#!/usr/bin/env python3
"""Generate `contract_gen.cuh`: straight-line CUDA device code for the eight Gaussian-moment
contractions (apax/layers/descriptor/gaussian_moment_descriptor.py:56-101) on PACKED symmetric
moments, and the exact reverse-mode of the same program.

Packed layout per atom (n_radial = 5 channels x 20 monomials, slot = r*20 + k):
  k = 0          : 1                                  (rank 0)
  k = 1..3       : x y z                              (rank 1)
  k = 4..9       : xx xy xz yy yz zz                  (rank 2, i<=j)
  k = 10..19     : xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz  (rank 3, i<=j<=k)

Every contraction is written as  feature = sum_p  I[p] * M[t, slot(p)]  with I bilinear in M,
grouped into independent blocks (one (r,s) prefix per block).  The forward emitter evaluates a
block's intermediates and features; the backward emitter recomputes the intermediates of the
block and then walks the block's statements in reverse, accumulating into the packed adjoint
A[100] = dE/dM_packed.  Both come from the same statement list, so they cannot disagree.

Feature order = reference concat order c0|c1|c2|c3|c4|c5|c6|c7 with the reference's own
triangular index sets (apax/layers/descriptor/triangular_indices.py:28-49).
"""
import itertools
import os
import sys

NR = 5
NM = 20


def pack2(i, j):
    i, j = sorted((i, j))
    return {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}[(i, j)]


def pack3(i, j, k):
    i, j, k = sorted((i, j, k))
    order = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
    return order.index((i, j, k))


def M0(r):
    return ("m", r * NM)


def M1(r, i):
    return ("m", r * NM + 1 + i)


def M2(r, i, j):
    return ("m", r * NM + 4 + pack2(i, j))


def M2p(r, p):
    return ("m", r * NM + 4 + p)


def M3(r, i, j, k):
    return ("m", r * NM + 10 + pack3(i, j, k))


def tril_2d(n):
    out = []
    for i in range(n):
        out.append((i, i))
        for j in range(i + 1, n):
            out.append((i, j))
    return out


def tril_3d(n):
    out = []
    for i in range(n):
        out.append((i, i, i))
        for j in range(n):
            if j != i:
                out.append((i, j, j))
        for j in range(i + 1, n):
            for k in range(j + 1, n):
                out.append((i, j, k))
    return out


class Block:
    """statements: (target, [(coef, a, b)]) with target ("t", name) or ("f", feature_index);
    a, b operands ("m", slot) or ("t", name)."""

    def __init__(self):
        self.stmts = []

    def define(self, target, terms):
        merged = {}
        for coef, a, b in terms:
            key = tuple(sorted((a, b)))
            merged[key] = merged.get(key, 0) + coef
        self.stmts.append((target, [(c, k[0], k[1]) for k, c in merged.items() if c != 0]))


def build_program():
    t2 = tril_2d(NR)
    t3 = tril_3d(NR)
    off = {}
    pos = 0
    for name, n in (("c0", NR), ("c1", len(t2)), ("c2", len(t2)), ("c3", len(t2)), ("c4", len(t3)),
                    ("c5", len(t2) * NR), ("c6", len(t2) * NR), ("c7", NR ** 3)):
        off[name] = pos
        pos += n
    n_feat = pos
    blocks = []
    ijs = list(itertools.product(range(3), repeat=2))
    ijks = list(itertools.product(range(3), repeat=3))

    # c0: features are the rank-0 moments themselves (coef * m * 1)
    b = Block()
    for r in range(NR):
        b.define(("f", off["c0"] + r), [(1, M0(r), ("one", 0))])
    blocks.append(b)

    # c1, c2, c3: plain dot products over the 15 (r,s) pairs
    for q, (r, s) in enumerate(t2):
        b = Block()
        b.define(("f", off["c1"] + q), [(1, M1(r, i), M1(s, i)) for i in range(3)])
        b.define(("f", off["c2"] + q), [(1, M2(r, i, j), M2(s, i, j)) for i, j in ijs])
        b.define(("f", off["c3"] + q), [(1, M3(r, i, j, k), M3(s, i, j, k)) for i, j, k in ijks])
        blocks.append(b)

    # c4[r,s,t] = sum_ijk M2[r,ij] M2[s,ik] M2[t,jk]; X_rs[(jk)] = sum_i M2[r,ij] M2[s,ik] (+ j<->k)
    by_prefix = {}
    for q, (r, s, t) in enumerate(t3):
        by_prefix.setdefault((r, s), []).append((q, t))
    for (r, s), uses in by_prefix.items():
        b = Block()
        for p in range(6):
            terms = []
            for j, k in ijs:
                if pack2(j, k) != p:
                    continue
                for i in range(3):
                    terms.append((1, M2(r, i, j), M2(s, i, k)))
            b.define(("t", f"X{p}"), terms)
        for q, t in uses:
            b.define(("f", off["c4"] + q), [(1, ("t", f"X{p}"), M2p(t, p)) for p in range(6)])
        blocks.append(b)

    # c5[(rs),t] = sum_ij M1[r,i] M1[s,j] M2[t,ij]
    for q, (r, s) in enumerate(t2):
        b = Block()
        for p in range(6):
            b.define(("t", f"V{p}"), [(1, M1(r, i), M1(s, j)) for i, j in ijs if pack2(i, j) == p])
        for t in range(NR):
            b.define(("f", off["c5"] + q * NR + t), [(1, ("t", f"V{p}"), M2p(t, p)) for p in range(6)])
        blocks.append(b)

    # c6[(rs),t] = sum_ijkl M3[r,ijk] M3[s,ijl] M2[t,kl]
    for q, (r, s) in enumerate(t2):
        b = Block()
        for p in range(6):
            terms = []
            for k, l in ijs:
                if pack2(k, l) != p:
                    continue
                for i, j in ijs:
                    terms.append((1, M3(r, i, j, k), M3(s, i, j, l)))
            b.define(("t", f"U{p}"), terms)
        for t in range(NR):
            b.define(("f", off["c6"] + q * NR + t), [(1, ("t", f"U{p}"), M2p(t, p)) for p in range(6)])
        blocks.append(b)

    # c7[r,s,t] = sum_ijk M3[r,ijk] M2[s,ij] M1[t,k]
    for r in range(NR):
        for s in range(NR):
            b = Block()
            for k in range(3):
                b.define(("t", f"T{k}"), [(1, M3(r, i, j, k), M2(s, i, j)) for i, j in ijs])
            for t in range(NR):
                b.define(("f", off["c7"] + (r * NR + s) * NR + t), [(1, ("t", f"T{k}"), M1(t, k)) for k in range(3)])
            blocks.append(b)
    return blocks, n_feat


def opnd(x):
    kind, v = x
    if kind == "m":
        return f"m[{v}]"
    if kind == "t":
        return v
    if kind == "one":
        return "1.0f"
    raise ValueError(x)


def cf(c):
    return f"{float(c):.1f}f"


def emit_sum(terms):
    """Right-nested fmaf chain; coefficient-grouped so equal-coefficient terms share one multiply."""
    groups = {}
    for c, a, b in terms:
        groups.setdefault(c, []).append((a, b))
    parts = []
    for c, ts in groups.items():
        expr = None
        for a, b in ts:
            if expr is None:
                expr = f"{opnd(a)}*{opnd(b)}"
            else:
                expr = f"fmaf({opnd(a)},{opnd(b)},{expr})"
        parts.append((c, expr))
    out = None
    for c, expr in parts:
        if out is None:
            out = expr if c == 1 else f"{cf(c)}*({expr})"
        else:
            out = f"({out})+{expr}" if c == 1 else f"fmaf({cf(c)},{expr},{out})"
    return out


def emit_forward(blocks):
    lines = []
    for bi, b in enumerate(blocks):
        if bi % 4 == 0:
            lines.append("  GM_CONTRACT_SYNC();")
        lines.append("  {")
        for target, terms in b.stmts:
            if target[0] == "t":
                lines.append(f"    const float {target[1]} = {emit_sum(terms)};")
            else:
                lines.append(f"    store({target[1]}, {emit_sum(terms)});")
        lines.append("  }")
    return lines


def split_blocks(blocks, n_parts):
    """Greedy balance of the (independent) blocks over n_parts by product-term count -> list of block lists."""
    cost = lambda b: sum(len(terms) for _, terms in b.stmts)
    parts, load = [[] for _ in range(n_parts)], [0] * n_parts
    for b in sorted(blocks, key=cost, reverse=True):
        k = load.index(min(load))
        parts[k].append(b)
        load[k] += cost(b)
    return parts, load


def emit_forward_parts(blocks, n_parts):
    """The forward program cut into n_parts independent pieces (whole blocks each) behind a warp-uniform switch: small
    systems run one atom on n_parts warps.  No barriers inside: the warps of a CTA execute different pieces."""
    parts, load = split_blocks(blocks, n_parts)
    lines = [f"  // product terms per part: {load}", "  switch (part) {"]
    for k, bl in enumerate(parts):
        lines.append(f"  case {k}: {{")
        for b in bl:
            lines.append("  {")
            for target, terms in b.stmts:
                if target[0] == "t":
                    lines.append(f"    const float {target[1]} = {emit_sum(terms)};")
                else:
                    lines.append(f"    store({target[1]}, {emit_sum(terms)});")
            lines.append("  }")
        lines.append("  } break;")
    lines.append("  default: break;")
    lines.append("  }")
    return lines


def emit_backward(blocks, prefetch=2):
    """Reverse mode, block by block.  The feature adjoints gbar(f) of a block are loaded `prefetch`
    blocks ahead of their use (function-scope variables g<f>), so the global-load latency overlaps the
    FMA work of the preceding blocks instead of stalling at the point of use."""
    lines = []

    def loads(b):
        return [f"  const float g{t[1]} = gbar({t[1]});" for t, _ in b.stmts if t[0] == "f"]

    for b in blocks[:prefetch]:
        lines.extend(loads(b))
    for bi, b in enumerate(blocks):
        if bi + prefetch < len(blocks):
            lines.extend(loads(blocks[bi + prefetch]))
        if bi % 2 == 0:
            lines.append("  GM_CONTRACT_SYNC();")
        lines.append("  {")
        temps = [t for t, _ in b.stmts if t[0] == "t"]
        for target, terms in b.stmts:
            if target[0] == "t":
                lines.append(f"    const float {target[1]} = {emit_sum(terms)};")
        for t in temps:
            lines.append(f"    float {t[1]}_b = 0.0f;")
        for target, terms in reversed(b.stmts):
            lines.append("    {")
            bar = f"g{target[1]}" if target[0] == "f" else f"{target[1]}_b"
            for c, a, bb in terms:
                scale = bar if c == 1 else f"({cf(c)}*{bar})"
                for x, y in ((a, bb), (bb, a)):
                    if x[0] == "one":
                        continue
                    other = opnd(y)
                    dst = f"a[{x[1]}]" if x[0] == "m" else f"{x[1]}_b"
                    if a == bb:
                        # d(c*x*x) = 2*c*x: emit once with doubled coefficient
                        sc2 = f"({cf(2 * c)}*{bar})"
                        lines.append(f"      {dst} = fmaf({sc2},{other},{dst});")
                        break
                    lines.append(f"      {dst} = fmaf({scale},{other},{dst});")
            lines.append("    }")
        lines.append("  }")
    return lines


def _val(x, m, env):
    kind, v = x
    if kind == "m":
        return m[v]
    if kind == "t":
        return env[v]
    return 1.0


def run_forward(blocks, m, n_feat=360):
    """NumPy interpreter of the generated program: m [100, N] -> features [n_feat, N]."""
    import numpy as np

    out = np.zeros((n_feat,) + m.shape[1:], dtype=m.dtype)
    for b in blocks:
        env = {}
        for target, terms in b.stmts:
            acc = 0
            for c, x, y in terms:
                acc = acc + c * _val(x, m, env) * _val(y, m, env)
            if target[0] == "t":
                env[target[1]] = acc
            else:
                out[target[1]] = acc
    return out


def run_backward(blocks, m, gbar):
    """NumPy interpreter of the reverse program: (m [100,N], gbar [n_feat,N]) -> a [100,N]."""
    import numpy as np

    a = np.zeros_like(m)
    for b in blocks:
        env, bars = {}, {}
        for target, terms in b.stmts:
            if target[0] == "t":
                env[target[1]] = sum(c * _val(x, m, env) * _val(y, m, env) for c, x, y in terms)
                bars[target[1]] = 0
        for target, terms in reversed(b.stmts):
            bar = gbar[target[1]] if target[0] == "f" else bars[target[1]]
            for c, x, y in terms:
                for u, w in ((x, y), (y, x)):
                    if u[0] == "m":
                        a[u[1]] = a[u[1]] + c * bar * _val(w, m, env)
                    elif u[0] == "t":
                        bars[u[1]] = bars[u[1]] + c * bar * _val(w, m, env)
    return a


# ------------------------------------------------------------------------------------------------
# Looped reverse mode.  The straight-line reverse program is ~200 KB of code and runs at 12 % issue
# utilisation (84 % of the stall samples are instruction-fetch stalls).  All blocks of one contraction
# type have the same shape and differ only in WHICH radial channels (r, s) they read, so the looped
# form keeps one generic block per type and loops over the (r, s) instances: the two channels are
# copied into fixed registers through a warp-uniform switch, the generic code runs on them, and their
# adjoint increments are added back through the same switch.  Operands indexed by the third channel t
# stay statically indexed (all five t are unrolled).  ~15 KB of code: it stays in the instruction cache.
# ------------------------------------------------------------------------------------------------
def _sym_M(kind):
    """moment accessors for a symbolic channel ('R' -> mr[], 'S' -> ms[]) or a numeric one"""
    def slot0(r):
        return (("mr" if r == "R" else "ms"), 0) if isinstance(r, str) else ("m", r * NM)
    def m1(r, i):
        return (("mr" if r == "R" else "ms"), 1 + i) if isinstance(r, str) else ("m", r * NM + 1 + i)
    def m2(r, i, j):
        return (("mr" if r == "R" else "ms"), 4 + pack2(i, j)) if isinstance(r, str) else ("m", r * NM + 4 + pack2(i, j))
    def m2p(r, p):
        return (("mr" if r == "R" else "ms"), 4 + p) if isinstance(r, str) else ("m", r * NM + 4 + p)
    def m3(r, i, j, k):
        return (("mr" if r == "R" else "ms"), 10 + pack3(i, j, k)) if isinstance(r, str) else ("m", r * NM + 10 + pack3(i, j, k))
    return slot0, m1, m2, m2p, m3


def build_looped():
    """-> (types, n_feat); types = list of dict(name, block (generic, local feature targets), krange_r, krange_s,
    instances = [(r, s, [absolute feature index or -1 per local feature])])."""
    _, m1, m2, m2p, m3 = _sym_M(None)
    t2 = tril_2d(NR)
    t3 = tril_3d(NR)
    off = {}
    pos = 0
    for name, n in (("c0", NR), ("c1", len(t2)), ("c2", len(t2)), ("c3", len(t2)), ("c4", len(t3)),
                    ("c5", len(t2) * NR), ("c6", len(t2) * NR), ("c7", NR ** 3)):
        off[name] = pos
        pos += n
    ijs = list(itertools.product(range(3), repeat=2))
    ijks = list(itertools.product(range(3), repeat=3))
    types = []

    b = Block()   # c1, c2, c3
    b.define(("f", 0), [(1, m1("R", i), m1("S", i)) for i in range(3)])
    b.define(("f", 1), [(1, m2("R", i, j), m2("S", i, j)) for i, j in ijs])
    b.define(("f", 2), [(1, m3("R", i, j, k), m3("S", i, j, k)) for i, j, k in ijks])
    types.append(dict(name="dot", block=b, kr=(1, 20), ks=(1, 20), nf=3,
                      instances=[(r, s, [off["c1"] + q, off["c2"] + q, off["c3"] + q]) for q, (r, s) in enumerate(t2)]))

    b = Block()   # c4
    for p in range(6):
        terms = []
        for j, k in ijs:
            if pack2(j, k) != p:
                continue
            for i in range(3):
                terms.append((1, m2("R", i, j), m2("S", i, k)))
        b.define(("t", f"X{p}"), terms)
    for t in range(NR):
        b.define(("f", t), [(1, ("t", f"X{p}"), m2p(t, p)) for p in range(6)])
    by_prefix = {}
    for q, (r, s, t) in enumerate(t3):
        by_prefix.setdefault((r, s), [-1] * NR)[t] = off["c4"] + q
    types.append(dict(name="c4", block=b, kr=(4, 10), ks=(4, 10), nf=NR,
                      instances=[(r, s, feats) for (r, s), feats in by_prefix.items()]))

    b = Block()   # c5
    for p in range(6):
        b.define(("t", f"V{p}"), [(1, m1("R", i), m1("S", j)) for i, j in ijs if pack2(i, j) == p])
    for t in range(NR):
        b.define(("f", t), [(1, ("t", f"V{p}"), m2p(t, p)) for p in range(6)])
    types.append(dict(name="c5", block=b, kr=(1, 4), ks=(1, 4), nf=NR,
                      instances=[(r, s, [off["c5"] + q * NR + t for t in range(NR)]) for q, (r, s) in enumerate(t2)]))

    b = Block()   # c6
    for p in range(6):
        terms = []
        for k, l in ijs:
            if pack2(k, l) != p:
                continue
            for i, j in ijs:
                terms.append((1, m3("R", i, j, k), m3("S", i, j, l)))
        b.define(("t", f"U{p}"), terms)
    for t in range(NR):
        b.define(("f", t), [(1, ("t", f"U{p}"), m2p(t, p)) for p in range(6)])
    types.append(dict(name="c6", block=b, kr=(10, 20), ks=(10, 20), nf=NR,
                      instances=[(r, s, [off["c6"] + q * NR + t for t in range(NR)]) for q, (r, s) in enumerate(t2)]))

    b = Block()   # c7
    for k in range(3):
        b.define(("t", f"T{k}"), [(1, m3("R", i, j, k), m2("S", i, j)) for i, j in ijs])
    for t in range(NR):
        b.define(("f", t), [(1, ("t", f"T{k}"), m1(t, k)) for k in range(3)])
    types.append(dict(name="c7", block=b, kr=(10, 20), ks=(4, 10), nf=NR,
                      instances=[(r, s, [off["c7"] + (r * NR + s) * NR + t for t in range(NR)])
                                 for r in range(NR) for s in range(NR)]))
    return types, off, pos


def expand_looped(types, off):
    """Concrete blocks (absolute slots / features) of the looped program, for the NumPy interpreters."""
    blocks = []
    b = Block()
    for r in range(NR):
        b.define(("f", off["c0"] + r), [(1, M0(r), ("one", 0))])
    blocks.append(b)

    def conc(x, r, s):
        if x[0] == "mr":
            return ("m", r * NM + x[1])
        if x[0] == "ms":
            return ("m", s * NM + x[1])
        return x

    for ty in types:
        for r, s, feats in ty["instances"]:
            nb = Block()
            for target, terms in ty["block"].stmts:
                if target[0] == "f":
                    if feats[target[1]] < 0:
                        continue
                    target = ("f", feats[target[1]])
                nb.stmts.append((target, [(c, conc(x, r, s), conc(y, r, s)) for c, x, y in terms]))
            blocks.append(nb)
    return blocks


def _lop(x):
    kind, v = x
    return {"m": f"m[{v}]", "mr": f"mr[{v}]", "ms": f"ms[{v}]", "t": f"{v}", "one": "1.0f"}[kind] if kind != "t" else v


def _ldst(x):
    kind, v = x
    return {"m": f"a[{v}]", "mr": f"ar[{v}]", "ms": f"as_[{v}]"}[kind] if kind != "t" else f"{v}_b"


def _lsum(terms):
    groups = {}
    for c, a, b in terms:
        groups.setdefault(c, []).append((a, b))
    out = None
    for c, ts in groups.items():
        expr = None
        for a, b in ts:
            expr = f"{_lop(a)}*{_lop(b)}" if expr is None else f"fmaf({_lop(a)},{_lop(b)},{expr})"
        if out is None:
            out = expr if c == 1 else f"{cf(c)}*({expr})"
        else:
            out = f"({out})+{expr}" if c == 1 else f"fmaf({cf(c)},{expr},{out})"
    return out


GM_RING_DEPTH = 6   # instances whose feature adjoints are in flight (cp.async into a shared-memory ring)
GM_RING_NF = 5      # feature adjoints per instance, at most


def emit_backward_looped(types, off):
    """The feature adjoints of an instance are only ~100-250 instructions of work away from the next instance's, far
    less than a DRAM round trip, and registers are exhausted (255), so they travel through a per-thread ring in
    shared memory filled by cp.async GM_RING_DEPTH - 1 instances ahead."""
    D, NFX = GM_RING_DEPTH, GM_RING_NF
    L = []
    L.append("  // c0: the rank-0 moments are features themselves")
    L.append("  if (part == 0) {")
    for r in range(NR):
        L.append(f"    a[{r * NM}] += gt[int64_t({off['c0'] + r}) * ld];")
    L.append("  }")
    for ty in types:
        n, nf, name = len(ty["instances"]), ty["nf"], ty["name"]
        (r0, r1), (s0, s1) = ty["kr"], ty["ks"]
        L.append(f"  // ---- {name}: {n} (r, s) instances of one generic block; this part takes q = part, part + NP, ...")
        L.append("  // (one fetch pipeline running across the type boundaries was measured: no gain, the fill latency is not what binds)")
        L.append("  {")
        L.append("#pragma unroll 1")
        L.append(f"    for (int k = 0; k < {D - 1}; ++k) {{")
        L.append("      const int q = part + k * NP;")
        L.append(f"      if (q < {n}) {{")
        for j in range(nf):
            L.append(f"        if (gm_tab_{name}_f[q][{j}] >= 0) gm_cp_async4(ring + (k * {NFX} + {j}) * ring_stride, gt + int64_t(gm_tab_{name}_f[q][{j}]) * ld);")
        L.append("      }")
        L.append("      gm_cp_async_commit();")
        L.append("    }")
        L.append("#pragma unroll 1")
        L.append(f"    for (int k = 0, q = part; q < {n}; ++k, q += NP) {{")
        L.append(f"      const int qn = q + {D - 1} * NP, slot = k % {D}, slot_n = (k + {D - 1}) % {D};")
        L.append(f"      if (qn < {n}) {{")
        for j in range(nf):
            L.append(f"        if (gm_tab_{name}_f[qn][{j}] >= 0) gm_cp_async4(ring + (slot_n * {NFX} + {j}) * ring_stride, gt + int64_t(gm_tab_{name}_f[qn][{j}]) * ld);")
        L.append("      }")
        L.append("      gm_cp_async_commit();")
        L.append(f"      gm_cp_async_wait<{D - 1}>();   // everything but the {D - 1} youngest groups has landed: instance q is complete")
        L.append(f"      float gc[{nf}];")
        for j in range(nf):
            L.append(f"      gc[{j}] = gm_tab_{name}_f[q][{j}] >= 0 ? ring[(slot * {NFX} + {j}) * ring_stride] : 0.0f;")
        L.append(f"      const int r = gm_tab_{name}_rs[q][0], s = gm_tab_{name}_rs[q][1];")
        L.append("      float mr[20], ms[20], ar[20], as_[20];")
        L.append(f"      gm_load_channel<{r0}, {r1}>(m, r, mr);")
        L.append(f"      gm_load_channel<{s0}, {s1}>(m, s, ms);")
        L.append("#pragma unroll")
        L.append(f"      for (int k2 = {r0}; k2 < {r1}; ++k2) ar[k2] = 0.0f;")
        L.append("#pragma unroll")
        L.append(f"      for (int k2 = {s0}; k2 < {s1}; ++k2) as_[k2] = 0.0f;")
        b = ty["block"]
        temps = [t for t, _ in b.stmts if t[0] == "t"]
        for target, terms in b.stmts:
            if target[0] == "t":
                L.append(f"      const float {target[1]} = {_lsum(terms)};")
        for t in temps:
            L.append(f"      float {t[1]}_b = 0.0f;")
        for target, terms in reversed(b.stmts):
            bar = f"gc[{target[1]}]" if target[0] == "f" else f"{target[1]}_b"
            for c, x, y in terms:
                scale = bar if c == 1 else f"({cf(c)}*{bar})"
                for u, w in ((x, y), (y, x)):
                    if u[0] == "one":
                        continue
                    L.append(f"      {_ldst(u)} = fmaf({scale},{_lop(w)},{_ldst(u)});")
        L.append(f"      gm_add_channel<{r0}, {r1}>(a, r, ar);")
        L.append(f"      gm_add_channel<{s0}, {s1}>(a, s, as_);")
        L.append("    }")
        L.append("    gm_cp_async_wait<0>();")
        L.append("  }")
    return L


def emit_looped_tables(types):
    L = []
    for ty in types:
        name, n, nf = ty["name"], len(ty["instances"]), ty["nf"]
        rs = ", ".join("{%d, %d}" % (r, s) for r, s, _ in ty["instances"])
        fs = ", ".join("{" + ", ".join(str(f) for f in feats) + "}" for _, _, feats in ty["instances"])
        L.append(f"__constant__ int gm_tab_{name}_rs[{n}][2] = {{{rs}}};")
        L.append(f"__constant__ int gm_tab_{name}_f[{n}][{nf}] = {{{fs}}};")
    return L


LOOPED_PRELUDE = """__device__ __forceinline__ void gm_cp_async4(float* smem, const float* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void gm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gm_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// radial channel r of the packed moments -> fixed registers (warp-uniform switch: r is a loop counter)
template <int K0, int K1>
__device__ __forceinline__ void gm_load_channel(const float (&m)[GM_N_PACKED], int r, float (&out)[20]) {
  switch (r) {
#define GM_CH(R) case R: _Pragma("unroll") for (int k = K0; k < K1; ++k) out[k] = m[R * 20 + k]; break;
    GM_CH(0) GM_CH(1) GM_CH(2) GM_CH(3) default: _Pragma("unroll") for (int k = K0; k < K1; ++k) out[k] = m[80 + k]; break;
#undef GM_CH
  }
}
template <int K0, int K1>
__device__ __forceinline__ void gm_add_channel(float (&a)[GM_N_PACKED], int r, const float (&inc)[20]) {
  switch (r) {
#define GM_CH(R) case R: _Pragma("unroll") for (int k = K0; k < K1; ++k) a[R * 20 + k] += inc[k]; break;
    GM_CH(0) GM_CH(1) GM_CH(2) GM_CH(3) default: _Pragma("unroll") for (int k = K0; k < K1; ++k) a[80 + k] += inc[k]; break;
#undef GM_CH
  }
}"""


def count_ops(blocks):
    fwd = sum(len(terms) for b in blocks for _, terms in b.stmts)
    return fwd


def main(out_path):
    blocks, n_feat = build_program()
    assert n_feat == 360, n_feat
    src = []
    src.append("// GENERATED by gen_contract.py -- do not edit.  Packed Gaussian-moment contractions")
    src.append("// (apax/layers/descriptor/gaussian_moment_descriptor.py:56-101) and their reverse mode.")
    src.append(f"// {len(blocks)} blocks, {count_ops(blocks)} forward product terms, {n_feat} features.")
    src.append("#pragma once")
    src.append(f"#define GM_N_FEATURES {n_feat}")
    src.append("#define GM_N_PACKED 100")
    src.append("// m = packed moments of one atom, a = packed adjoint accumulators (caller zero-initialises).")
    src.append("// The program is ~100-200 KB of straight-line code, far beyond the instruction cache: GM_CONTRACT_SYNC()")
    src.append("// (a block barrier, defined by the including kernel) keeps the warps of a CTA walking it in step so")
    src.append("// that one instruction fetch serves all of them.")
    src.append("#ifndef GM_CONTRACT_SYNC")
    src.append("#define GM_CONTRACT_SYNC() __syncthreads()")
    src.append("#endif")
    src.append("template <class Store>")
    src.append("__device__ __forceinline__ void gm_contract_forward(const float (&m)[GM_N_PACKED], Store store) {")
    src.extend(emit_forward(blocks))
    src.append("}")
    src.append("")
    src.append("#define GM_FWD_PARTS 4")
    src.append("// the same program in GM_FWD_PARTS independent pieces (whole blocks; `part` is warp-uniform): no barriers inside")
    src.append("template <class Store>")
    src.append("__device__ __forceinline__ void gm_contract_forward_part(const float (&m)[GM_N_PACKED], int part, Store store) {")
    src.extend(emit_forward_parts(blocks, 4))
    src.append("}")
    src.append("")
    src.append("template <class LoadGbar>")
    src.append("__device__ __forceinline__ void gm_contract_backward(const float (&m)[GM_N_PACKED], float (&a)[GM_N_PACKED], LoadGbar gbar) {")
    src.extend(emit_backward(blocks))
    src.append("}")
    src.append("")
    types, off, nf2 = build_looped()
    assert nf2 == n_feat
    src.append("// ---- looped reverse mode: one generic block per contraction type, instances from constant tables")
    src.extend(emit_looped_tables(types))
    src.append(LOOPED_PRELUDE)
    src.append(f"#define GM_RING_FLOATS_PER_THREAD {GM_RING_DEPTH * GM_RING_NF}")
    src.append("// gt: this atom's column of the feature-major adjoints (row stride ld); ring: this thread's slot 0 of a shared-memory")
    src.append("// ring of GM_RING_FLOATS_PER_THREAD floats per thread, consecutive slots ring_stride floats apart")
    src.append("// NP > 1: the instances of every type are dealt round-robin to NP parts (one warp each, `part` warp-uniform); the")
    src.append("// caller sums the NP partial adjoints.  NP == 1, part == 0 is the whole program.")
    src.append("template <int NP>")
    src.append("__device__ __forceinline__ void gm_contract_backward_looped(const float (&m)[GM_N_PACKED], float (&a)[GM_N_PACKED],")
    src.append("                                                            const float* gt, int64_t ld, float* ring, int ring_stride,")
    src.append("                                                            int part = 0) {")
    src.extend(emit_backward_looped(types, off))
    src.append("}")
    src.append("")
    with open(out_path, "w") as f:
        f.write("\n".join(src) + "\n")
    return blocks, n_feat


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "contract_gen.cuh"))
