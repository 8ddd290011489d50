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
    src.append("template <class LoadGbar>")
    src.append("__device__ __forceinline__ void gm_contract_backward(const float (&m)[GM_N_PACKED], float (&a)[GM_N_PACKED], LoadGbar gbar) {")
    src.extend(emit_backward(blocks))
    src.append("}")
    src.append("")
    with open(out_path, "w") as f:
        f.write("\n".join(src) + "\n")
    return blocks, n_feat


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "contract_gen.cuh"))
