// Neighbour lists on the GPU: a cell list in fractional space feeding the two list semantics apax
// uses, with predicates evaluated in fp64 in a fixed, unfused operation order so that the pair SET
// is bit-identical to the CPU oracle's restatement of the reference:
//   min-image Sparse list : apax/utils/jax_md_reduced/partition.py:478-787 (predicate :640-664,
//                           metric apax/utils/jax_md_reduced/space.py:146-157,260-282)
//   full list with shifts : vesin NeighborList(full_list=True) as called at apax/md/ase_calc.py:337-355
// The reference itself has no cell list (partition.py:693-699 raises NotImplementedError) and
// enumerates all N^2 candidates; this file is the O(N) replacement.
#include <math.h>

#include "nl_internal.cuh"

namespace apax {

size_t carve_nl_workspace(void* base, int64_t n_atoms, int64_t capacity, NlWorkspace* ws) {
  Carver c(base);
  NlWorkspace w{};
  const int64_t N = n_atoms > 0 ? n_atoms : 1, C = capacity > 0 ? capacity : 1;
  w.n_atoms = n_atoms;
  w.capacity = capacity;
  w.grid = c.take<NlGrid>(1);
  w.cell_of = c.take<int32_t>(N);
  w.wrap = c.take<int32_t>(3 * N);
  w.cell_cnt = c.take<int32_t>(N + 1);
  w.cellptr = c.take<int32_t>(N + 1);
  w.cell_tmp = c.take<int32_t>(N);
  w.cell_atoms = c.take<int32_t>(N);
  w.cell_pos = c.take<float4>(N);
  w.rowcnt = c.take<int32_t>(N + 1);
  w.rowptr = c.take<int32_t>(N + 1);
  w.rowptr_c = c.take<int32_t>(N + 1);
  w.nbr_tmp = c.take<int32_t>(C);
  w.nbr = c.take<int32_t>(C);
  w.shift_tmp = c.take<int32_t>(C);
  w.shift = c.take<int32_t>(C);
  w.scan_scratch = c.take<int32_t>(scan_scratch_ints(N + 1));
  if (ws) *ws = w;
  return c.bytes();
}

namespace {

constexpr int LG = 8;  // lanes cooperating on one atom's candidate scan

__device__ __forceinline__ double wrap_unit_rn(double x) {
  // np.mod(x + 0.5, 1.0) - 0.5 with numpy/jnp semantics (fmod, then +1 if negative), unfused
  const double t = __dadd_rn(x, 0.5);
  double m = __dsub_rn(t, trunc(t));  // == fmod(t, 1.0), exact
  if (m < 0.0) m = __dadd_rn(m, 1.0);
  return __dsub_rn(m, 0.5);
}

__device__ __forceinline__ double metric_sq_minimage(const double* __restrict__ B, double ax, double ay, double az,
                                                     double bx, double by, double bz) {
  // oracle/neighbor_oracle.py:metric_sq_minimage, same operation order, no FMA contraction
  const double w0 = wrap_unit_rn(__dsub_rn(ax, bx));
  const double w1 = wrap_unit_rn(__dsub_rn(ay, by));
  const double w2 = wrap_unit_rn(__dsub_rn(az, bz));
  double d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    d[i] = __dadd_rn(__dadd_rn(__dmul_rn(B[3 * i + 0], w0), __dmul_rn(B[3 * i + 1], w1)), __dmul_rn(B[3 * i + 2], w2));
  return __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
}

__device__ __forceinline__ double dist_sq_image(const double* __restrict__ C /*rows = lattice vectors*/, double cx,
                                                double cy, double cz, double ix, double iy, double iz, int s0, int s1,
                                                int s2) {
  // oracle/neighbor_oracle.py:vesin_full_list: D = (r_c - r_i) + ((S0*c0k + S1*c1k) + S2*c2k)
  const double rc[3] = {cx, cy, cz}, ri[3] = {ix, iy, iz};
  double acc = 0.0;
  double D[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double t = __dadd_rn(__dadd_rn(__dmul_rn(double(s0), C[0 + k]), __dmul_rn(double(s1), C[3 + k])),
                               __dmul_rn(double(s2), C[6 + k]));
    D[k] = __dadd_rn(__dsub_rn(rc[k], ri[k]), t);
  }
  acc = __dadd_rn(__dadd_rn(__dmul_rn(D[0], D[0]), __dmul_rn(D[1], D[1])), __dmul_rn(D[2], D[2]));
  return acc;
}

__device__ void invert3(const double* a, double* inv) {
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                     a[2] * (a[3] * a[7] - a[4] * a[6]);
  const double id = 1.0 / det;
  inv[0] = (a[4] * a[8] - a[5] * a[7]) * id;
  inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) * id;
  inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) * id;
  inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// mode 0: min-image (box columns = lattice vectors); mode 1: images (cell rows = lattice vectors)
__global__ void k_nl_setup(const double* __restrict__ box_or_cell, double cutoff, int64_t n_atoms, int mode, int pbc,
                           NlGrid* __restrict__ grid) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  NlGrid g;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) g.B[3 * i + j] = mode == 0 ? box_or_cell[3 * i + j] : box_or_cell[3 * j + i];
  g.cutoff_sq = cutoff * cutoff;
  g.periodic = pbc;
  if (!pbc) {
    // open boundaries: one cell, no images (gas-phase molecules on this path are a few atoms)
    for (int k = 0; k < 3; ++k) { g.nc[k] = 1; g.m[k] = 0; }
    for (int i = 0; i < 9; ++i) g.Binv[i] = 0.0;
    g.ncell = 1;
    *grid = g;
    return;
  }
  invert3(g.B, g.Binv);
  int64_t total = 1;
  for (int k = 0; k < 3; ++k) {
    const double nrm = sqrt(g.Binv[3 * k] * g.Binv[3 * k] + g.Binv[3 * k + 1] * g.Binv[3 * k + 1] +
                            g.Binv[3 * k + 2] * g.Binv[3 * k + 2]);
    const double h = 1.0 / nrm;  // perpendicular height along fractional axis k
    int nc = int(floor(h / cutoff));
    if (nc < 1) nc = 1;
    if (nc > 1024) nc = 1024;
    if (mode == 0) {
      // min-image: the +-1 stencil is only duplicate-free with >= 3 cells; otherwise scan the axis
      if (nc < 3) nc = 1;
      g.m[k] = nc >= 3 ? 1 : 0;
    } else {
      const double width = h / nc;
      int m = int(ceil(cutoff / width));
      if (m < 1) m = 1;
      g.m[k] = m;
    }
    g.nc[k] = nc;
    total *= nc;
  }
  // cap the number of cells at n_atoms (arrays are sized by n_atoms): shrink the largest axis
  while (total > n_atoms && total > 1) {
    int kmax = 0;
    for (int k = 1; k < 3; ++k)
      if (g.nc[k] > g.nc[kmax]) kmax = k;
    total /= g.nc[kmax];
    g.nc[kmax] = g.nc[kmax] > 3 ? g.nc[kmax] - 1 : 1;
    if (mode == 0) {
      if (g.nc[kmax] < 3) { g.nc[kmax] = 1; g.m[kmax] = 0; }
    }
    total *= g.nc[kmax];
  }
  g.ncell = int(total);
  *grid = g;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_cell_assign(const double* __restrict__ pos, int64_t N,
                                                     const NlGrid* __restrict__ grid, int32_t* __restrict__ cell_of,
                                                     int32_t* __restrict__ wrap, int32_t* __restrict__ cell_cnt) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= N) return;
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  double f[3];
  if (MODE == 0 || !grid->periodic) {
    f[0] = x; f[1] = y; f[2] = z;
  } else {
    const double* Bi = grid->Binv;
    f[0] = Bi[0] * x + Bi[1] * y + Bi[2] * z;
    f[1] = Bi[3] * x + Bi[4] * y + Bi[5] * z;
    f[2] = Bi[6] * x + Bi[7] * y + Bi[8] * z;
  }
  int c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double fl = floor(f[k]);
    const int nc = grid->nc[k];
    int ck = int((f[k] - fl) * nc);
    ck = ck < 0 ? 0 : (ck >= nc ? nc - 1 : ck);
    c[k] = ck;
    if (MODE == 1) wrap[3 * i + k] = grid->periodic ? int(fl) : 0;
  }
  const int cid = (c[0] * grid->nc[1] + c[1]) * grid->nc[2] + c[2];
  cell_of[i] = cid;
  atomicAdd(&cell_cnt[cid], 1);
}

__global__ void __launch_bounds__(256) k_cell_fill(int64_t N, const int32_t* __restrict__ cell_of,
                                                   const int32_t* __restrict__ cellptr, int32_t* __restrict__ cursor,
                                                   int32_t* __restrict__ cell_tmp) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= N) return;
  const int cid = cell_of[i];
  cell_tmp[cellptr[cid] + atomicAdd(&cursor[cid], 1)] = int(i);
}

__device__ __forceinline__ int pack_shift(int s0, int s1, int s2) {
  return ((s0 + 128) << 16) | ((s1 + 128) << 8) | (s2 + 128);
}
__device__ __forceinline__ void unpack_shift(int p, int& s0, int& s1, int& s2) {
  s0 = ((p >> 16) & 255) - 128; s1 = ((p >> 8) & 255) - 128; s2 = (p & 255) - 128;
}

// cell-ordered fp32 copy of the wrapped fractional positions: the candidate stream of the pre-filter below
__global__ void __launch_bounds__(256) k_cell_pack(const double* __restrict__ pos, int64_t N,
                                                   const int32_t* __restrict__ cell_atoms, float4* __restrict__ cell_pos) {
  const int64_t t = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (t >= N) return;
  const int j = cell_atoms[t];
  const double x = pos[3 * int64_t(j)], y = pos[3 * int64_t(j) + 1], z = pos[3 * int64_t(j) + 2];
  cell_pos[t] = make_float4(float(x - floor(x)), float(y - floor(y)), float(z - floor(z)), __int_as_float(j));
}

// Candidate scan of one atom by LG lanes (order of the accepted pairs is irrelevant, rows are sorted afterwards).
//   PASS 0: count accepted pairs.
//   PASS 1: append them at rowptr[atom] (second pass of the two-pass build; rows whose count is <= stride were
//           already captured by a PASS 2 scan and are skipped).
//   PASS 2: count AND park the first `stride` accepted neighbours of every atom in a fixed-stride scratch row, so
//           that one scan suffices whenever no row exceeds the stride.
// LGT lanes per atom: LG for large systems; a whole warp for the few-hundred-atom image lists, whose 27+ image cells x N
// candidates per atom are otherwise walked by eight lanes (96 atoms: 84 + 107 us of a 0.47 ms calculator call).
template <int MODE, int PASS, int LGT = LG>
__global__ void __launch_bounds__(256) k_nl_scan(const double* __restrict__ pos, int64_t N,
                                                 const NlGrid* __restrict__ gridp, NlWorkspace w, int stride,
                                                 int32_t* __restrict__ scratch) {
  constexpr bool FILL = PASS != 0;
  __shared__ NlGrid g;
  __shared__ int s_cnt[256 / LGT];   // PASS 1 / 2: running slot counter of each lane group
  if (threadIdx.x == 0) g = *gridp;
  if (threadIdx.x < 256 / LGT) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t gid = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const bool valid = gid / LGT < N;
  const int64_t atom = valid ? gid / LGT : 0;  // out-of-range lanes stay alive for the shuffles below
  const int sub = int(gid % LGT);
  const double ax = pos[3 * atom], ay = pos[3 * atom + 1], az = pos[3 * atom + 2];
  int c[3];
  {
    int cid = w.cell_of[atom];
    c[2] = cid % g.nc[2]; cid /= g.nc[2];
    c[1] = cid % g.nc[1]; c[0] = cid / g.nc[1];
  }
  int wa[3] = {0, 0, 0};
  if (MODE == 1) { wa[0] = w.wrap[3 * atom]; wa[1] = w.wrap[3 * atom + 1]; wa[2] = w.wrap[3 * atom + 2]; }
  // fp32 pre-filter (min-image mode): candidates whose approximate distance is outside cutoff +- delta are decided
  // in fp32, only the thin shell around the cutoff pays for the exact fp64 predicate -- the pair set is unchanged.
  float fa[3] = {0.f, 0.f, 0.f}, Bf[9], c2_lo = 0.f, c2_hi = 0.f;
  if (MODE == 0) {
    fa[0] = float(ax - floor(ax)); fa[1] = float(ay - floor(ay)); fa[2] = float(az - floor(az));
    float bmax = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { Bf[i] = float(g.B[i]); bmax = fmaxf(bmax, fabsf(Bf[i])); }
    const float cut = sqrtf(float(g.cutoff_sq));
    const float delta = 1e-3f + 4e-6f * bmax;   // >> fp32 error of wrapped fractional differences times |box|
    c2_lo = (cut - delta) * (cut - delta);
    c2_hi = (cut + delta) * (cut + delta);
  }
  const int grp = threadIdx.x / LGT;
  const int64_t base = PASS == 1 ? int64_t(w.rowptr[atom]) : PASS == 2 ? atom * int64_t(stride) : 0;
  int count = 0;
  int m0 = (valid && !(PASS == 1 && w.rowcnt[atom] <= stride)) ? g.m[0] : -1;
  // what happens to an accepted candidate (both scan forms)
  auto accept = [&](int j, int packed) {
    if (FILL) {
      const int slot = atomicAdd(&s_cnt[grp], 1);   // shared-memory counter of the lane group
      if (PASS == 2) {
        if (slot < stride) scratch[base + slot] = j;
      } else if (base + slot < w.capacity) {
        w.nbr_tmp[base + slot] = j;
        if (MODE == 1) w.shift_tmp[base + slot] = packed;
      }
    } else {
      ++count;
    }
  };
  // Min-image scan over at most 27 cells, software-pipelined: the lane group first fetches all cell ranges at once (one
  // exposed latency instead of 27 dependent ones), then every lane walks its share of the candidates as ONE stream with the
  // next candidate's record already in flight while the current one is tested (ncu: 38 % of this kernel's stall samples sat
  // on the first use of the just-issued candidate load).  Same candidates and the same predicate as the loop nest below; the
  // order in which accepted pairs land is irrelevant, rows are sorted afterwards.
  if (MODE == 0) {
    constexpr int kMaxCells = 27;
    const int n1 = 2 * g.m[1] + 1, n2 = 2 * g.m[2] + 1;
    const int ncell = (2 * g.m[0] + 1) * n1 * n2;   // uniform over the grid
    if (ncell <= kMaxCells) {
      __shared__ int s_cb[256 / LGT][kMaxCells], s_ce[256 / LGT][kMaxCells];
      if (m0 >= 0) {
        for (int cc = sub; cc < ncell; cc += LGT) {
          const int q0 = c[0] + cc / (n2 * n1) - g.m[0], q1 = c[1] + (cc / n2) % n1 - g.m[1], q2 = c[2] + cc % n2 - g.m[2];
          const int s0 = q0 >= 0 ? q0 / g.nc[0] : -((-q0 + g.nc[0] - 1) / g.nc[0]);
          const int s1 = q1 >= 0 ? q1 / g.nc[1] : -((-q1 + g.nc[1] - 1) / g.nc[1]);
          const int s2 = q2 >= 0 ? q2 / g.nc[2] : -((-q2 + g.nc[2] - 1) / g.nc[2]);
          const int cid = ((q0 - s0 * g.nc[0]) * g.nc[1] + (q1 - s1 * g.nc[1])) * g.nc[2] + (q2 - s2 * g.nc[2]);
          s_cb[grp][cc] = w.cellptr[cid];
          s_ce[grp][cc] = w.cellptr[cid + 1];
        }
      }
      __syncwarp();
      if (m0 >= 0) {
        int cc = 0, t = s_cb[grp][0] + sub, ce = s_ce[grp][0];
        while (t >= ce && ++cc < ncell) { t = s_cb[grp][cc] + sub; ce = s_ce[grp][cc]; }
        float4 q_nxt = cc < ncell ? w.cell_pos[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        while (cc < ncell) {
          const float4 q = q_nxt;
          t += LGT;
          while (t >= ce && ++cc < ncell) { t = s_cb[grp][cc] + sub; ce = s_ce[grp][cc]; }
          if (cc < ncell) q_nxt = w.cell_pos[t];
          const int j = __float_as_int(q.w);
          float u0 = fa[0] - q.x, u1 = fa[1] - q.y, u2 = fa[2] - q.z;
          u0 -= rintf(u0); u1 -= rintf(u1); u2 -= rintf(u2);
          const float e0 = Bf[0] * u0 + Bf[1] * u1 + Bf[2] * u2;
          const float e1 = Bf[3] * u0 + Bf[4] * u1 + Bf[5] * u2;
          const float e2 = Bf[6] * u0 + Bf[7] * u1 + Bf[8] * u2;
          const float d2f = e0 * e0 + e1 * e1 + e2 * e2;
          bool keep;
          if (d2f > c2_hi || j == atom) {
            keep = false;
          } else if (d2f < c2_lo) {
            keep = true;
          } else {
            const double bx = pos[3 * int64_t(j)], by = pos[3 * int64_t(j) + 1], bz = pos[3 * int64_t(j) + 2];
            keep = metric_sq_minimage(g.B, ax, ay, az, bx, by, bz) < g.cutoff_sq;
          }
          if (keep) accept(j, 0);
        }
      }
      m0 = -1;   // done: the loop nest below is skipped
    }
  }
  for (int d0 = -m0; d0 <= m0; ++d0) {
    const int q0 = c[0] + d0;
    const int s0 = q0 >= 0 ? q0 / g.nc[0] : -((-q0 + g.nc[0] - 1) / g.nc[0]);
    const int k0 = q0 - s0 * g.nc[0];
    for (int d1 = -g.m[1]; d1 <= g.m[1]; ++d1) {
      const int q1 = c[1] + d1;
      const int s1 = q1 >= 0 ? q1 / g.nc[1] : -((-q1 + g.nc[1] - 1) / g.nc[1]);
      const int k1 = q1 - s1 * g.nc[1];
      for (int d2 = -g.m[2]; d2 <= g.m[2]; ++d2) {
        const int q2 = c[2] + d2;
        const int s2 = q2 >= 0 ? q2 / g.nc[2] : -((-q2 + g.nc[2] - 1) / g.nc[2]);
        const int k2 = q2 - s2 * g.nc[2];
        const int cid = (k0 * g.nc[1] + k1) * g.nc[2] + k2;
        const int cb = w.cellptr[cid], ce = w.cellptr[cid + 1];
        for (int t = cb + sub; t < ce; t += LGT) {
          int j = 0;
          bool keep = false;
          int packed = 0;
          if (MODE == 0) {
            const float4 q = w.cell_pos[t];
            j = __float_as_int(q.w);
            float u0 = fa[0] - q.x, u1 = fa[1] - q.y, u2 = fa[2] - q.z;
            u0 -= rintf(u0); u1 -= rintf(u1); u2 -= rintf(u2);
            const float e0 = Bf[0] * u0 + Bf[1] * u1 + Bf[2] * u2;
            const float e1 = Bf[3] * u0 + Bf[4] * u1 + Bf[5] * u2;
            const float e2 = Bf[6] * u0 + Bf[7] * u1 + Bf[8] * u2;
            const float d2f = e0 * e0 + e1 * e1 + e2 * e2;
            if (d2f > c2_hi || j == atom) {
              keep = false;
            } else if (d2f < c2_lo) {
              keep = true;
            } else {
              const double bx = pos[3 * int64_t(j)], by = pos[3 * int64_t(j) + 1], bz = pos[3 * int64_t(j) + 2];
              keep = metric_sq_minimage(g.B, ax, ay, az, bx, by, bz) < g.cutoff_sq;
            }
          } else {
            j = w.cell_atoms[t];
            const double bx = pos[3 * int64_t(j)], by = pos[3 * int64_t(j) + 1], bz = pos[3 * int64_t(j) + 2];
            // S such that D = r_c - r_i + S.cell for the image reached through box shift (s0,s1,s2)
            const int S0 = w.wrap[3 * int64_t(j)] - wa[0] - s0;
            const int S1 = w.wrap[3 * int64_t(j) + 1] - wa[1] - s1;
            const int S2 = w.wrap[3 * int64_t(j) + 2] - wa[2] - s2;
            // g.B holds cell^T; dist_sq_image wants rows = lattice vectors: C[a][k] = B[k][a]
            double C[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int k = 0; k < 3; ++k) C[3 * a + k] = g.B[3 * k + a];
            keep = !(j == atom && S0 == 0 && S1 == 0 && S2 == 0) &&
                   dist_sq_image(C, ax, ay, az, bx, by, bz, S0, S1, S2) < g.cutoff_sq &&
                   S0 >= -127 && S0 <= 127 && S1 >= -127 && S1 <= 127 && S2 >= -127 && S2 <= 127;
            packed = pack_shift(S0, S1, S2);
          }
          if (keep) accept(j, packed);
        }
      }
    }
  }
  if (PASS == 0) {
#pragma unroll
    for (int o = LGT / 2; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (valid && sub == 0) w.rowcnt[atom] = count;
  } else if (PASS == 2) {
    __syncwarp();
    if (valid && sub == 0) w.rowcnt[atom] = s_cnt[grp];
  }
}

__global__ void __launch_bounds__(256) k_nl_finish_count(int64_t N, int64_t capacity, const int32_t* __restrict__ rowptr,
                                                         int32_t* __restrict__ rowptr_c, int32_t* __restrict__ n_found,
                                                         uint8_t* __restrict__ overflow) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i > N) return;
  const int v = rowptr[i];
  rowptr_c[i] = v < capacity ? v : int(capacity);
  if (i == N) {
    *n_found = v;
    if (v > capacity) *overflow = uint8_t(*overflow | 1);  // PEC.NEIGHBOR_LIST_OVERFLOW (partition.py:734)
  }
}

// rank sort of each row by (neighbour, packed shift); warp per row
__global__ void __launch_bounds__(256) k_nl_rowsort(const int32_t* __restrict__ rowptr_c, int64_t N,
                                                    const int32_t* __restrict__ nbr_in,
                                                    const int32_t* __restrict__ shift_in, int32_t* __restrict__ nbr_out,
                                                    int32_t* __restrict__ shift_out) {
  const int64_t row = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int b = rowptr_c[row], e = rowptr_c[row + 1];
  for (int i = b + lane; i < e; i += 32) {
    const int64_t key = (int64_t(nbr_in[i]) << 32) | uint32_t(shift_in ? shift_in[i] : 0);
    int rank = 0;
    for (int j = b; j < e; ++j) {
      const int64_t kj = (int64_t(nbr_in[j]) << 32) | uint32_t(shift_in ? shift_in[j] : 0);
      rank += (kj < key);
    }
    nbr_out[b + rank] = nbr_in[i];
    if (shift_in) shift_out[b + rank] = shift_in[i];
  }
}

// Row sort from the single-scan scratch (min-image lists): one warp per row.  Rows of up to 256 neighbours are sorted in
// registers by a bitonic network (element e = q * 32 + lane: exchanges at distance >= 32 are register swaps inside a lane,
// the others one shuffle each; padding keys INT_MAX sink to the end) -- 28 / 36 compare-exchange steps instead of the
// cnt^2 / 32 comparisons per lane of a rank sort (ncu: that kernel was issue-bound at 96 %, 43 k instructions per row).
// Only the first len_c (capacity-clamped) elements are written, i.e. a truncated list keeps the FIRST pairs in the
// reference's order (partition.py:656-664).  Longer rows: rank sort straight from global memory.
constexpr int kRowSortMax = 256;

template <int NQ>
__device__ __forceinline__ void warp_bitonic_sort(int (&k)[NQ], int lane) {
#pragma unroll
  for (int size = 2; size <= NQ * 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int dq = stride >> 5;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int pq = q ^ dq;
          if (q < pq) {
            const bool up = ((q * 32) & size) == 0;          // bit `size` of e lives in q here (size >= 64)
            const int lo = min(k[q], k[pq]), hi = max(k[q], k[pq]);
            k[q] = up ? lo : hi;
            k[pq] = up ? hi : lo;
          }
        }
      } else {
        const bool lower = (lane & stride) == 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int other = __shfl_xor_sync(0xffffffffu, k[q], stride);
          const bool up = ((q * 32 + lane) & size) == 0;
          k[q] = (lower == up) ? min(k[q], other) : max(k[q], other);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_nl_rowsort_strided(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ rowptr_c,
                                                            const int32_t* __restrict__ rowcnt, int64_t N,
                                                            const int32_t* __restrict__ scratch, int stride,
                                                            const int32_t* __restrict__ nbr_tmp,
                                                            int32_t* __restrict__ nbr_out) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + wib;
  if (row >= N) return;
  const int cnt = rowcnt[row];
  const int b = rowptr_c[row], len_c = rowptr_c[row + 1] - b;
  const int32_t* src = cnt <= stride ? scratch + row * int64_t(stride) : nbr_tmp + rowptr[row];
  if (cnt <= 128) {
    int k[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = q * 32 + lane < cnt ? src[q * 32 + lane] : 0x7fffffff;
    warp_bitonic_sort<4>(k, lane);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q * 32 + lane < len_c) nbr_out[b + q * 32 + lane] = k[q];      // len_c <= cnt: padding keys are never written
  } else if (cnt <= kRowSortMax) {
    int k[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) k[q] = q * 32 + lane < cnt ? src[q * 32 + lane] : 0x7fffffff;
    warp_bitonic_sort<8>(k, lane);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q * 32 + lane < len_c) nbr_out[b + q * 32 + lane] = k[q];
  } else {   // very long rows: rank straight from global memory
    for (int i = lane; i < cnt; i += 32) {
      const int key = src[i];
      int rank = 0;
      for (int j = 0; j < cnt; ++j) rank += (src[j] < key);
      if (rank < len_c) nbr_out[b + rank] = key;
    }
  }
}

// COO emission.  pad_value: n_atoms for the JaxMD list (partition.py:656), 0 for the vesin path
// (ase_calc.py:349-355).
__global__ void __launch_bounds__(256) k_nl_emit(const int32_t* __restrict__ rowptr_c, int64_t N, int64_t capacity,
                                                 const int32_t* __restrict__ nbr, const int32_t* __restrict__ shift,
                                                 const NlGrid* __restrict__ grid, int pad_value,
                                                 int32_t* __restrict__ idx, int32_t* __restrict__ shifts_out,
                                                 double* __restrict__ offsets_out) {
  const int64_t wid = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid < N) {
    const int b = rowptr_c[wid], e = rowptr_c[wid + 1];
    for (int s = b + lane; s < e; s += 32) {
      idx[s] = nbr[s];
      idx[capacity + s] = int(wid);
      if (shift) {
        int s0, s1, s2;
        unpack_shift(shift[s], s0, s1, s2);
        if (shifts_out) { shifts_out[3 * int64_t(s)] = s0; shifts_out[3 * int64_t(s) + 1] = s1; shifts_out[3 * int64_t(s) + 2] = s2; }
        if (offsets_out) {
          // offsets = S @ cell (ase_calc.py:354); cell[a][k] = B[k][a]
#pragma unroll
          for (int k = 0; k < 3; ++k)
            offsets_out[3 * int64_t(s) + k] =
                __dadd_rn(__dadd_rn(__dmul_rn(double(s0), grid->B[3 * k + 0]), __dmul_rn(double(s1), grid->B[3 * k + 1])),
                          __dmul_rn(double(s2), grid->B[3 * k + 2]));
        }
      }
    }
  }
  // tail padding, grid-stride over [n_valid, capacity)
  const int64_t n_valid = rowptr_c[N];
  const int64_t tid = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * 256;
  for (int64_t s = n_valid + tid; s < capacity; s += nthreads) {
    idx[s] = pad_value;
    idx[capacity + s] = pad_value;
    if (shifts_out) { shifts_out[3 * s] = 0; shifts_out[3 * s + 1] = 0; shifts_out[3 * s + 2] = 0; }
    if (offsets_out) { offsets_out[3 * s] = 0.0; offsets_out[3 * s + 1] = 0.0; offsets_out[3 * s + 2] = 0.0; }
  }
}

__global__ void __launch_bounds__(256) k_nl_offsets(const int32_t* __restrict__ rowptr_c, int64_t N,
                                                    const int32_t* __restrict__ shift, const NlGrid* __restrict__ grid,
                                                    double* __restrict__ offsets) {
  const int64_t s = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (s >= rowptr_c[N]) return;
  int s0, s1, s2;
  unpack_shift(shift[s], s0, s1, s2);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    offsets[3 * s + k] =
        __dadd_rn(__dadd_rn(__dmul_rn(double(s0), grid->B[3 * k + 0]), __dmul_rn(double(s1), grid->B[3 * k + 1])),
                  __dmul_rn(double(s2), grid->B[3 * k + 2]));
}

__global__ void __launch_bounds__(256) k_needs_rebuild(const double* __restrict__ frac, const double* __restrict__ ref,
                                                       const double* __restrict__ box, int64_t N, double thr_sq,
                                                       int32_t* __restrict__ flag) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= N) return;
  double B[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) B[k] = box[k];
  const double d2 = metric_sq_minimage(B, frac[3 * i], frac[3 * i + 1], frac[3 * i + 2], ref[3 * i], ref[3 * i + 1],
                                       ref[3 * i + 2]);
  if (d2 > thr_sq) *flag = 1;  // any(d > (dr_threshold/2)^2) (partition.py:771-773)
}

}  // namespace

template <int MODE>
static int nl_build_common(cudaStream_t stream, const double* pos, const double* box_or_cell, int pbc, int64_t N,
                           double cutoff, int64_t capacity, const NlWorkspace& w, int32_t* idx, int32_t* shifts,
                           double* offsets, int32_t* n_found, uint8_t* overflow) {
  APAX_LAUNCH(k_nl_setup, dim3(1), dim3(32), 0, stream, box_or_cell, cutoff, N, MODE, pbc, w.grid);
  APAX_CUDA_TRY(cudaMemsetAsync(w.cell_cnt, 0, sizeof(int32_t) * (N + 1), stream));
  APAX_LAUNCH(k_cell_assign<MODE>, dim3((unsigned)cdiv(N, 256)), dim3(256), 0, stream, pos, N, w.grid, w.cell_of, w.wrap,
              w.cell_cnt);
  APAX_TRY(exclusive_scan_i32(stream, w.cell_cnt, w.cellptr, N, w.scan_scratch));
  APAX_CUDA_TRY(cudaMemsetAsync(w.cell_cnt, 0, sizeof(int32_t) * (N + 1), stream));
  APAX_LAUNCH(k_cell_fill, dim3((unsigned)cdiv(N, 256)), dim3(256), 0, stream, N, w.cell_of, w.cellptr, w.cell_cnt,
              w.cell_tmp);
  APAX_TRY(rowsort_i32(stream, w.cellptr, N, w.cell_tmp, nullptr, w.cell_atoms, nullptr));
  if (MODE == 0) APAX_LAUNCH(k_cell_pack, dim3((unsigned)cdiv(N, 256)), dim3(256), 0, stream, pos, N, w.cell_atoms, w.cell_pos);
  const bool wide = MODE == 1 && N <= 2048;      // small image lists: a warp per atom
  const dim3 grid_scan((unsigned)cdiv(N * (wide ? 32 : LG), 256));
  // Min-image lists carry no shifts, so the two shift buffers (2 x capacity ints, contiguous up to alignment padding)
  // serve as fixed-stride scratch rows and ONE candidate scan both counts and captures the neighbours; only rows
  // longer than the stride (none for ordinary densities: the stride is ~twice the mean row) are re-scanned.
  int stride = 0;
  int32_t* scratch = nullptr;
  if (MODE == 0) {
    const int64_t span = (reinterpret_cast<char*>(w.shift) - reinterpret_cast<char*>(w.shift_tmp)) / int64_t(sizeof(int32_t)) + capacity;
    const int64_t per_row = span / N;
    if (per_row >= 32) {
      stride = int(per_row < kRowSortMax ? per_row : kRowSortMax);
      scratch = w.shift_tmp;
    }
  }
  if (stride > 0) {
    auto k_nl_scan_single = k_nl_scan<MODE, 2>;
    APAX_LAUNCH(k_nl_scan_single, grid_scan, dim3(256), 0, stream, pos, N, w.grid, w, stride, scratch);
  } else {
    APAX_CUDA_TRY(cudaMemsetAsync(w.rowcnt, 0, sizeof(int32_t) * (N + 1), stream));
    auto k_nl_scan_count = wide ? k_nl_scan<MODE, 0, 32> : k_nl_scan<MODE, 0>;
    APAX_LAUNCH(k_nl_scan_count, grid_scan, dim3(256), 0, stream, pos, N, w.grid, w, 0, scratch);
  }
  APAX_TRY(exclusive_scan_i32(stream, w.rowcnt, w.rowptr, N, w.scan_scratch));
  APAX_LAUNCH(k_nl_finish_count, dim3((unsigned)cdiv(N + 1, 256)), dim3(256), 0, stream, N, capacity, w.rowptr,
              w.rowptr_c, n_found, overflow);
  {
    auto k_nl_scan_fill = wide ? k_nl_scan<MODE, 1, 32> : k_nl_scan<MODE, 1>;   // with a stride: only the rows the single scan could not hold
    APAX_LAUNCH(k_nl_scan_fill, grid_scan, dim3(256), 0, stream, pos, N, w.grid, w, stride > 0 ? stride : -1, scratch);
  }
  if (stride > 0) {
    APAX_LAUNCH(k_nl_rowsort_strided, dim3((unsigned)cdiv(N, 8)), dim3(256), 0, stream, w.rowptr, w.rowptr_c, w.rowcnt, N,
                scratch, stride, w.nbr_tmp, w.nbr);
  } else {
    APAX_LAUNCH(k_nl_rowsort, dim3((unsigned)cdiv(N * 32, 256)), dim3(256), 0, stream, w.rowptr_c, N, w.nbr_tmp,
                MODE == 1 ? w.shift_tmp : nullptr, w.nbr, MODE == 1 ? w.shift : nullptr);
  }
  if (idx) {
    const int64_t blocks = cdiv(N * 32, 256);
    APAX_LAUNCH(k_nl_emit, dim3((unsigned)blocks), dim3(256), 0, stream, w.rowptr_c, N, capacity, w.nbr,
                MODE == 1 ? w.shift : nullptr, w.grid, MODE == 0 ? int(N) : 0, idx, shifts, offsets);
  }
  return APAX_OK;
}

int nl_emit_offsets(cudaStream_t stream, const NlWorkspace& w, int64_t n_atoms, double* offsets) {
  APAX_LAUNCH(k_nl_offsets, dim3((unsigned)cdiv(w.capacity, 256)), dim3(256), 0, stream, w.rowptr_c, n_atoms, w.shift,
              w.grid, offsets);
  return APAX_OK;
}

int nl_build_minimage_csr(cudaStream_t stream, const double* frac, const double* box, int64_t n_atoms, double cutoff,
                          int64_t capacity, const NlWorkspace& w, int32_t* idx, int32_t* n_found, uint8_t* overflow) {
  return nl_build_common<0>(stream, frac, box, 1, n_atoms, cutoff, capacity, w, idx, nullptr, nullptr, n_found,
                            overflow);
}

int nl_build_images_csr(cudaStream_t stream, const double* pos, const double* cell, int32_t pbc, int64_t n_atoms,
                        double cutoff, int64_t capacity, const NlWorkspace& w, int32_t* idx, int32_t* shifts,
                        double* offsets, int32_t* n_found, uint8_t* overflow) {
  return nl_build_common<1>(stream, pos, cell, pbc, n_atoms, cutoff, capacity, w, idx, shifts, offsets, n_found,
                            overflow);
}

// ================================================================================================
// Batched image lists: many small independent periodic cells (BASELINE config 5: 4096 x 128 atoms,
// apax/md/ase_calc.py:395-481 evaluates them through vesin + padding).  Atoms of all structures are
// concatenated; every structure is scanned by brute force over its own atoms and image shifts with the
// same fp64 predicate as the single-structure path.  Rows (central atoms) come out in a fixed
// enumeration order, so no sort is needed and the CSR can be consumed directly.
// ================================================================================================
struct StructGrid {
  double C[9];     // rows = lattice vectors
  double Cinv[9];  // frac = pos . Cinv  (row-vector convention)
  int m[3];        // image range per axis
  int pad;
};

namespace {

__global__ void __launch_bounds__(128) k_nlb_setup(const double* __restrict__ pos, const double* __restrict__ cells,
                                                   const int32_t* __restrict__ struct_ptr, int n_structs, double cutoff,
                                                   StructGrid* __restrict__ grids, int32_t* __restrict__ sid,
                                                   int32_t* __restrict__ wrap) {
  const int b = blockIdx.x;
  if (b >= n_structs) return;
  __shared__ StructGrid g;
  if (threadIdx.x == 0) {
    double Bm[9], Binv[9];
    for (int i = 0; i < 9; ++i) g.C[i] = cells[9 * int64_t(b) + i];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Bm[3 * i + j] = g.C[3 * j + i];  // columns = lattice vectors
    invert3(Bm, Binv);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) g.Cinv[3 * j + i] = Binv[3 * i + j];  // frac_k = sum_j pos_j Cinv[j][k]
    for (int k = 0; k < 3; ++k) {
      const double nrm = sqrt(Binv[3 * k] * Binv[3 * k] + Binv[3 * k + 1] * Binv[3 * k + 1] + Binv[3 * k + 2] * Binv[3 * k + 2]);
      int m = int(ceil(cutoff * nrm));  // cutoff / height
      g.m[k] = m < 1 ? 1 : (m > 100 ? 100 : m);
    }
    g.pad = 0;
    grids[b] = g;
  }
  __syncthreads();
  for (int a = struct_ptr[b] + threadIdx.x; a < struct_ptr[b + 1]; a += blockDim.x) {
    sid[a] = b;
    const double x = pos[3 * int64_t(a)], y = pos[3 * int64_t(a) + 1], z = pos[3 * int64_t(a) + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k) wrap[3 * int64_t(a) + k] = int(floor(x * g.Cinv[k] + y * g.Cinv[3 + k] + z * g.Cinv[6 + k]));
  }
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_nlb_scan(const double* __restrict__ pos, int64_t N,
                                                  const int32_t* __restrict__ struct_ptr,
                                                  const StructGrid* __restrict__ grids, const int32_t* __restrict__ sid,
                                                  const int32_t* __restrict__ wrap, double cutoff_sq, int64_t capacity,
                                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ rowcnt,
                                                  int32_t* __restrict__ nbr, int32_t* __restrict__ shift) {
  const int64_t c = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (c >= N) return;
  const int b = sid[c];
  const StructGrid g = grids[b];
  const double cx = pos[3 * c], cy = pos[3 * c + 1], cz = pos[3 * c + 2];
  const int wc0 = wrap[3 * c], wc1 = wrap[3 * c + 1], wc2 = wrap[3 * c + 2];
  int count = 0;
  const int base = FILL ? rowptr[c] : 0;
  // fp32 pre-filter (as in the min-image scan): the displacement to the image with S = d is formed in fp64 once per partner
  // and rounded; the 27+ lattice translations are subtracted in fp32.  Candidates whose approximate distance is outside
  // cutoff +- delta are decided there, only the thin shell around the cutoff pays for the exact fp64 predicate -- the pair
  // set is unchanged.  delta >> the fp32 error of a few subtractions at the magnitude of the farthest image.
  float Cf[9], mag = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) Cf[k] = float(g.C[k]);
#pragma unroll
  for (int a = 0; a < 3; ++a) mag += float(g.m[a] + 2) * (fabsf(Cf[3 * a]) + fabsf(Cf[3 * a + 1]) + fabsf(Cf[3 * a + 2]));
  const float cut = sqrtf(float(cutoff_sq));
  const float delta = 1e-3f + 1e-5f * mag;
  const float c_in = fmaxf(cut - delta, 0.f);
  const float c2_lo = c_in * c_in, c2_hi = (cut + delta) * (cut + delta);
  for (int i = struct_ptr[b]; i < struct_ptr[b + 1]; ++i) {
    const double ix = pos[3 * int64_t(i)], iy = pos[3 * int64_t(i) + 1], iz = pos[3 * int64_t(i) + 2];
    const int d0 = wrap[3 * int64_t(i)] - wc0, d1 = wrap[3 * int64_t(i) + 1] - wc1, d2 = wrap[3 * int64_t(i) + 2] - wc2;
    const float bx = float((cx - ix) + (double(d0) * g.C[0] + double(d1) * g.C[3] + double(d2) * g.C[6]));
    const float by = float((cy - iy) + (double(d0) * g.C[1] + double(d1) * g.C[4] + double(d2) * g.C[7]));
    const float bz = float((cz - iz) + (double(d0) * g.C[2] + double(d1) * g.C[5] + double(d2) * g.C[8]));
    for (int n0 = -g.m[0]; n0 <= g.m[0]; ++n0) {
      const float t0x = bx - float(n0) * Cf[0], t0y = by - float(n0) * Cf[1], t0z = bz - float(n0) * Cf[2];
      for (int n1 = -g.m[1]; n1 <= g.m[1]; ++n1) {
        const float t1x = t0x - float(n1) * Cf[3], t1y = t0y - float(n1) * Cf[4], t1z = t0z - float(n1) * Cf[5];
        for (int n2 = -g.m[2]; n2 <= g.m[2]; ++n2) {
          const float ex = t1x - float(n2) * Cf[6], ey = t1y - float(n2) * Cf[7], ez = t1z - float(n2) * Cf[8];
          const float d2f = ex * ex + ey * ey + ez * ez;
          if (d2f > c2_hi) continue;
          const int S0 = d0 - n0, S1 = d1 - n1, S2 = d2 - n2;  // D = r_c - r_i + S.cell
          if (i == c && S0 == 0 && S1 == 0 && S2 == 0) continue;
          if (S0 < -127 || S0 > 127 || S1 < -127 || S1 > 127 || S2 < -127 || S2 > 127) continue;
          if (d2f < c2_lo || dist_sq_image(g.C, cx, cy, cz, ix, iy, iz, S0, S1, S2) < cutoff_sq) {
            if (FILL) {
              const int slot = base + count;
              if (slot < capacity) {
                nbr[slot] = i;
                shift[slot] = pack_shift(S0, S1, S2);
              }
            }
            ++count;
          }
        }
      }
    }
  }
  if (!FILL) rowcnt[c] = count;
}

__global__ void __launch_bounds__(256) k_nlb_emit(const int32_t* __restrict__ rowptr_c, int64_t N, int64_t capacity,
                                                  const int32_t* __restrict__ nbr, const int32_t* __restrict__ shift,
                                                  const StructGrid* __restrict__ grids, const int32_t* __restrict__ sid,
                                                  int32_t* __restrict__ idx, int32_t* __restrict__ shifts_out,
                                                  double* __restrict__ offsets_out) {
  const int64_t wid = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid < N) {
    const int b = rowptr_c[wid], e = rowptr_c[wid + 1];
    const double* C = grids[sid[wid]].C;
    for (int s = b + lane; s < e; s += 32) {
      if (idx) { idx[s] = nbr[s]; idx[capacity + s] = int(wid); }
      int s0, s1, s2;
      unpack_shift(shift[s], s0, s1, s2);
      if (shifts_out) { shifts_out[3 * int64_t(s)] = s0; shifts_out[3 * int64_t(s) + 1] = s1; shifts_out[3 * int64_t(s) + 2] = s2; }
      if (offsets_out) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          offsets_out[3 * int64_t(s) + k] = __dadd_rn(__dadd_rn(__dmul_rn(double(s0), C[k]), __dmul_rn(double(s1), C[3 + k])),
                                                      __dmul_rn(double(s2), C[6 + k]));
      }
    }
  }
  const int64_t n_valid = rowptr_c[N];
  const int64_t tid = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * 256;
  for (int64_t s = n_valid + tid; s < capacity; s += nthreads) {
    if (idx) { idx[s] = 0; idx[capacity + s] = 0; }
    if (shifts_out) { shifts_out[3 * s] = 0; shifts_out[3 * s + 1] = 0; shifts_out[3 * s + 2] = 0; }
    if (offsets_out) { offsets_out[3 * s] = 0.0; offsets_out[3 * s + 1] = 0.0; offsets_out[3 * s + 2] = 0.0; }
  }
}

}  // namespace

size_t carve_nlb_workspace(void* base, int64_t n_atoms, int64_t capacity, int64_t n_structs, NlWorkspace* ws,
                           StructGrid** grids) {
  size_t off = carve_nl_workspace(base, n_atoms, capacity, ws);
  Carver c(base);
  c.off = off;
  StructGrid* g = c.take<StructGrid>(n_structs > 0 ? n_structs : 1);
  if (grids) *grids = g;
  return c.bytes();
}

int nl_build_images_batched(cudaStream_t stream, const double* pos, const double* cells, const int32_t* struct_ptr,
                            int64_t n_structs, int64_t N, double cutoff, int64_t capacity, const NlWorkspace& w,
                            StructGrid* grids, int32_t* idx, int32_t* shifts, double* offsets, int32_t* n_found,
                            uint8_t* overflow) {
  APAX_LAUNCH(k_nlb_setup, dim3((unsigned)n_structs), dim3(128), 0, stream, pos, cells, struct_ptr, int(n_structs), cutoff,
              grids, w.cell_of, w.wrap);
  const dim3 grid((unsigned)cdiv(N, 128));
  {
    auto k_nlb_count = k_nlb_scan<false>;
    APAX_LAUNCH(k_nlb_count, grid, dim3(128), 0, stream, pos, N, struct_ptr, grids, w.cell_of, w.wrap, cutoff * cutoff,
                capacity, w.rowptr, w.rowcnt, w.nbr, w.shift);
  }
  APAX_TRY(exclusive_scan_i32(stream, w.rowcnt, w.rowptr, N, w.scan_scratch));
  APAX_LAUNCH(k_nl_finish_count, dim3((unsigned)cdiv(N + 1, 256)), dim3(256), 0, stream, N, capacity, w.rowptr, w.rowptr_c,
              n_found, overflow);
  {
    auto k_nlb_fill = k_nlb_scan<true>;
    APAX_LAUNCH(k_nlb_fill, grid, dim3(128), 0, stream, pos, N, struct_ptr, grids, w.cell_of, w.wrap, cutoff * cutoff,
                capacity, w.rowptr, w.rowcnt, w.nbr, w.shift);
  }
  if (idx || shifts || offsets) {
    APAX_LAUNCH(k_nlb_emit, dim3((unsigned)cdiv(N * 32, 256)), dim3(256), 0, stream, w.rowptr_c, N, capacity, w.nbr,
                w.shift, grids, w.cell_of, idx, shifts, offsets);
  }
  return APAX_OK;
}

}  // namespace apax

using namespace apax;

extern "C" size_t apax_nl_workspace_bytes(int64_t n_atoms, int64_t capacity) {
  if (n_atoms < 0 || capacity < 0) return 0;
  return carve_nl_workspace(nullptr, n_atoms, capacity, nullptr);
}

static int nl_args_ok(int64_t n_atoms, int64_t capacity, double cutoff, void* ws, size_t bytes, NlWorkspace* w) {
  if (n_atoms < 1 || capacity < 0 || !(cutoff > 0.0) || !ws) return APAX_ERR_INVALID_ARG;
  if (n_atoms >= (int64_t(1) << 31) / 32 || capacity >= (int64_t(1) << 31) - 256) return APAX_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return APAX_ERR_INVALID_ARG;
  if (carve_nl_workspace(ws, n_atoms, capacity, w) > bytes) return APAX_ERR_WORKSPACE;
  return APAX_OK;
}

extern "C" int apax_nl_build_minimage(apax_stream_t stream, const double* frac, const double* box, int64_t n_atoms,
                                      double cutoff, int64_t capacity, int32_t* idx, int32_t* n_found,
                                      uint8_t* overflow, void* workspace, size_t workspace_bytes) {
  NlWorkspace w;
  APAX_TRY(nl_args_ok(n_atoms, capacity, cutoff, workspace, workspace_bytes, &w));
  if (!frac || !box || !idx || !n_found || !overflow) return APAX_ERR_INVALID_ARG;
  return nl_build_minimage_csr(stream, frac, box, n_atoms, cutoff, capacity, w, idx, n_found, overflow);
}

extern "C" int apax_nl_build_images(apax_stream_t stream, const double* pos, const double* cell, int32_t pbc,
                                    int64_t n_atoms, double cutoff, int64_t capacity, int32_t* idx, int32_t* shifts,
                                    double* offsets, int32_t* n_found, uint8_t* overflow, void* workspace,
                                    size_t workspace_bytes) {
  NlWorkspace w;
  APAX_TRY(nl_args_ok(n_atoms, capacity, cutoff, workspace, workspace_bytes, &w));
  if (!pos || !cell || !idx || !n_found || !overflow) return APAX_ERR_INVALID_ARG;
  return nl_build_images_csr(stream, pos, cell, pbc, n_atoms, cutoff, capacity, w, idx, shifts, offsets, n_found,
                             overflow);
}

extern "C" size_t apax_nl_batched_workspace_bytes(int64_t n_atoms, int64_t capacity, int64_t n_structs) {
  if (n_atoms < 0 || capacity < 0 || n_structs < 0) return 0;
  return carve_nlb_workspace(nullptr, n_atoms, capacity, n_structs, nullptr, nullptr);
}

extern "C" int apax_nl_build_images_batched(apax_stream_t stream, const double* pos, const double* cells,
                                            const int32_t* struct_ptr, int64_t n_structs, int64_t n_atoms,
                                            double cutoff, int64_t capacity, int32_t* idx, int32_t* shifts,
                                            double* offsets, int32_t* n_found, uint8_t* overflow, void* workspace,
                                            size_t workspace_bytes) {
  NlWorkspace w;
  if (n_structs < 1) return APAX_ERR_INVALID_ARG;
  APAX_TRY(nl_args_ok(n_atoms, capacity, cutoff, workspace, workspace_bytes, &w));
  StructGrid* grids = nullptr;
  if (carve_nlb_workspace(workspace, n_atoms, capacity, n_structs, &w, &grids) > workspace_bytes) return APAX_ERR_WORKSPACE;
  if (!pos || !cells || !struct_ptr || !n_found || !overflow) return APAX_ERR_INVALID_ARG;
  return nl_build_images_batched(stream, pos, cells, struct_ptr, n_structs, n_atoms, cutoff, capacity, w, grids, idx,
                                 shifts, offsets, n_found, overflow);
}

extern "C" int apax_nl_needs_rebuild(apax_stream_t stream, const double* frac, const double* ref, const double* box,
                                     int64_t n_atoms, double dr_threshold, int32_t* flag) {
  if (!frac || !ref || !box || !flag || n_atoms < 1) return APAX_ERR_INVALID_ARG;
  APAX_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int32_t), stream));
  const double half = dr_threshold / 2.0;
  APAX_LAUNCH(k_needs_rebuild, dim3((unsigned)cdiv(n_atoms, 256)), dim3(256), 0, stream, frac, ref, box, n_atoms,
              half * half, flag);
  return APAX_OK;
}
