// Training step of the GMNN path (BASELINE configs[1]; SURVEY.md s8e "Training", s8f rank 1): the loss of a batch of
// independent structures and its gradient w.r.t. every parameter, and the optax-Adam update -- hand-derived, no autodiff.
//
// Replaces, for the GMNN model with an energy + force MSE loss:
//   calc_loss + jax.value_and_grad(params)      apax/train/trainer.py:253-263, 332-336
//   jax.vmap(model.apply) over structures       apax/train/run.py:204
//   Loss.__call__ (mse, atoms_exponent)         apax/train/loss.py:12-23, 199-244
//   get_opt (clip -> adam(linear schedule) -> zero_nans per parameter group)   apax/optimizer/get_optimizer.py:62-176
//
// The force loss differentiates F = -dE/dR w.r.t. the weights.  With u = dL/dF (a fixed vector once E and F of the batch are
// known) and alpha_s = dL/dE_s,   grad_theta L = grad_theta [ sum_s alpha_s E_s  -  d/d(eps) E(R + eps u) ],
// so one FORWARD-mode (tangent) sweep of the energy along u followed by one reverse sweep of the dual program gives every
// gradient; the adjoint of the tangent quantities is minus the adjoint chain that produced the forces, so it is reused.
//
// Workload shape: hundreds of atoms per step (32 x 9-atom molecules per GPU), i.e. launch-latency bound.  All kernels are
// plain fp32 CUDA-core code in atom-major layout, deterministic (no floating-point atomics), never allocate and never
// synchronise, so a whole step (loss+grad, all-reduce, Adam) can be captured in a CUDA graph; the learning-rate schedule is
// evaluated on the device from a device-resident step counter for the same reason.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <vector>

#include "common.cuh"
#include "gm_internal.cuh"

namespace apax {
namespace {

constexpr int ONE_SLOT = NPK;           // pseudo-moment that is identically 1 (tangent 0)
constexpr float kSwish = 1.6765324703310907f;   // apax/layers/activation.py:8
constexpr int TR_CH = 32;               // pairs staged per chunk in the moment kernels

// ---------------------------------------------------------------------------------------------------------------------
// Contraction program as a flat table of monomials in the packed moments (built on the host, once per plan).
// feature f = sum_t coef_t * M[a_t] * M[b_t] * M[c_t]   (slot 100 == 1 for lower degrees)
// Reference: apax/layers/descriptor/gaussian_moment_descriptor.py:56-101, triangular_indices.py:28-49.
// ---------------------------------------------------------------------------------------------------------------------
struct Term {
  int f;
  float coef;
  uint8_t a, b, c;
};

int pack2(int i, int j) {
  if (i > j) std::swap(i, j);
  static const int t[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  return t[i][j];
}
int pack3(int i, int j, int k) {
  int v[3] = {i, j, k};
  std::sort(v, v + 3);
  static const int order[10][3] = {{0, 0, 0}, {0, 0, 1}, {0, 0, 2}, {0, 1, 1}, {0, 1, 2},
                                   {0, 2, 2}, {1, 1, 1}, {1, 1, 2}, {1, 2, 2}, {2, 2, 2}};
  for (int q = 0; q < 10; ++q)
    if (order[q][0] == v[0] && order[q][1] == v[1] && order[q][2] == v[2]) return q;
  return -1;
}
int s0(int r) { return r * 20; }
int s1(int r, int i) { return r * 20 + 1 + i; }
int s2(int r, int i, int j) { return r * 20 + 4 + pack2(i, j); }
int s3(int r, int i, int j, int k) { return r * 20 + 10 + pack3(i, j, k); }

std::vector<Term> build_terms() {
  std::vector<std::array<int, 2>> t2;
  std::vector<std::array<int, 3>> t3;
  for (int i = 0; i < NR; ++i) {
    t2.push_back({i, i});
    for (int j = i + 1; j < NR; ++j) t2.push_back({i, j});
  }
  for (int i = 0; i < NR; ++i) {
    t3.push_back({i, i, i});
    for (int j = 0; j < NR; ++j)
      if (j != i) t3.push_back({i, j, j});
    for (int j = i + 1; j < NR; ++j)
      for (int k = j + 1; k < NR; ++k) t3.push_back({i, j, k});
  }
  std::map<std::array<int, 4>, int> acc;   // (feature, sorted slots) -> multiplicity
  auto add = [&](int f, int a, int b, int c) {
    int v[3] = {a, b, c};
    std::sort(v, v + 3);
    acc[{f, v[0], v[1], v[2]}] += 1;
  };
  int off = 0;
  for (int r = 0; r < NR; ++r) add(off + r, s0(r), ONE_SLOT, ONE_SLOT);   // c0
  off += NR;
  const int nt2 = int(t2.size()), nt3 = int(t3.size());
  for (int q = 0; q < nt2; ++q)   // c1 = ari,asi->ars
    for (int i = 0; i < 3; ++i) add(off + q, s1(t2[q][0], i), s1(t2[q][1], i), ONE_SLOT);
  off += nt2;
  for (int q = 0; q < nt2; ++q)   // c2 = arij,asij->ars
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) add(off + q, s2(t2[q][0], i, j), s2(t2[q][1], i, j), ONE_SLOT);
  off += nt2;
  for (int q = 0; q < nt2; ++q)   // c3 = arijk,asijk->ars
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) add(off + q, s3(t2[q][0], i, j, k), s3(t2[q][1], i, j, k), ONE_SLOT);
  off += nt2;
  for (int q = 0; q < nt3; ++q)   // c4 = arij,asik,atjk->arst
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) add(off + q, s2(t3[q][0], i, j), s2(t3[q][1], i, k), s2(t3[q][2], j, k));
  off += nt3;
  for (int q = 0; q < nt2; ++q)   // c5 = ari,asj,atij->arst at the (r,s) pairs, all t
    for (int t = 0; t < NR; ++t)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) add(off + q * NR + t, s1(t2[q][0], i), s1(t2[q][1], j), s2(t, i, j));
  off += nt2 * NR;
  for (int q = 0; q < nt2; ++q)   // c6 = arijk,asijl,atkl->arst at the (r,s) pairs, all t
    for (int t = 0; t < NR; ++t)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 3; ++l) add(off + q * NR + t, s3(t2[q][0], i, j, k), s3(t2[q][1], i, j, l), s2(t, k, l));
  off += nt2 * NR;
  for (int r = 0; r < NR; ++r)    // c7 = arijk,asij,atk->arst, all (r,s,t)
    for (int s = 0; s < NR; ++s)
      for (int t = 0; t < NR; ++t)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) add(off + (r * NR + s) * NR + t, s3(r, i, j, k), s2(s, i, j), s1(t, k));
  off += NR * NR * NR;
  std::vector<Term> out;
  if (off != NF) return out;
  for (auto& kv : acc) out.push_back({kv.first[0], float(kv.second), uint8_t(kv.first[1]), uint8_t(kv.first[2]), uint8_t(kv.first[3])});
  return out;   // std::map order: by feature first
}

// device tables
struct Tables {
  // by feature
  const int32_t* f_ptr;     // [NF+1]
  const float* f_coef;
  const uchar4* f_slots;    // (a, b, c, -)
  // by moment slot: every occurrence of slot i in a term: d term / d M_i = coef * M[o1] * M[o2]
  const int32_t* o_ptr;     // [NPK+1]
  const float* o_coef;
  const int32_t* o_pack;    // f | o1 << 16 | o2 << 24
};

constexpr int MAXNB = 16;        // basis functions per radial channel: 7 Gaussians by default, Bessel's config default is 16
enum { BASIS_GAUSS_LINEAR = 0, BASIS_GAUSS_EXP = 1, BASIS_BESSEL = 2 };

struct DescConstT {
  float r_max, beta, norm, two_beta, pi_f, inv_sqrt_nb;
  int nb, kind;
  float mu[MAXNB];   // Gaussian centres; Bessel: prefactors a_n b_n (basis_functions.py:74-75)
};

}  // namespace
}  // namespace apax

struct apax_train_plan {
  apax_gm_config cfg;
  apax::Tables tab;
  apax::DescConstT dc;
  void* dev_block;          // one allocation holding all tables
  int64_t off32[8];         // emb, w0, b0, w1, b1, w2, b2, total
  int64_t off64[3];         // scale, shift, total
};

namespace apax {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------------------------
struct TrWs {
  int64_t n, P, B;
  int32_t *rs, *re, *ts, *te;     // [n] row / transpose-row ranges into rowp / trowp
  int32_t *rowp, *trowp;          // [P] COO positions grouped by central atom / by neighbour atom
  float4* geo;                    // [P] (dx,dy,dz, species-pair key as int bits); w < 0: dropped entry
  float4* Q;                      // [P] per-pair force contribution
  float* cbar;                    // [P][5 n_basis] per-pair embedding-gradient contributions
  int32_t* present;               // [128] species present, [128] ticket of the loss reduction
  float *M, *Md, *Mdb, *Mb2;      // [n][100]
  float *GS, *GB;                 // [2n][360]
  float *Y1, *H1, *Y2, *H2, *YB2, *HB1, *YB1;   // [2n][256]
  float* e_atom;                  // [2n]: readout output e_a, then its tangent
  double* Es;                     // [B]
  double* alpha;                  // [B]   dL/dE_s
  double* cF;                     // [B]   force-loss coefficient of the structure
  double* u;                      // [n][3] dL/dF
  double* loss_s;                 // [B]
};

size_t carve_tr(void* base, int64_t n, int64_t P, int64_t B, int nb, TrWs* w) {
  Carver c(base);
  TrWs t{};
  t.n = n; t.P = P; t.B = B;
  t.rs = c.take<int32_t>(n); t.re = c.take<int32_t>(n); t.ts = c.take<int32_t>(n); t.te = c.take<int32_t>(n);
  t.rowp = c.take<int32_t>(P); t.trowp = c.take<int32_t>(P);
  t.geo = c.take<float4>(P); t.Q = c.take<float4>(P);
  t.cbar = c.take<float>(P * NR * nb);
  t.present = c.take<int32_t>(132);
  t.M = c.take<float>(n * NPK); t.Md = c.take<float>(n * NPK); t.Mdb = c.take<float>(n * NPK); t.Mb2 = c.take<float>(n * NPK);
  t.GS = c.take<float>(2 * n * NF); t.GB = c.take<float>(2 * n * NF);
  t.Y1 = c.take<float>(2 * n * NH); t.H1 = c.take<float>(2 * n * NH); t.Y2 = c.take<float>(2 * n * NH);
  t.H2 = c.take<float>(2 * n * NH); t.YB2 = c.take<float>(2 * n * NH); t.HB1 = c.take<float>(2 * n * NH);
  t.YB1 = c.take<float>(2 * n * NH);
  t.e_atom = c.take<float>(2 * n);
  t.Es = c.take<double>(B); t.alpha = c.take<double>(B); t.cF = c.take<double>(B); t.u = c.take<double>(3 * n);
  t.loss_s = c.take<double>(B);
  if (w) *w = t;
  return c.bytes();
}

// ---------------------------------------------------------------------------------------------------------------------
// list grouping: one block per structure.  Every thread owns atoms of the structure and scans the structure's COO range:
// row (central == a) and transpose row (neighbour == a) in COO order, so all later sums have a fixed order.
// Dropped: entries with idx[0]==idx[1] (the (0,0) padding, apax/layers/masking.py:10-16) or an index outside the structure.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GR_STAGE = 3072;   // COO entries of a structure staged in shared memory (24 KB); longer lists are read in place
__global__ void __launch_bounds__(128) k_tr_group(const int32_t* __restrict__ idx, int64_t P, const int32_t* __restrict__ struct_ptr,
                                                  const int32_t* __restrict__ pair_ptr, const int32_t* __restrict__ Z,
                                                  const double* __restrict__ R, const double* __restrict__ offsets, int S,
                                                  int32_t* __restrict__ atom_struct, TrWs w) {
  __shared__ int32_t s_i0[GR_STAGE], s_i1[GR_STAGE];
  const int s = blockIdx.x;
  const int a0 = struct_ptr[s], a1 = struct_ptr[s + 1];
  const int p0 = pair_ptr[s], p1 = pair_ptr[s + 1], np = p1 - p0;
  const bool staged = np <= GR_STAGE;
  // per COO entry: fp64 displacement R[i1] - R[i0] (+ offsets) cast to fp32 (apax/layers/distances.py:48-67,
  // gaussian_moment_descriptor.py:33) and the embedding block of the pair: central species first (basis_functions.py:139-141)
  for (int q = threadIdx.x; q < np; q += blockDim.x) {
    const int p = p0 + q;
    const int j = idx[p], c = idx[P + p];
    if (staged) { s_i0[q] = j; s_i1[q] = c; }
    float4 g = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (j != c && j >= a0 && j < a1 && c >= a0 && c < a1) {
      double d0 = R[3 * int64_t(c) + 0] - R[3 * int64_t(j) + 0];
      double d1 = R[3 * int64_t(c) + 1] - R[3 * int64_t(j) + 1];
      double d2 = R[3 * int64_t(c) + 2] - R[3 * int64_t(j) + 2];
      if (offsets) { d0 += offsets[3 * int64_t(p) + 0]; d1 += offsets[3 * int64_t(p) + 1]; d2 += offsets[3 * int64_t(p) + 2]; }
      g = make_float4(float(d0), float(d1), float(d2), __int_as_float(Z[c] * S + Z[j]));
    }
    w.geo[p] = g;
  }
  __syncthreads();
  const int32_t* i0 = staged ? s_i0 : idx + p0;        // neighbour
  const int32_t* i1 = staged ? s_i1 : idx + P + p0;    // central
  for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
    atom_struct[a] = s;
    w.present[Z[a] & 127] = 1;    // benign race: every writer stores 1
    int nc = 0, nt = 0;
#pragma unroll 4
    for (int q = 0; q < np; ++q) {
      const int j = i0[q], c = i1[q];
      const bool ok = j != c && j >= a0 && j < a1 && c >= a0 && c < a1;
      nc += ok && c == a;
      nt += ok && j == a;
    }
    w.rs[a] = nc;
    w.ts[a] = nt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // serial prefix over the structure's atoms (tens to hundreds)
    int rc = p0, tc = p0;
    for (int a = a0; a < a1; ++a) {
      const int nc = w.rs[a], nt = w.ts[a];
      w.rs[a] = rc; w.re[a] = rc + nc; rc += nc;
      w.ts[a] = tc; w.te[a] = tc + nt; tc += nt;
    }
  }
  __syncthreads();
  for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
    int rc = w.rs[a], tc = w.ts[a];
    for (int q = 0; q < np; ++q) {
      const int j = i0[q], c = i1[q];
      const bool ok = j != c && j >= a0 && j < a1 && c >= a0 && c < a1;
      if (ok && c == a) w.rowp[rc++] = p0 + q;
      if (ok && j == a) w.trowp[tc++] = p0 + q;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-pair scalar pieces
// ---------------------------------------------------------------------------------------------------------------------
struct PairT {
  float r, inv_re;        // |d|, 1/(r + 1e-5)
  float n[3];             // d / (r + 1e-5)     (gaussian_moment_descriptor.py:46-48)
  float G[MAXNB], dG[MAXNB];   // radial basis and d/dr (basis_functions.py:19-79)
  float fc, dfc;               // cosine cutoff and d/dr (basis_functions.py:82-85; zero slope beyond r_max)
};

__device__ __forceinline__ void pair_terms(const float4 g, const DescConstT& dc, PairT& t) {
  const float r2 = g.x * g.x + g.y * g.y + g.z * g.z;
  t.r = r2 > 0.f ? sqrtf(r2) : 0.f;                      // space.distance / safe_mask (space.py:314-323)
  t.inv_re = 1.0f / (t.r + 1e-5f);
  t.n[0] = g.x * t.inv_re; t.n[1] = g.y * t.inv_re; t.n[2] = g.z * t.inv_re;
  if (dc.kind == BASIS_BESSEL) {
    // BesselBasis (basis_functions.py:59-79): a_n b_n (sinc((n+1) r/rc) + sinc((n+2) r/rc)), jnp.sinc(z) = sin(pi z)/(pi z);
    // d/dr sinc(k r/rc) = (cos(pi k r/rc) - sinc(k r/rc)) / r
    float sk[MAXNB + 1], dk[MAXNB + 1];
#pragma unroll
    for (int k = 0; k <= MAXNB; ++k) {
      if (k <= dc.nb) {
        const float y = dc.pi_f * (float(k + 1) * t.r / dc.r_max);
        float sn, cs;
        sincosf(y, &sn, &cs);
        sk[k] = y != 0.f ? sn / y : 1.0f;
        dk[k] = t.r > 0.f ? (cs - sk[k]) / t.r : 0.f;
      } else {
        sk[k] = 0.f; dk[k] = 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < MAXNB; ++b) {
      t.G[b] = b < dc.nb ? dc.mu[b] * (sk[b] + sk[b + 1]) : 0.f;
      t.dG[b] = b < dc.nb ? dc.mu[b] * (dk[b] + dk[b + 1]) : 0.f;
    }
  } else {
    // GaussianBasis: exp(-beta (mu_b - s(r))^2) with s(r) = r (linear spacing) or exp(-r) (exponential, :29-37)
    const float x = dc.kind == BASIS_GAUSS_EXP ? expf(-t.r) : t.r;
    const float dxdr = dc.kind == BASIS_GAUSS_EXP ? -x : 1.0f;
#pragma unroll
    for (int b = 0; b < MAXNB; ++b) {
      if (b < dc.nb) {
        const float dm = dc.mu[b] - x;
        t.G[b] = dc.norm * expf(-dc.beta * (dm * dm));
        t.dG[b] = t.G[b] * dc.two_beta * dm * dxdr;
      } else {
        t.G[b] = 0.f; t.dG[b] = 0.f;
      }
    }
  }
  if (t.r < dc.r_max) {
    float sn, cs;
    sincosf(dc.pi_f * t.r / dc.r_max, &sn, &cs);
    t.fc = 0.5f * (cs + 1.0f);
    t.dfc = -0.5f * (dc.pi_f / dc.r_max) * sn;
  } else {
    t.fc = 0.f;
    t.dfc = 0.f;
  }
}

__device__ __forceinline__ void monomials20(const float (&n)[3], float (&m)[20]) {
  const float x = n[0], y = n[1], z = n[2];
  m[0] = 1.0f; m[1] = x; m[2] = y; m[3] = z;
  const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  m[4] = xx; m[5] = xy; m[6] = xz; m[7] = yy; m[8] = yz; m[9] = zz;
  m[10] = xx * x; m[11] = xx * y; m[12] = xx * z; m[13] = xy * y; m[14] = xy * z;
  m[15] = xz * z; m[16] = yy * y; m[17] = yy * z; m[18] = yz * z; m[19] = zz * z;
}

// d mono_k / d n_i  (k as in monomials20)
__device__ __forceinline__ void monomial_grads(const float (&n)[3], float (&gx)[20], float (&gy)[20], float (&gz)[20]) {
  const float x = n[0], y = n[1], z = n[2];
  const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  //        1     x     y     z     xx      xy    xz    yy      yz    zz
  gx[0] = 0; gx[1] = 1; gx[2] = 0; gx[3] = 0; gx[4] = 2 * x; gx[5] = y; gx[6] = z; gx[7] = 0; gx[8] = 0; gx[9] = 0;
  gy[0] = 0; gy[1] = 0; gy[2] = 1; gy[3] = 0; gy[4] = 0; gy[5] = x; gy[6] = 0; gy[7] = 2 * y; gy[8] = z; gy[9] = 0;
  gz[0] = 0; gz[1] = 0; gz[2] = 0; gz[3] = 1; gz[4] = 0; gz[5] = 0; gz[6] = x; gz[7] = 0; gz[8] = y; gz[9] = 2 * z;
  //         xxx        xxy        xxz        xyy        xyz       xzz        yyy        yyz        yzz        zzz
  gx[10] = 3 * xx; gx[11] = 2 * xy; gx[12] = 2 * xz; gx[13] = yy; gx[14] = yz; gx[15] = zz; gx[16] = 0; gx[17] = 0; gx[18] = 0; gx[19] = 0;
  gy[10] = 0; gy[11] = xx; gy[12] = 0; gy[13] = 2 * xy; gy[14] = xz; gy[15] = 0; gy[16] = 3 * yy; gy[17] = 2 * yz; gy[18] = zz; gy[19] = 0;
  gz[10] = 0; gz[11] = 0; gz[12] = xx; gz[13] = 0; gz[14] = xy; gz[15] = 2 * xz; gz[16] = 0; gz[17] = yy; gz[18] = 2 * yz; gz[19] = 3 * zz;
}

// radial channels: rho_r = (sum_b c_rb G_b) * fc,  drho_r/dr;  c = (1/sqrt(n_basis)) * W[Z_central, Z_neighbour]
__device__ __forceinline__ void radial(const float* __restrict__ emb, int key, const DescConstT& dc, const PairT& t,
                                       float (&rho)[NR], float (&drho)[NR]) {
  const float* c = emb + int64_t(key) * (NR * dc.nb);
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float s = 0.f, ds = 0.f;
#pragma unroll
    for (int b = 0; b < MAXNB; ++b) {
      if (b < dc.nb) {
        const float cb = dc.inv_sqrt_nb * c[r * dc.nb + b];
        s += cb * t.G[b];
        ds += cb * t.dG[b];
      }
    }
    rho[r] = s * t.fc;
    drho[r] = ds * t.fc + s * t.dfc;
  }
}

// tangent of the pair geometry along ddot = u[central] - u[neighbour]
__device__ __forceinline__ void pair_tangent(const float4 g, const PairT& t, const double* __restrict__ u, int c, int j,
                                             float& rdot, float (&ndot)[3]) {
  const float e0 = float(u[3 * c + 0] - u[3 * j + 0]);
  const float e1 = float(u[3 * c + 1] - u[3 * j + 1]);
  const float e2 = float(u[3 * c + 2] - u[3 * j + 2]);
  rdot = t.r > 0.f ? (g.x * e0 + g.y * e1 + g.z * e2) / t.r : 0.f;
  const float k = rdot * t.inv_re * t.inv_re;
  ndot[0] = e0 * t.inv_re - g.x * k;
  ndot[1] = e1 * t.inv_re - g.y * k;
  ndot[2] = e2 * t.inv_re - g.z * k;
}

// ---------------------------------------------------------------------------------------------------------------------
// moments (moments.py:5-29) and their tangent: block per atom; pairs are staged chunk-wise (thread per pair), then
// thread (r, k) sums its slot over the chunk in list order.
//   TANGENT = 0:  M[a][r][k]  = sum_p rho_r mono_k
//   TANGENT = 1:  Md[a][r][k] = sum_p rhodot_r mono_k + rho_r monodot_k
// ---------------------------------------------------------------------------------------------------------------------
template <int TANGENT>
__global__ void __launch_bounds__(128) k_tr_moments(TrWs w, const float* __restrict__ emb, const int32_t* __restrict__ idx,
                                                    DescConstT dc, float* __restrict__ out) {
  __shared__ float s_rho[TR_CH][NR], s_mono[TR_CH][20];
  __shared__ float s_rhod[TANGENT ? TR_CH : 1][NR], s_monod[TANGENT ? TR_CH : 1][20];
  const int a = blockIdx.x;
  const int b0 = w.rs[a], b1 = w.re[a];
  const int t = threadIdx.x, r = t / 20, k = t % 20;
  float acc = 0.f;
  for (int base = b0; base < b1; base += TR_CH) {
    const int cnt = min(TR_CH, b1 - base);
    if (t < cnt) {
      const int p = w.rowp[base + t];
      const float4 g = w.geo[p];
      PairT pt;
      pair_terms(g, dc, pt);
      float rho[NR], drho[NR], mono[20];
      radial(emb, __float_as_int(g.w), dc, pt, rho, drho);
      monomials20(pt.n, mono);
#pragma unroll
      for (int q = 0; q < NR; ++q) s_rho[t][q] = rho[q];
#pragma unroll
      for (int q = 0; q < 20; ++q) s_mono[t][q] = mono[q];
      if (TANGENT) {
        float rdot, nd[3], gx[20], gy[20], gz[20];
        pair_tangent(g, pt, w.u, a, idx[p], rdot, nd);
        monomial_grads(pt.n, gx, gy, gz);
#pragma unroll
        for (int q = 0; q < NR; ++q) s_rhod[t][q] = drho[q] * rdot;
#pragma unroll
        for (int q = 0; q < 20; ++q) s_monod[t][q] = gx[q] * nd[0] + gy[q] * nd[1] + gz[q] * nd[2];
      }
    }
    __syncthreads();
    if (t < NPK) {
      for (int q = 0; q < cnt; ++q) {
        if (TANGENT) acc += s_rhod[q][r] * s_mono[q][k] + s_rho[q][r] * s_monod[q][k];
        else acc += s_rho[q][r] * s_mono[q][k];
      }
    }
    __syncthreads();
  }
  if (t < NPK) out[int64_t(a) * NPK + t] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// contractions: block per atom
//   fwd      g[f]  = sum_t coef M_a M_b M_c
//   tangent  gd[f] = sum_t coef (Md_a M_b M_c + M_a Md_b M_c + M_a M_b Md_c)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CT_ATOMS = 2;   // atoms per block: the term tables (80 KB / 240 KB) are walked once for both
template <int TANGENT>
__global__ void __launch_bounds__(384) k_tr_contract(Tables tab, const float* __restrict__ M, const float* __restrict__ Md,
                                                     float* __restrict__ out, int64_t n) {
  __shared__ float m[CT_ATOMS][NPK + 1], md[CT_ATOMS][NPK + 1];
  const int64_t a0 = int64_t(blockIdx.x) * CT_ATOMS;
  const int t = threadIdx.x;
  for (int e = t; e < CT_ATOMS * (NPK + 1); e += blockDim.x) {
    const int u = e / (NPK + 1), i = e % (NPK + 1);
    const bool ok = a0 + u < n && i < NPK;
    m[u][i] = i == NPK ? 1.f : (ok ? M[(a0 + u) * NPK + i] : 0.f);
    md[u][i] = (TANGENT && ok) ? Md[(a0 + u) * NPK + i] : 0.f;
  }
  __syncthreads();
  if (t >= NF) return;
  float acc[CT_ATOMS];
#pragma unroll
  for (int u = 0; u < CT_ATOMS; ++u) acc[u] = 0.f;
  const int q1 = tab.f_ptr[t + 1];
#pragma unroll 4
  for (int q = tab.f_ptr[t]; q < q1; ++q) {
    const uchar4 sl = tab.f_slots[q];
    const float c = tab.f_coef[q];
#pragma unroll
    for (int u = 0; u < CT_ATOMS; ++u) {
      if (TANGENT) acc[u] += c * (md[u][sl.x] * m[u][sl.y] * m[u][sl.z] + m[u][sl.x] * (md[u][sl.y] * m[u][sl.z] + m[u][sl.y] * md[u][sl.z]));
      else acc[u] += c * (m[u][sl.x] * m[u][sl.y] * m[u][sl.z]);
    }
  }
#pragma unroll
  for (int u = 0; u < CT_ATOMS; ++u)
    if (a0 + u < n) out[(a0 + u) * NF + t] = acc[u];
}

// reverse:  out[i] = sum_occ coef * ( gb1[f] * M[o1] M[o2]  +  gb2[f] * (Md[o1] M[o2] + M[o1] Md[o2]) )
// gb2 == nullptr: plain transpose-Jacobian product J(M)^T gb1.  With gb2 = adjoint of the tangent features this adds the
// second-order term d/d(eps) J(M + eps Md)^T gb2.
constexpr int CB_LANES = 8;   // lanes that share one moment slot's occurrence list
__global__ void __launch_bounds__(NPK * CB_LANES) k_tr_contract_bwd(Tables tab, const float* __restrict__ M, const float* __restrict__ Md,
                                                                   const float* __restrict__ gb1, const float* __restrict__ gb2,
                                                                   float* __restrict__ out, int64_t n) {
  __shared__ float m[CT_ATOMS][NPK + 1], md[CT_ATOMS][NPK + 1], g1[CT_ATOMS][NF], g2[CT_ATOMS][NF];
  const int64_t a0 = int64_t(blockIdx.x) * CT_ATOMS;
  const int t = threadIdx.x;
  const bool second = gb2 != nullptr;
  for (int e = t; e < CT_ATOMS * NF; e += blockDim.x) {
    const int u = e / NF, i = e % NF;
    const bool ok = a0 + u < n;
    g1[u][i] = ok ? gb1[(a0 + u) * NF + i] : 0.f;
    g2[u][i] = (second && ok) ? gb2[(a0 + u) * NF + i] : 0.f;
  }
  for (int e = t; e < CT_ATOMS * (NPK + 1); e += blockDim.x) {
    const int u = e / (NPK + 1), i = e % (NPK + 1);
    const bool ok = a0 + u < n && i < NPK;
    m[u][i] = i == NPK ? 1.f : (ok ? M[(a0 + u) * NPK + i] : 0.f);
    md[u][i] = (second && ok) ? Md[(a0 + u) * NPK + i] : 0.f;
  }
  __syncthreads();
  const int slot = t / CB_LANES, sub = t % CB_LANES;   // blockDim = 800 = 25 full warps
  float acc[CT_ATOMS];
#pragma unroll
  for (int u = 0; u < CT_ATOMS; ++u) acc[u] = 0.f;
  const int q1 = tab.o_ptr[slot + 1];
#pragma unroll 4
  for (int q = tab.o_ptr[slot] + sub; q < q1; q += CB_LANES) {
    const int pk = tab.o_pack[q];
    const float c = tab.o_coef[q];
    const int f = pk & 0xffff, o1 = (pk >> 16) & 0xff, o2 = (pk >> 24) & 0xff;
#pragma unroll
    for (int u = 0; u < CT_ATOMS; ++u) {
      float v = g1[u][f] * (m[u][o1] * m[u][o2]);
      if (second) v += g2[u][f] * (md[u][o1] * m[u][o2] + m[u][o1] * md[u][o2]);
      acc[u] += c * v;
    }
  }
#pragma unroll
  for (int u = 0; u < CT_ATOMS; ++u) {
#pragma unroll
    for (int o = CB_LANES / 2; o; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    if (sub == 0 && a0 + u < n) out[(a0 + u) * NPK + slot] = acc[u];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// small fp32 GEMM with strided operands:  C[i][j] = sum_k A(i,k) B(k,j)  (+ bias[j]), 32x32 tile, 2x2 per thread
// ---------------------------------------------------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_ACT = 1, EPI_TAN = 2, EPI_MULD = 3, EPI_SECOND = 4 };

__device__ __forceinline__ float sigm(float y) { return 1.0f / (1.0f + expf(-y)); }
__device__ __forceinline__ float swish(float y) { return kSwish * y * sigm(y); }                       // activation.py:7-9
__device__ __forceinline__ float dswish(float y) { const float s = sigm(y); return kSwish * (s + y * s * (1.f - s)); }
__device__ __forceinline__ float d2swish(float y) {
  const float s = sigm(y), q = s * (1.f - s);
  return kSwish * (2.f * q + y * q * (1.f - 2.f * s));
}

struct GemmT {
  const float* A; int64_t a_rs, a_cs;
  const float* B; int64_t b_rs, b_cs;
  float* C; int64_t ldc;
  int M, N, K;
  const float* bias;
  int epi;
  float* C2;          // second output, same shape/ld as C
  const float* X1;    // [M][ldc] pre-activations of the primal rows
  const float* X2;    // EPI_SECOND: adjoint of the tangent activations
  const float* X3;    // EPI_SECOND: tangent pre-activations
};

// Block tile 32 x 32, k-tile 32, 256 threads = 4 k-groups x 64 threads; a thread owns a 4 x 4 micro-tile and, inside every
// k-tile, the two 4-wide k-chunks of its group (intra-block split-K: these GEMMs have only ~72-96 output tiles, so the
// k-loop must be spread over the warps of an SM); the four partial tiles are added in a fixed order through shared
// memory.  Operands arrive by 16-byte cp.async in a 3-stage ring, in the orientation they have in global memory:
//   A_KC : A(i,k) contiguous along k (activations [rows][features])  -> As[i][k];  else contiguous along i -> As[k][i]
//   B_JC : B(k,j) contiguous along j (weights [in][out], adjoints)   -> Bs[k][j];  else contiguous along k -> Bs[j][k]
// Rows of a micro-tile are interleaved (ty + 8 r) where the operand is read along k, so that the float4 reads of a warp
// hit distinct banks (row stride 36 floats).
constexpr int TG_STAGES = 3, TG_LD = 36;

__device__ __forceinline__ void cp_async16_zfill(float* smem, const float* gmem, bool valid) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int bytes = valid ? 16 : 0;   // src-size 0: the destination is zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}

template <bool A_KC, bool B_JC>
__global__ void __launch_bounds__(256) k_tr_gemm(GemmT g) {
  __shared__ __align__(16) float As[TG_STAGES][32][TG_LD], Bs[TG_STAGES][32][TG_LD];
  __shared__ float red[4][32][33];
  const int t = threadIdx.x, grp = t >> 6, tt = t & 63, ty = tt >> 3, tx = tt & 7;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  // one 16-byte chunk of each operand per thread and k-tile
  const int c_hi = t >> 3, c_lo = (t & 7) * 4;   // chunk (row c_hi, elements c_lo .. c_lo+3) of a [32][32] tile
  auto load_tile = [&](int stage, int k0) {
    if (A_KC) {   // row = i, elements along k
      const int gi = i0 + c_hi, gk = k0 + c_lo;
      const bool ok = gi < g.M && gk < g.K;
      cp_async16_zfill(&As[stage][c_hi][c_lo], ok ? g.A + int64_t(gi) * g.a_rs + gk : g.A, ok);
    } else {      // row = k, elements along i
      const int gk = k0 + c_hi, gi = i0 + c_lo;
      const bool ok = gi < g.M && gk < g.K;
      cp_async16_zfill(&As[stage][c_hi][c_lo], ok ? g.A + int64_t(gk) * g.a_cs + gi : g.A, ok);
    }
    if (B_JC) {   // row = k, elements along j
      const int gk = k0 + c_hi, gj = j0 + c_lo;
      const bool ok = gj < g.N && gk < g.K;
      cp_async16_zfill(&Bs[stage][c_hi][c_lo], ok ? g.B + int64_t(gk) * g.b_rs + gj : g.B, ok);
    } else {      // row = j, elements along k
      const int gj = j0 + c_hi, gk = k0 + c_lo;
      const bool ok = gj < g.N && gk < g.K;
      cp_async16_zfill(&Bs[stage][c_hi][c_lo], ok ? g.B + int64_t(gj) * g.b_cs + gk : g.B, ok);
    }
  };
  const int nk = (g.K + 31) / 32;
#pragma unroll
  for (int st = 0; st < TG_STAGES - 1; ++st) {
    if (st < nk) load_tile(st, st * 32);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(TG_STAGES - 2) : "memory");
    __syncthreads();   // tile kt has landed for every thread, and everyone is done with tile kt-1
    if (kt + TG_STAGES - 1 < nk) load_tile((kt + TG_STAGES - 1) % TG_STAGES, (kt + TG_STAGES - 1) * 32);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int st = kt % TG_STAGES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = (grp + 4 * h) * 4;
      float a[4][4], b[4][4];   // a[r][kq], b[kq][c]
      if (A_KC) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(&As[st][ty + 8 * r][kk]);
          a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&As[st][kk + q][4 * ty]);
          a[0][q] = v.x; a[1][q] = v.y; a[2][q] = v.z; a[3][q] = v.w;
        }
      }
      if (B_JC) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[st][kk + q][4 * tx]);
          b[q][0] = v.x; b[q][1] = v.y; b[q][2] = v.z; b[q][3] = v.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[st][tx + 8 * c][kk]);
          b[0][c] = v.x; b[1][c] = v.y; b[2][c] = v.z; b[3][c] = v.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] += a[r][q] * b[q][c];
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) red[grp][A_KC ? ty + 8 * r : 4 * ty + r][B_JC ? 4 * tx + c : tx + 8 * c] = acc[r][c];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = t + 256 * q, il = e >> 5, jl = e & 31;
    const int i = i0 + il, j = j0 + jl;
    if (i >= g.M || j >= g.N) continue;
    float v = (red[0][il][jl] + red[1][il][jl]) + (red[2][il][jl] + red[3][il][jl]);
    if (g.bias) v += g.bias[j];
    const int64_t o = int64_t(i) * g.ldc + j;
    g.C[o] = v;
    switch (g.epi) {
      case EPI_ACT: g.C2[o] = swish(v); break;
      case EPI_TAN: g.C2[o] = dswish(g.X1[o]) * v; break;
      case EPI_MULD: g.C2[o] = dswish(g.X1[o]) * v; break;
      case EPI_SECOND: g.C2[o] = dswish(g.X1[o]) * v + g.X2[o] * d2swish(g.X1[o]) * g.X3[o]; break;
      default: break;
    }
  }
}

int gemm(cudaStream_t s, const GemmT& g) {
  dim3 grid((unsigned)cdiv(g.N, 32), (unsigned)cdiv(g.M, 32));
  const bool a_kc = g.a_cs == 1, b_jc = g.b_cs == 1;
  // 16-byte chunks: the contiguous extent of each operand must be a multiple of 4 elements
  if ((a_kc ? g.K : g.M) % 4 || (b_jc ? g.N : g.K) % 4 || (a_kc ? g.a_rs : g.a_cs) % 4 || (b_jc ? g.b_rs : g.b_cs) % 4)
    return APAX_ERR_INVALID_ARG;
  void (*const k_tr_gemm_nn)(GemmT) = k_tr_gemm<true, true>;     // activations x weights
  void (*const k_tr_gemm_nt)(GemmT) = k_tr_gemm<true, false>;    // adjoints x weights^T
  void (*const k_tr_gemm_tn)(GemmT) = k_tr_gemm<false, true>;    // activations^T x adjoints (weight gradients)
  if (a_kc && b_jc) { APAX_LAUNCH(k_tr_gemm_nn, grid, dim3(256), 0, s, g); }
  else if (a_kc && !b_jc) { APAX_LAUNCH(k_tr_gemm_nt, grid, dim3(256), 0, s, g); }
  else if (!a_kc && b_jc) { APAX_LAUNCH(k_tr_gemm_tn, grid, dim3(256), 0, s, g); }
  else return APAX_ERR_INVALID_ARG;
  return APAX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// head, energies, first adjoint row block
// ---------------------------------------------------------------------------------------------------------------------
// warp per atom: e_a = h2[a] . w2 + b2;  YB2[n + a][k] = -s~_a w2[k] swish'(y2[a][k])  (adjoint of the tangent pre-activation)
__global__ void __launch_bounds__(256) k_tr_head(TrWs w, const float* __restrict__ w2, const float* __restrict__ b2,
                                                 const int32_t* __restrict__ Z, const double* __restrict__ scale, int mask_atoms) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= w.n) return;
  const int z = Z[a];
  const float st = float((mask_atoms && z == 0) ? 0.0 : scale[z]);
  float acc = 0.f;
  for (int k = lane; k < NH; k += 32) {
    const int64_t o = int64_t(a) * NH + k;
    acc += w.H2[o] * w2[k];
    w.YB2[int64_t(w.n) * NH + o] = -st * w2[k] * dswish(w.Y2[o]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) w.e_atom[a] = acc + b2[0];
}

// block per structure: E_s = sum_a mask (scale[Z] e_a + shift[Z]) in fp64 (scaling.py:40-50, models.py:109-122)
__global__ void __launch_bounds__(128) k_tr_energy(TrWs w, const int32_t* __restrict__ struct_ptr, const int32_t* __restrict__ Z,
                                                   const double* __restrict__ scale, const double* __restrict__ shift,
                                                   int mask_atoms, double* __restrict__ energy_out) {
  __shared__ double red[128];
  const int s = blockIdx.x;
  double acc = 0.0;
  for (int a = struct_ptr[s] + threadIdx.x; a < struct_ptr[s + 1]; a += blockDim.x) {
    const int z = Z[a];
    if (!(mask_atoms && z == 0)) acc += scale[z] * double(w.e_atom[a]) + shift[z];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) { w.Es[s] = red[0]; energy_out[s] = red[0]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-pair reverse of the moments: q = d(-E)/d(dr) given Mdb = -dE/dM.  Thread per pair, block per atom.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_tr_pair_bwd(TrWs w, const float* __restrict__ emb, DescConstT dc) {
  __shared__ float A[NPK];
  const int a = blockIdx.x;
  for (int i = threadIdx.x; i < NPK; i += blockDim.x) A[i] = w.Mdb[int64_t(a) * NPK + i];
  __syncthreads();
  for (int s = w.rs[a] + threadIdx.x; s < w.re[a]; s += blockDim.x) {
    const int p = w.rowp[s];
    const float4 g = w.geo[p];
    PairT pt;
    pair_terms(g, dc, pt);
    float rho[NR], drho[NR], mono[20], gx[20], gy[20], gz[20];
    radial(emb, __float_as_int(g.w), dc, pt, rho, drho);
    monomials20(pt.n, mono);
    monomial_grads(pt.n, gx, gy, gz);
    float rbar = 0.f, nb0 = 0.f, nb1 = 0.f, nb2 = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float rb = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f;
#pragma unroll
      for (int k = 0; k < 20; ++k) {
        const float av = A[r * 20 + k];
        rb += av * mono[k];
        x0 += av * gx[k]; x1 += av * gy[k]; x2 += av * gz[k];
      }
      rbar += rb * drho[r];
      nb0 += rho[r] * x0; nb1 += rho[r] * x1; nb2 += rho[r] * x2;
    }
    // dn/dd^T nbar = nbar/(r+eps) - d (d.nbar) / (r (r+eps)^2);  dr/dd = d/r (0 at r = 0)
    float q0 = nb0 * pt.inv_re, q1 = nb1 * pt.inv_re, q2 = nb2 * pt.inv_re;
    if (pt.r > 0.f) {
      const float dn = g.x * nb0 + g.y * nb1 + g.z * nb2;
      const float k = rbar / pt.r - dn * pt.inv_re * pt.inv_re / pt.r;
      q0 += g.x * k; q1 += g.y * k; q2 += g.z * k;
    }
    w.Q[p] = make_float4(q0, q1, q2, 0.f);
  }
}

// block per structure: forces (fp64), u = dL/dF, alpha = dL/dE, loss of the structure (loss.py:231-243)
struct LossArgs {
  double wE, expE, wF, expF, batch_divisor;
};
__global__ void __launch_bounds__(128) k_tr_forces_loss(TrWs w, const int32_t* __restrict__ struct_ptr,
                                                        const int32_t* __restrict__ n_real, const double* __restrict__ E_lab,
                                                        const double* __restrict__ F_lab, LossArgs la, double* __restrict__ F_out,
                                                        double* __restrict__ loss) {
  __shared__ double red[128];
  const int s = blockIdx.x;
  const double na = double(n_real[s]);
  const double cE = la.wE / (pow(na, la.expE) * la.batch_divisor);
  const double cF = la.wF / (pow(na, la.expF) * la.batch_divisor);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double acc = 0.0;
  // warp per atom: lanes stride the atom's row (+) and transpose row (-), fixed butterfly sum in fp64
  for (int a = struct_ptr[s] + wid; a < struct_ptr[s + 1]; a += 4) {
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    const int r0 = w.rs[a], nr = w.re[a] - r0, t0 = w.ts[a], nt = w.te[a] - t0;
    for (int q = lane; q < nr + nt; q += 32) {
      const bool row = q < nr;
      const float4 v = w.Q[row ? w.rowp[r0 + q] : w.trowp[t0 + q - nr]];
      const double sg = row ? 1.0 : -1.0;
      f0 += sg * double(v.x); f1 += sg * double(v.y); f2 += sg * double(v.z);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      f0 += __shfl_xor_sync(0xffffffffu, f0, o);
      f1 += __shfl_xor_sync(0xffffffffu, f1, o);
      f2 += __shfl_xor_sync(0xffffffffu, f2, o);
    }
    if (lane == 0) {
      F_out[3 * int64_t(a) + 0] = f0; F_out[3 * int64_t(a) + 1] = f1; F_out[3 * int64_t(a) + 2] = f2;
      if (!F_lab) continue;   // energy + forces only (apax_train_energy_forces)
      const double e0 = f0 - F_lab[3 * int64_t(a) + 0], e1 = f1 - F_lab[3 * int64_t(a) + 1], e2 = f2 - F_lab[3 * int64_t(a) + 2];
      w.u[3 * int64_t(a) + 0] = 2.0 * cF * e0; w.u[3 * int64_t(a) + 1] = 2.0 * cF * e1; w.u[3 * int64_t(a) + 2] = 2.0 * cF * e2;
      acc += cF * (e0 * e0 + e1 * e1 + e2 * e2);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && E_lab) {
    const double dE = w.Es[s] - E_lab[s];
    w.alpha[s] = 2.0 * cE * dE;
    w.cF[s] = cF;
    w.loss_s[s] = red[0] + cE * dE * dE;
    // the last block to get here adds the per-structure losses up in index order (deterministic)
    __threadfence();
    if (atomicAdd(w.present + 128, 1) == int(gridDim.x) - 1 && loss) {
      __threadfence();
      const volatile double* ls = w.loss_s;
      double acc = 0.0;
      for (int64_t q = 0; q < w.B; ++q) acc += ls[q];
      loss[0] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// second half of the head: tangent output, first row block of the reverse sweep
// warp per atom: edot_a = hdot2[a] . w2;  YB2[a][k] = alpha s~ w2[k] swish'(y2) - s~ w2[k] swish''(y2) ydot2
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tr_head2(TrWs w, const float* __restrict__ w2, const int32_t* __restrict__ Z,
                                                  const int32_t* __restrict__ atom_struct, const double* __restrict__ scale,
                                                  int mask_atoms) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (a >= w.n) return;
  const int z = Z[a];
  const float st = float((mask_atoms && z == 0) ? 0.0 : scale[z]);
  const float al = float(w.alpha[atom_struct[a]]);
  float acc = 0.f;
  for (int k = lane; k < NH; k += 32) {
    const int64_t o = int64_t(a) * NH + k, od = int64_t(w.n) * NH + o;
    acc += w.H2[od] * w2[k];
    const float y = w.Y2[o];
    w.YB2[o] = st * w2[k] * (al * dswish(y) - d2swish(y) * w.Y2[od]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) w.e_atom[w.n + a] = acc;
}

// gradients of the output layer and of scale / shift, and (second kernel) the bias gradients.
// Column reductions over atoms use blocks of 8 columns x 32 atom groups with a fixed-order tree, so they are
// deterministic and spread over 32 blocks.
constexpr int CR_COLS = 8, CR_GROUPS = 32;

__device__ __forceinline__ float colreduce_finish(float acc, float (*red)[CR_COLS]) {
  const int c = threadIdx.x % CR_COLS, g = threadIdx.x / CR_COLS;
  red[g][c] = acc;
  __syncthreads();
  for (int o = CR_GROUPS / 2; o; o >>= 1) {
    if (g < o) red[g][c] += red[g + o][c];
    __syncthreads();
  }
  return red[0][c];
}

//   blocks [0, 32): w2bar[k] = sum_a s~_a (alpha h2[a][k] - hdot2[a][k]);  block 0 also b2bar = sum_a alpha s~_a
//   other blocks: warp per species z: scalebar[z] = sum_{a: Z_a = z} mask (alpha e_a - edot_a), shiftbar[z] = sum alpha mask
__global__ void __launch_bounds__(256) k_tr_head_grads(TrWs w, const int32_t* __restrict__ Z, const int32_t* __restrict__ atom_struct,
                                                       const double* __restrict__ scale, int mask_atoms, int S,
                                                       float* __restrict__ w2bar, float* __restrict__ b2bar,
                                                       double* __restrict__ scalebar, double* __restrict__ shiftbar) {
  __shared__ float red[CR_GROUPS][CR_COLS];
  if (blockIdx.x < NH / CR_COLS) {
    const int c = threadIdx.x % CR_COLS, g = threadIdx.x / CR_COLS, k = blockIdx.x * CR_COLS + c;
    float acc = 0.f, accb = 0.f;
    for (int64_t a = g; a < w.n; a += CR_GROUPS) {
      const int z = Z[a];
      const float st = float((mask_atoms && z == 0) ? 0.0 : scale[z]);
      const float al = float(w.alpha[atom_struct[a]]);
      acc += st * (al * w.H2[a * NH + k] - w.H2[(w.n + a) * NH + k]);
      accb += al * st;
    }
    const float tot = colreduce_finish(acc, red);
    if (g == 0) w2bar[k] = tot;
    if (blockIdx.x == 0) {
      __syncthreads();
      const float totb = colreduce_finish(accb, red);
      if (threadIdx.x == 0) b2bar[0] = totb;
    }
    return;
  }
  const int z = (blockIdx.x - NH / CR_COLS) * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (z >= S) return;
  double sc = 0.0, sh = 0.0;
  if (!(mask_atoms && z == 0)) {
    for (int64_t a = lane; a < w.n; a += 32) {
      if (Z[a] != z) continue;
      const double al = w.alpha[atom_struct[a]];
      sc += al * double(w.e_atom[a]) - double(w.e_atom[w.n + a]);
      sh += al;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
  }
  if (lane == 0) { scalebar[z] = sc; shiftbar[z] = sh; }
}

// bias gradients: column sums of the first n rows of the two pre-activation adjoints (blockIdx.y picks the layer)
__global__ void __launch_bounds__(256) k_tr_colsum(const float* __restrict__ X0, float* __restrict__ out0,
                                                   const float* __restrict__ X1, float* __restrict__ out1, int64_t n, int ld) {
  __shared__ float red[CR_GROUPS][CR_COLS];
  const float* X = blockIdx.y ? X1 : X0;
  float* out = blockIdx.y ? out1 : out0;
  const int c = threadIdx.x % CR_COLS, g = threadIdx.x / CR_COLS, k = blockIdx.x * CR_COLS + c;
  float acc = 0.f;
  for (int64_t a = g; a < n; a += CR_GROUPS) acc += X[a * ld + k];
  const float tot = colreduce_finish(acc, red);
  if (g == 0) out[k] = tot;
}

// ---------------------------------------------------------------------------------------------------------------------
// embedding gradient.  Per pair:  cbar[r][b] = rhobar_r fc G_b + rhodotbar_r rdot (fc G'_b + fc' G_b)
//   rhobar_r    = sum_k Mb2[r][k] mono_k + Mdb[r][k] monodot_k       rhodotbar_r = sum_k Mdb[r][k] mono_k
// then a fixed-order reduction per species pair into the [S][S][5][7] table.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_tr_emb_pair(TrWs w, const int32_t* __restrict__ idx, DescConstT dc) {
  __shared__ float A1[NPK], A2[NPK];
  const int a = blockIdx.x;
  for (int i = threadIdx.x; i < NPK; i += blockDim.x) {
    A1[i] = w.Mb2[int64_t(a) * NPK + i];
    A2[i] = w.Mdb[int64_t(a) * NPK + i];
  }
  __syncthreads();
  for (int s = w.rs[a] + threadIdx.x; s < w.re[a]; s += blockDim.x) {
    const int p = w.rowp[s];
    const float4 g = w.geo[p];
    PairT pt;
    pair_terms(g, dc, pt);
    float mono[20], gx[20], gy[20], gz[20], rdot, nd[3];
    monomials20(pt.n, mono);
    monomial_grads(pt.n, gx, gy, gz);
    pair_tangent(g, pt, w.u, a, idx[p], rdot, nd);
    float* out = w.cbar + int64_t(p) * (NR * dc.nb);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float rb = 0.f, rdb = 0.f;
#pragma unroll
      for (int k = 0; k < 20; ++k) {
        const float md = gx[k] * nd[0] + gy[k] * nd[1] + gz[k] * nd[2];
        rb += A1[r * 20 + k] * mono[k] + A2[r * 20 + k] * md;
        rdb += A2[r * 20 + k] * mono[k];
      }
#pragma unroll
      for (int b = 0; b < MAXNB; ++b)
        if (b < dc.nb)
          out[r * dc.nb + b] = dc.inv_sqrt_nb * (rb * pt.fc * pt.G[b] + rdb * rdot * (pt.fc * pt.dG[b] + pt.dfc * pt.G[b]));
    }
  }
}

// 256 blocks share the (present species)^2 table blocks; each sums the cbar rows whose key matches: thread-strided
// partial sums, then a fixed tree.
constexpr int ER_BLOCKS = 256;
__global__ void __launch_bounds__(128) k_tr_emb_reduce(TrWs w, int S, int ne, float* __restrict__ embbar) {
  constexpr int MAXE = NR * MAXNB;   // coefficients per species pair
  __shared__ float red[128][MAXE + 1];
  __shared__ int list[128];
  __shared__ int n_list;
  {   // ordered compaction of the present species: one flag per thread, ballot prefix per warp
    __shared__ int warp_cnt[4];
    const int z = threadIdx.x, lane = z & 31, wid = z >> 5;
    const bool on = z < S && w.present[z] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0;
    for (int q = 0; q < wid; ++q) base += warp_cnt[q];
    if (on) list[base + __popc(bal & ((1u << lane) - 1u))] = z;
    if (threadIdx.x == 0) n_list = warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
  }
  __syncthreads();
  const int np = n_list;
  for (int q = blockIdx.x; q < np * np; q += gridDim.x) {
    const int key = list[q / np] * S + list[q % np];
    float acc[MAXE];
#pragma unroll
    for (int i = 0; i < MAXE; ++i) acc[i] = 0.f;
    for (int64_t base = 0; base < w.P; base += 32 * int64_t(blockDim.x)) {
      unsigned match = 0;   // 32 independent key loads per thread, then only the matching rows are read
#pragma unroll 8
      for (int b = 0; b < 32; ++b) {
        const int64_t p = base + int64_t(b) * blockDim.x + threadIdx.x;
        if (p < w.P && __float_as_int(w.geo[p].w) == key) match |= 1u << b;
      }
      while (match) {
        const int b = __ffs(match) - 1;
        match &= match - 1;
        const float* c = w.cbar + (base + int64_t(b) * blockDim.x + threadIdx.x) * ne;
#pragma unroll
        for (int i = 0; i < MAXE; ++i)
          if (i < ne) acc[i] += c[i];
      }
    }
#pragma unroll
    for (int i = 0; i < MAXE; ++i) red[threadIdx.x][i] = acc[i];
    __syncthreads();
    for (int o = 64; o; o >>= 1) {
      if (threadIdx.x < o)
        for (int i = 0; i < ne; ++i) red[threadIdx.x][i] += red[threadIdx.x + o][i];
      __syncthreads();
    }
    if (threadIdx.x < ne) embbar[int64_t(key) * ne + threadIdx.x] = red[0][threadIdx.x];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// optimizer: optax.chain(clip(c), adam(linear_schedule), zero_nans()) per group (get_optimizer.py:73-85)
// ---------------------------------------------------------------------------------------------------------------------
struct AdamArgs {
  double lr_emb, lr_nn, lr_scale, lr_shift, clip, end_value, b1, b2, eps;
  int64_t transition_steps, transition_begin;
  int64_t emb_count;    // elements [0, emb_count) of the fp32 block are the embedding group
  int64_t n32, n64, S;
};

__device__ __forceinline__ double sched(double init, int64_t count, const AdamArgs& a) {
  // optax.linear_schedule == polynomial_schedule(power = 1)
  int64_t c = count - a.transition_begin;
  c = c < 0 ? 0 : (c > a.transition_steps ? a.transition_steps : c);
  const double frac = 1.0 - double(c) / double(a.transition_steps);
  return (init - a.end_value) * frac + a.end_value;
}

struct AdamScalars {
  double lr[4];      // scheduled learning rates: emb, nn, scale, shift (negative: group frozen)
  double bc1, bc2;   // 1 - b1^t, 1 - b2^t
};
static_assert(sizeof(AdamScalars) == 6 * sizeof(int64_t), "opt_state holds count + AdamScalars");

// opt_state: i64 [8] = { optax count, AdamScalars of the step being applied, spare }
__global__ void k_tr_adam_prep(AdamArgs a, int64_t* opt_state) {
  const int64_t count = opt_state[0];
  AdamScalars* sc = reinterpret_cast<AdamScalars*>(opt_state + 1);
  const double lr0[4] = {a.lr_emb, a.lr_nn, a.lr_scale, a.lr_shift};
  for (int i = 0; i < 4; ++i) sc->lr[i] = lr0[i] <= 1e-7 ? -1.0 : sched(lr0[i], count, a);   // set_to_zero (get_optimizer.py:74-75)
  const double t = double(count + 1);
  sc->bc1 = 1.0 - pow(a.b1, t);
  sc->bc2 = 1.0 - pow(a.b2, t);
  opt_state[0] = count + 1;
}

template <class T>
__device__ __forceinline__ void adam_one(T& p, T g, T& m, T& v, double lr, const AdamScalars& sc, const AdamArgs& a) {
  if (lr < 0.0) return;
  const T c = T(a.clip);
  g = g < -c ? -c : (g > c ? c : g);
  m = T(a.b1) * m + T(1.0 - a.b1) * g;
  v = T(a.b2) * v + T(1.0 - a.b2) * g * g;
  const T mh = m / T(sc.bc1);
  const T vh = v / T(sc.bc2);
  T upd = T(-lr) * (mh / (sqrt(vh) + T(a.eps)));
  if (upd != upd) upd = T(0);   // optax.zero_nans
  p += upd;
}

// four fp32 parameters per thread (the block length and the embedding/dense boundary are multiples of 64 elements)
__global__ void __launch_bounds__(256) k_tr_adam(AdamArgs a, float4* __restrict__ th32, const float4* __restrict__ g32,
                                                 float4* __restrict__ m32, float4* __restrict__ v32, double* __restrict__ th64,
                                                 const double* __restrict__ g64, double* __restrict__ m64,
                                                 double* __restrict__ v64, const int64_t* __restrict__ opt_state) {
  const AdamScalars sc = *reinterpret_cast<const AdamScalars*>(opt_state + 1);
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (4 * i < a.n32) {
    float4 p = th32[i], m = m32[i], v = v32[i];
    const float4 g = g32[i];
    const double lr = 4 * i < a.emb_count ? sc.lr[0] : sc.lr[1];
    adam_one<float>(p.x, g.x, m.x, v.x, lr, sc, a);
    adam_one<float>(p.y, g.y, m.y, v.y, lr, sc, a);
    adam_one<float>(p.z, g.z, m.z, v.z, lr, sc, a);
    adam_one<float>(p.w, g.w, m.w, v.w, lr, sc, a);
    th32[i] = p; m32[i] = m; v32[i] = v;
  }
  if (i < a.n64) {
    double p = th64[i], m = m64[i], v = v64[i];
    adam_one<double>(p, g64[i], m, v, i < a.S ? sc.lr[2] : sc.lr[3], sc, a);
    th64[i] = p; m64[i] = m; v64[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Data parallelism without a library collective: all-reduce + Adam in ONE kernel over NVLink peer memory.
// Every rank's gradient block lives in memory that its peers have mapped (CUDA IPC).  A rank (1) tells every peer that
// its block of this step is complete (release store of the step's epoch into the peer's flag array), (2) waits until all
// peers have said the same (acquire loads of its OWN flag array -- local memory), then (3) each thread loads its four
// parameters' gradients from every rank's block IN RANK ORDER (remote loads over NVLink for the peers, bypassing L1),
// adds them up and applies Adam to the local replica.  The sum order is the same on every rank, so replicas stay bitwise
// identical.  Gradient blocks are double-buffered by the step's parity: a rank can be at most one step ahead of a peer
// (it cannot pass the barrier of step t+1 before every peer has signalled t+1, i.e. finished reading step t), so the block
// being read is never the one being written.
// ---------------------------------------------------------------------------------------------------------------------
struct PeerArgs {
  const float4* g32[8];
  const double* g64[8];
  uint32_t* flags[8];   // flags[p]: rank p's flag array u32 [8] as mapped here; slot [r] is written by rank r
  int world, rank;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_tr_adam_peer(AdamArgs a, PeerArgs pa, float4* __restrict__ th32, float4* __restrict__ m32,
                                                      float4* __restrict__ v32, double* __restrict__ th64, double* __restrict__ m64,
                                                      double* __restrict__ v64, const int64_t* __restrict__ opt_state,
                                                      double* __restrict__ loss_out) {
  if (blockIdx.x == 0 && threadIdx.x < pa.world) {
    __threadfence_system();
    st_release_sys(pa.flags[threadIdx.x] + pa.rank, pa.epoch);
  }
  if (threadIdx.x < pa.world)   // bounded: a time-out sets the sticky error word behind the eight flags (apax_train_peers_status)
    peer_spin_until(pa.flags[pa.rank] + threadIdx.x, pa.epoch, pa.flags[pa.rank] + 8);
  __syncthreads();
  const AdamScalars sc = *reinterpret_cast<const AdamScalars*>(opt_state + 1);
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (4 * i < a.n32) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < pa.world; ++r) {
      const float4 t = __ldcg(pa.g32[r] + i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    float4 p = th32[i], m = m32[i], v = v32[i];
    const double lr = 4 * i < a.emb_count ? sc.lr[0] : sc.lr[1];
    adam_one<float>(p.x, g.x, m.x, v.x, lr, sc, a);
    adam_one<float>(p.y, g.y, m.y, v.y, lr, sc, a);
    adam_one<float>(p.z, g.z, m.z, v.z, lr, sc, a);
    adam_one<float>(p.w, g.w, m.w, v.w, lr, sc, a);
    th32[i] = p; m32[i] = m; v32[i] = v;
  }
  if (i <= a.n64) {   // i == n64: the loss rides behind the scale/shift gradients
    double g = 0.0;
    for (int r = 0; r < pa.world; ++r) g += __ldcg(pa.g64[r] + i);
    if (i == a.n64) {
      loss_out[0] = g;
    } else {
      double p = th64[i], m = m64[i], v = v64[i];
      adam_one<double>(p, g, m, v, i < a.S ? sc.lr[2] : sc.lr[3], sc, a);
      th64[i] = p; m64[i] = m; v64[i] = v;
    }
  }
}

}  // namespace
}  // namespace apax

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace apax;

extern "C" int apax_train_plan_create(const apax_gm_config* cfg, apax_train_plan** out) {
  if (!cfg || !out) return APAX_ERR_INVALID_ARG;
  if (cfg->n_radial != NR || cfg->n_basis < 1 || cfg->n_basis > MAXNB || cfg->basis_kind < 0 || cfg->basis_kind > 2 ||
      cfg->n_contr != 8 || cfg->n_hidden1 != NH || cfg->n_hidden2 != NH ||
      cfg->n_out != 1 || cfg->use_ntk != 0 || cfg->n_species < 1 || cfg->n_species > 128)
    return APAX_ERR_UNSUPPORTED;
  std::vector<Term> terms = build_terms();
  if (terms.empty()) return APAX_ERR_INVALID_ARG;
  const int nt = int(terms.size());
  std::vector<int32_t> f_ptr(NF + 1, 0), o_ptr(NPK + 1, 0);
  std::vector<float> f_coef(nt);
  std::vector<uchar4> f_slots(nt);
  for (int q = 0; q < nt; ++q) {
    f_ptr[terms[q].f + 1] += 1;
    f_coef[q] = terms[q].coef;
    f_slots[q] = make_uchar4(terms[q].a, terms[q].b, terms[q].c, 0);
  }
  for (int f = 0; f < NF; ++f) f_ptr[f + 1] += f_ptr[f];
  // occurrences by slot (the constant slot has none)
  std::vector<std::vector<std::pair<float, int32_t>>> occ(NPK);
  for (int q = 0; q < nt; ++q) {
    const int sl[3] = {terms[q].a, terms[q].b, terms[q].c};
    for (int w = 0; w < 3; ++w) {
      if (sl[w] == ONE_SLOT) continue;
      const int o1 = sl[(w + 1) % 3], o2 = sl[(w + 2) % 3];
      occ[sl[w]].push_back({terms[q].coef, int32_t(terms[q].f | (o1 << 16) | (o2 << 24))});
    }
  }
  std::vector<float> o_coef;
  std::vector<int32_t> o_pack;
  for (int i = 0; i < NPK; ++i) {
    o_ptr[i] = int32_t(o_coef.size());
    for (auto& e : occ[i]) { o_coef.push_back(e.first); o_pack.push_back(e.second); }
  }
  o_ptr[NPK] = int32_t(o_coef.size());
  const int no = int(o_coef.size());

  apax_train_plan* p = new apax_train_plan();
  p->cfg = *cfg;
  Carver c(nullptr);
  c.take<int32_t>(NF + 1); c.take<float>(nt); c.take<uchar4>(nt); c.take<int32_t>(NPK + 1); c.take<float>(no); c.take<int32_t>(no);
  const size_t bytes = c.bytes();
  if (cudaMalloc(&p->dev_block, bytes) != cudaSuccess) { delete p; return APAX_ERR_CUDA; }
  Carver d(p->dev_block);
  int32_t* d_fptr = d.take<int32_t>(NF + 1); float* d_fcoef = d.take<float>(nt); uchar4* d_fslots = d.take<uchar4>(nt);
  int32_t* d_optr = d.take<int32_t>(NPK + 1); float* d_ocoef = d.take<float>(no); int32_t* d_opack = d.take<int32_t>(no);
  bool ok = true;
  ok &= cudaMemcpy(d_fptr, f_ptr.data(), sizeof(int32_t) * (NF + 1), cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(d_fcoef, f_coef.data(), sizeof(float) * nt, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(d_fslots, f_slots.data(), sizeof(uchar4) * nt, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(d_optr, o_ptr.data(), sizeof(int32_t) * (NPK + 1), cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(d_ocoef, o_coef.data(), sizeof(float) * no, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(d_opack, o_pack.data(), sizeof(int32_t) * no, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { cudaFree(p->dev_block); delete p; return APAX_ERR_CUDA; }
  p->tab = Tables{d_fptr, d_fcoef, d_fslots, d_optr, d_ocoef, d_opack};
  {
    const int nb = cfg->n_basis;
    const double r_min = cfg->r_min, r_max = cfg->r_max;
    DescConstT& dc = p->dc;
    dc.r_max = cfg->r_max;
    dc.pi_f = float(M_PI);
    dc.nb = nb;
    dc.kind = cfg->basis_kind;
    dc.inv_sqrt_nb = cfg->use_embed_norm ? float(1.0 / sqrt(double(nb))) : 1.0f;
    for (int b = 0; b < MAXNB; ++b) dc.mu[b] = 0.f;
    if (cfg->basis_kind == BASIS_GAUSS_LINEAR) {          // basis_functions.py:22-28
      const double beta = double(nb) * nb / (r_max * r_max);
      dc.beta = float(beta); dc.two_beta = float(2.0 * beta); dc.norm = float(pow(2.0 * beta / M_PI, 0.25));
      for (int b = 0; b < nb; ++b) dc.mu[b] = float(r_min + (r_max - r_min) / nb * b);
    } else if (cfg->basis_kind == BASIS_GAUSS_EXP) {      // :29-37
      const double beta = pow(2.0 / nb * (exp(-r_min) - exp(-r_max)), -2.0);
      dc.beta = float(beta); dc.two_beta = float(2.0 * beta); dc.norm = 1.0f;
      for (int b = 0; b < nb; ++b)
        dc.mu[b] = float(nb > 1 ? exp(-r_max) + (exp(-r_min) - exp(-r_max)) * b / (nb - 1) : exp(-r_max));   // np.linspace
    } else {                                              // BesselBasis :72-78 (b_n is 1: the square root is of the PRODUCT)
      dc.beta = dc.two_beta = 0.f; dc.norm = 1.0f;
      for (int n = 0; n < nb; ++n) {
        const double a = ((n & 1) ? -1.0 : 1.0) * (sqrt(2.0) * M_PI / pow(r_max, 1.5));
        const double bb = double(n + 1) * (n + 2) / sqrt(double(n + 1) * (n + 1) * double(n + 2) * (n + 2));
        dc.mu[n] = float(float(a) * float(bb));
      }
    }
  }
  // parameter block layout: every array starts on a 64-element boundary
  const int64_t S = cfg->n_species;
  const int64_t sizes[7] = {S * S * NR * cfg->n_basis, int64_t(NF) * NH, NH, int64_t(NH) * NH, NH, NH, 1};
  int64_t off = 0;
  for (int i = 0; i < 7; ++i) { p->off32[i] = off; off = round_up(off + sizes[i], 64); }
  p->off32[7] = off;
  p->off64[0] = 0; p->off64[1] = S; p->off64[2] = 2 * S;
  *out = p;
  return APAX_OK;
}

extern "C" void apax_train_plan_destroy(apax_train_plan* p) {
  if (!p) return;
  cudaFree(p->dev_block);
  delete p;
}

extern "C" int apax_train_param_layout(const apax_train_plan* p, int64_t* offsets32, int64_t* offsets64) {
  if (!p) return APAX_ERR_INVALID_ARG;
  if (offsets32) for (int i = 0; i < 8; ++i) offsets32[i] = p->off32[i];
  if (offsets64) for (int i = 0; i < 3; ++i) offsets64[i] = p->off64[i];
  return APAX_OK;
}

extern "C" size_t apax_train_workspace_bytes(const apax_train_plan* p, int64_t n_atoms, int64_t n_pairs, int64_t n_structs) {
  const int nb = p ? p->cfg.n_basis : MAXNB;
  size_t b = carve_tr(nullptr, n_atoms, n_pairs, n_structs, nb, nullptr);
  return b + size_t(round_up(n_atoms, 64)) * sizeof(int32_t) + 256;   // + atom -> structure map
}

// grads == false: energy and forces only (E_label, F_label, lc, loss, grad32, grad64 unused)
static int train_run(bool grads, apax_stream_t stream_, const apax_train_plan* p, const float* theta32,
                     const double* theta64, const double* R, const int32_t* Z, const int32_t* idx,
                     const double* offsets, const int32_t* struct_ptr, const int32_t* pair_ptr,
                     const int32_t* n_real, int64_t n_structs, int64_t n_atoms, int64_t n_pairs,
                     const double* E_label, const double* F_label, const apax_train_loss_config* lc,
                     double batch_divisor, double* loss, double* energy, double* forces, float* grad32,
                     double* grad64, void* workspace, size_t workspace_bytes) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!p || !theta32 || !theta64 || !R || !Z || !idx || !struct_ptr || !pair_ptr || !energy || !forces || !workspace)
    return APAX_ERR_INVALID_ARG;
  if (grads && (!n_real || !E_label || !F_label || !lc || !loss || !grad32 || !grad64)) return APAX_ERR_INVALID_ARG;
  if (n_structs < 1 || n_atoms < 1 || n_pairs < 1 || n_atoms > (int64_t(1) << 24) || n_pairs > (int64_t(1) << 30)) return APAX_ERR_INVALID_ARG;
  if (workspace_bytes < apax_train_workspace_bytes(p, n_atoms, n_pairs, n_structs)) return APAX_ERR_WORKSPACE;
  TrWs w;
  const size_t used = carve_tr(workspace, n_atoms, n_pairs, n_structs, p->cfg.n_basis, &w);
  int32_t* atom_struct = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + used);
  const int64_t n = n_atoms, P = n_pairs, B = n_structs;
  const int S = p->cfg.n_species, mask = p->cfg.mask_atoms;
  const float* emb = theta32 + p->off32[0];
  const float *w0 = theta32 + p->off32[1], *b0 = theta32 + p->off32[2], *w1 = theta32 + p->off32[3], *b1 = theta32 + p->off32[4],
              *w2 = theta32 + p->off32[5], *b2 = theta32 + p->off32[6];
  const double *scale = theta64 + p->off64[0], *shift = theta64 + p->off64[1];
  float* const g0 = grads ? grad32 : nullptr;
  float *g_emb = g0 + p->off32[0], *g_w0 = g0 + p->off32[1], *g_b0 = g0 + p->off32[2], *g_w1 = g0 + p->off32[3],
        *g_b1 = g0 + p->off32[4], *g_w2 = g0 + p->off32[5], *g_b2 = g0 + p->off32[6];
  const Tables& tab = p->tab;
  const DescConstT& dc = p->dc;

  if (grads) APAX_CUDA_TRY(cudaMemsetAsync(grad32, 0, sizeof(float) * size_t(p->off32[7]), s));
  APAX_CUDA_TRY(cudaMemsetAsync(w.present, 0, sizeof(int32_t) * 132, s));
  // lists and geometry
  APAX_LAUNCH(k_tr_group, dim3((unsigned)B), dim3(128), 0, s, idx, P, struct_ptr, pair_ptr, Z, R, offsets, S, atom_struct, w);
  // ---- energy
  APAX_LAUNCH(k_tr_moments<0>, dim3((unsigned)n), dim3(128), 0, s, w, emb, idx, dc, w.M);
  APAX_LAUNCH(k_tr_contract<0>, dim3((unsigned)cdiv(n, CT_ATOMS)), dim3(384), 0, s, tab, w.M, nullptr, w.GS, n);
  GemmT g{};
  g = GemmT{w.GS, NF, 1, w0, NH, 1, w.Y1, NH, int(n), NH, NF, b0, EPI_ACT, w.H1, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  g = GemmT{w.H1, NH, 1, w1, NH, 1, w.Y2, NH, int(n), NH, NH, b1, EPI_ACT, w.H2, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  APAX_LAUNCH(k_tr_head, dim3((unsigned)cdiv(n * 32, 256)), dim3(256), 0, s, w, w2, b2, Z, scale, mask);
  APAX_LAUNCH(k_tr_energy, dim3((unsigned)B), dim3(128), 0, s, w, struct_ptr, Z, scale, shift, mask, energy);
  // ---- forces: adjoint chain of the TANGENT quantities (= minus the energy-gradient chain), rows [n, 2n)
  g = GemmT{w.YB2 + n * NH, NH, 1, w1, 1, NH, w.HB1 + n * NH, NH, int(n), NH, NH, nullptr, EPI_MULD, w.YB1 + n * NH, w.Y1, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  g = GemmT{w.YB1 + n * NH, NH, 1, w0, 1, NH, w.GB + n * NF, NF, int(n), NF, NH, nullptr, EPI_NONE, nullptr, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  APAX_LAUNCH(k_tr_contract_bwd, dim3((unsigned)cdiv(n, CT_ATOMS)), dim3(NPK * CB_LANES), 0, s, tab, w.M, nullptr, w.GB + n * NF, nullptr, w.Mdb, n);
  APAX_LAUNCH(k_tr_pair_bwd, dim3((unsigned)n), dim3(64), 0, s, w, emb, dc);
  if (!grads) {
    LossArgs la{0.0, 1.0, 0.0, 1.0, 1.0};
    APAX_LAUNCH(k_tr_forces_loss, dim3((unsigned)B), dim3(128), 0, s, w, struct_ptr, struct_ptr /* unused */, nullptr, nullptr, la,
                forces, loss);
    return APAX_OK;
  }
  LossArgs la{lc->energy_weight, lc->energy_atoms_exponent, lc->forces_weight, lc->forces_atoms_exponent, batch_divisor};
  APAX_LAUNCH(k_tr_forces_loss, dim3((unsigned)B), dim3(128), 0, s, w, struct_ptr, n_real, E_label, F_label, la, forces, loss);
  // ---- tangent sweep along u = dL/dF, rows [n, 2n)
  APAX_LAUNCH(k_tr_moments<1>, dim3((unsigned)n), dim3(128), 0, s, w, emb, idx, dc, w.Md);
  APAX_LAUNCH(k_tr_contract<1>, dim3((unsigned)cdiv(n, CT_ATOMS)), dim3(384), 0, s, tab, w.M, w.Md, w.GS + n * NF, n);
  g = GemmT{w.GS + n * NF, NF, 1, w0, NH, 1, w.Y1 + n * NH, NH, int(n), NH, NF, nullptr, EPI_TAN, w.H1 + n * NH, w.Y1, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  g = GemmT{w.H1 + n * NH, NH, 1, w1, NH, 1, w.Y2 + n * NH, NH, int(n), NH, NH, nullptr, EPI_TAN, w.H2 + n * NH, w.Y2, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  // ---- reverse sweep of  sum_s alpha_s E_s - Edot, rows [0, n)
  APAX_LAUNCH(k_tr_head2, dim3((unsigned)cdiv(n * 32, 256)), dim3(256), 0, s, w, w2, Z, atom_struct, scale, mask);
  APAX_LAUNCH(k_tr_head_grads, dim3((unsigned)(NH / CR_COLS + cdiv(S, 8))), dim3(256), 0, s, w, Z, atom_struct, scale, mask, S, g_w2, g_b2,
              grad64 + p->off64[0], grad64 + p->off64[1]);
  g = GemmT{w.YB2, NH, 1, w1, 1, NH, w.HB1, NH, int(n), NH, NH, nullptr, EPI_SECOND, w.YB1, w.Y1, w.HB1 + n * NH, w.Y1 + n * NH};
  APAX_TRY(gemm(s, g));
  g = GemmT{w.YB1, NH, 1, w0, 1, NH, w.GB, NF, int(n), NF, NH, nullptr, EPI_NONE, nullptr, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  // weight gradients: [x ; xdot]^T [ybar ; ydotbar] over all 2n rows
  g = GemmT{w.H1, 1, NH, w.YB2, NH, 1, g_w1, NH, NH, NH, int(2 * n), nullptr, EPI_NONE, nullptr, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  g = GemmT{w.GS, 1, NF, w.YB1, NH, 1, g_w0, NH, NF, NH, int(2 * n), nullptr, EPI_NONE, nullptr, nullptr, nullptr, nullptr};
  APAX_TRY(gemm(s, g));
  APAX_LAUNCH(k_tr_colsum, dim3(NH / CR_COLS, 2), dim3(256), 0, s, w.YB2, g_b1, w.YB1, g_b0, n, NH);
  // moments' adjoint incl. the second-order term, then the embedding gradient
  APAX_LAUNCH(k_tr_contract_bwd, dim3((unsigned)cdiv(n, CT_ATOMS)), dim3(NPK * CB_LANES), 0, s, tab, w.M, w.Md, w.GB, w.GB + n * NF, w.Mb2, n);
  APAX_LAUNCH(k_tr_emb_pair, dim3((unsigned)n), dim3(64), 0, s, w, idx, dc);
  APAX_LAUNCH(k_tr_emb_reduce, dim3(ER_BLOCKS), dim3(128), 0, s, w, S, NR * p->cfg.n_basis, g_emb);
  return APAX_OK;
}

extern "C" int apax_train_loss_and_grad(apax_stream_t stream, const apax_train_plan* p, const float* theta32,
                                        const double* theta64, const double* R, const int32_t* Z, const int32_t* idx,
                                        const double* offsets, const int32_t* struct_ptr, const int32_t* pair_ptr,
                                        const int32_t* n_real, int64_t n_structs, int64_t n_atoms, int64_t n_pairs,
                                        const double* E_label, const double* F_label, const apax_train_loss_config* lc,
                                        double batch_divisor, double* loss, double* energy, double* forces, float* grad32,
                                        double* grad64, void* workspace, size_t workspace_bytes) {
  return train_run(true, stream, p, theta32, theta64, R, Z, idx, offsets, struct_ptr, pair_ptr, n_real, n_structs, n_atoms, n_pairs,
                   E_label, F_label, lc, batch_divisor, loss, energy, forces, grad32, grad64, workspace, workspace_bytes);
}

extern "C" int apax_train_energy_forces(apax_stream_t stream, const apax_train_plan* p, const float* theta32,
                                        const double* theta64, const double* R, const int32_t* Z, const int32_t* idx,
                                        const double* offsets, const int32_t* struct_ptr, const int32_t* pair_ptr,
                                        int64_t n_structs, int64_t n_atoms, int64_t n_pairs, double* energy, double* forces,
                                        void* workspace, size_t workspace_bytes) {
  return train_run(false, stream, p, theta32, theta64, R, Z, idx, offsets, struct_ptr, pair_ptr, nullptr, n_structs, n_atoms, n_pairs,
                   nullptr, nullptr, nullptr, 1.0, nullptr, energy, forces, nullptr, nullptr, workspace, workspace_bytes);
}

extern "C" int apax_train_adam_step(apax_stream_t stream_, const apax_train_plan* p, const apax_train_opt_config* oc,
                                    float* theta32, double* theta64, const float* grad32, const double* grad64, float* m32,
                                    float* v32, double* m64, double* v64, int64_t* opt_state) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!p || !oc || !theta32 || !theta64 || !grad32 || !grad64 || !m32 || !v32 || !m64 || !v64 || !opt_state)
    return APAX_ERR_INVALID_ARG;
  if (oc->transition_steps < 1) return APAX_ERR_INVALID_ARG;
  AdamArgs a{oc->emb_lr, oc->nn_lr, oc->scale_lr, oc->shift_lr, oc->gradient_clipping, oc->end_value, oc->b1, oc->b2, oc->eps,
             oc->transition_steps, oc->transition_begin, p->off32[1], p->off32[7], p->off64[2], p->cfg.n_species};
  APAX_LAUNCH(k_tr_adam_prep, dim3(1), dim3(1), 0, s, a, opt_state);
  APAX_LAUNCH(k_tr_adam, dim3((unsigned)cdiv(std::max(a.n32 / 4, a.n64), 256)), dim3(256), 0, s, a,
              reinterpret_cast<float4*>(theta32), reinterpret_cast<const float4*>(grad32), reinterpret_cast<float4*>(m32),
              reinterpret_cast<float4*>(v32), theta64, grad64, m64, v64, opt_state);
  return APAX_OK;
}

// ---- peer-mapped memory (CUDA IPC) for the fused all-reduce + Adam
extern "C" int apax_peer_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_out) {
  if (bytes <= 0 || !dev_ptr || !ipc_handle_out) return APAX_ERR_INVALID_ARG;
  void* p = nullptr;
  APAX_CUDA_TRY(cudaMalloc(&p, size_t(bytes)));
  APAX_CUDA_TRY(cudaMemset(p, 0, size_t(bytes)));
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    g_last_error = cudaGetLastError();
    cudaFree(p);
    return APAX_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return APAX_OK;
}

extern "C" int apax_peer_open(const void* ipc_handle, void** dev_ptr) {
  if (!ipc_handle || !dev_ptr) return APAX_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  APAX_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return APAX_OK;
}

extern "C" int apax_peer_close(void* dev_ptr) {
  if (dev_ptr) APAX_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return APAX_OK;
}

extern "C" int apax_peer_free(void* dev_ptr) {
  if (dev_ptr) APAX_CUDA_TRY(cudaFree(dev_ptr));
  return APAX_OK;
}

extern "C" int apax_train_allreduce_adam_step(apax_stream_t stream_, const apax_train_plan* p, const apax_train_opt_config* oc,
                                              const apax_train_peers* peers, uint32_t epoch, float* theta32, double* theta64,
                                              float* m32, float* v32, double* m64, double* v64, int64_t* opt_state,
                                              double* loss_out) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!p || !oc || !peers || !theta32 || !theta64 || !m32 || !v32 || !m64 || !v64 || !opt_state || !loss_out)
    return APAX_ERR_INVALID_ARG;
  if (peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world || oc->transition_steps < 1)
    return APAX_ERR_INVALID_ARG;
  PeerArgs pa{};
  for (int r = 0; r < peers->world; ++r) {
    if (!peers->grad32[r] || !peers->grad64[r] || !peers->flags[r]) return APAX_ERR_INVALID_ARG;
    pa.g32[r] = reinterpret_cast<const float4*>(peers->grad32[r]);
    pa.g64[r] = peers->grad64[r];
    pa.flags[r] = peers->flags[r];
  }
  pa.world = peers->world; pa.rank = peers->rank; pa.epoch = epoch;
  AdamArgs a{oc->emb_lr, oc->nn_lr, oc->scale_lr, oc->shift_lr, oc->gradient_clipping, oc->end_value, oc->b1, oc->b2, oc->eps,
             oc->transition_steps, oc->transition_begin, p->off32[1], p->off32[7], p->off64[2], p->cfg.n_species};
  APAX_LAUNCH(k_tr_adam_prep, dim3(1), dim3(1), 0, s, a, opt_state);
  APAX_LAUNCH(k_tr_adam_peer, dim3((unsigned)cdiv(std::max(a.n32 / 4, a.n64 + 1), 256)), dim3(256), 0, s, a, pa,
              reinterpret_cast<float4*>(theta32), reinterpret_cast<float4*>(m32), reinterpret_cast<float4*>(v32), theta64, m64,
              v64, opt_state, loss_out);
  return APAX_OK;
}

extern "C" int apax_train_peers_status(const apax_train_peers* peers) {
  // synchronising query of this rank's sticky barrier-time-out word (0 = healthy)
  if (!peers || peers->rank < 0 || peers->rank >= 8 || !peers->flags[peers->rank]) return APAX_ERR_INVALID_ARG;
  uint32_t err = 0;
  APAX_CUDA_TRY(cudaMemcpy(&err, peers->flags[peers->rank] + 8, sizeof(err), cudaMemcpyDeviceToHost));
  return err == 0 ? APAX_OK : APAX_ERR_CUDA;
}
