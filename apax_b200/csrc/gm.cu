// GMNN energy+forces hot path on sm_100a: list preparation, fused pair kernels (moments forward,
// force backward), packed contractions, fp32 readout GEMMs, fp64 energy/force assembly.
//
// Reference semantics (apax, path:line): apax/layers/distances.py:48-67, apax/layers/descriptor/
// gaussian_moment_descriptor.py:31-103, basis_functions.py:19-158, moments.py:5-29,
// apax/layers/readout.py:22-50, ntk_linear.py:20-50, activation.py:7-9, scaling.py:40-50,
// apax/nn/models.py:87-171.  Nothing here is translated from the reference (which is pure JAX):
// the reference materialises [P,200] per-pair tensors and scatter-adds them; this file keeps the 100
// packed moment accumulators of an atom in registers, never materialises per-pair tensors, and
// assembles forces with a deterministic transpose gather instead of atomics.
#include <math.h>
#include <stdio.h>

#include "contract_gen.cuh"
#include "gm_internal.cuh"
#include "i8_gemm.cuh"
#include "tc_gemm.cuh"

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace apax {

// ================================================================================================
// workspace
// ================================================================================================
size_t carve_gm_workspace(void* base, int64_t n_atoms, int64_t n_pairs, int32_t n_out, int32_t n_species,
                          GmWorkspace* ws) {
  Carver c(base);
  GmWorkspace w{};
  const int64_t N = n_atoms, P = n_pairs > 0 ? n_pairs : 1;
  const int64_t ld = round_up(N > 0 ? N : 1, 128);
  w.n_atoms = N;
  w.n_central = N;
  w.n_pairs = n_pairs;
  w.ld = ld;
  w.n_out = n_out;
  w.rowptr = c.take<int32_t>(N + 1);
  w.trowptr = c.take<int32_t>(N + 1);
  w.nbr = c.take<int32_t>(P);
  w.src = c.take<int32_t>(P);
  w.tpos = c.take<int32_t>(P);
  w.tmp_a = c.take<int32_t>(P);
  w.tmp_b = c.take<int32_t>(P);
  w.cnt = c.take<int32_t>(N + 1);
  w.tcnt = c.take<int32_t>(N + 1);
  w.scan_scratch = c.take<int32_t>(scan_scratch_ints(N + 1));
  w.sp_present = c.take<int32_t>(128);
  w.sp_map = c.take<int32_t>(128);
  w.sp_z = c.take<int32_t>(128);
  w.n_sp = c.take<int32_t>(4);
  w.embc = c.take<float>(int64_t(n_species) * n_species * EMB_STRIDE_GEN);
  w.P = c.take<double4>(N + 1);
  w.G = c.take<float4>(P);
  w.D = c.take<float4>(P);
  w.MT = c.take<float>(NPK * ld);
  w.AT = c.take<float>(NPK * ld);
  w.GT = c.take<float>(NFP * ld);
  w.H1T = c.take<float>(NH * ld);
  w.D1T = c.take<float>(NH * ld);
  w.H2T = c.take<float>(NH * ld);
  w.D2T = c.take<float>(NH * ld);
  w.GLT = c.take<float>(NFP * ld);
  w.H1LT = c.take<float>(NH * ld);
  w.H2LT = c.take<float>(NH * ld);
  w.ebar = c.take<float>(ld);
  w.Eatom = c.take<double>(int64_t(n_out) * ld);
  w.partial = c.take<double>(int64_t(n_out) * 256);
  w.Frow = c.take<double>(N * 3 + 3);
  if (ws) *ws = w;
  return c.bytes();
}

// ================================================================================================
// list preparation: COO [2][P] -> CSR by central atom (rows sorted by original position) and the
// transpose map (for every atom, the CSR slots in which it is the neighbour).  Entries with
// idx[0]==idx[1] or an index outside [0,N) are dropped: they are exactly the (0,0)/(N,N) padding,
// which the reference masks to zero (apax/layers/masking.py:10-16, partition.py:656).
// ================================================================================================
namespace {

__global__ void __launch_bounds__(256) k_coo_count(const int32_t* __restrict__ idx, int64_t P, int N,
                                                   int32_t* __restrict__ cnt, int32_t* __restrict__ tcnt) {
  const int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (e >= P) return;
  const int i0 = idx[e], i1 = idx[P + e];
  if (i0 != i1 && i0 >= 0 && i0 < N && i1 >= 0 && i1 < N) {
    atomicAdd(&cnt[i1], 1);
    atomicAdd(&tcnt[i0], 1);
  }
}

__global__ void __launch_bounds__(256) k_coo_fill(const int32_t* __restrict__ idx, int64_t P, int N,
                                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                                  int32_t* __restrict__ nbr_tmp, int32_t* __restrict__ src_tmp) {
  const int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (e >= P) return;
  const int i0 = idx[e], i1 = idx[P + e];
  if (i0 != i1 && i0 >= 0 && i0 < N && i1 >= 0 && i1 < N) {
    const int slot = rowptr[i1] + atomicAdd(&cursor[i1], 1);
    nbr_tmp[slot] = i0;
    src_tmp[slot] = int(e);
  }
}

__global__ void __launch_bounds__(256) k_transpose_fill(const int32_t* __restrict__ rowptr, int N,
                                                        const int32_t* __restrict__ nbr,
                                                        const int32_t* __restrict__ trowptr,
                                                        int32_t* __restrict__ cursor, int32_t* __restrict__ tpos_tmp) {
  const int64_t s = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (s >= rowptr[N]) return;
  const int j = nbr[s];
  tpos_tmp[trowptr[j] + atomicAdd(&cursor[j], 1)] = int(s);
}

__global__ void __launch_bounds__(128) k_species_mark(const int32_t* __restrict__ Z, int64_t N, int S,
                                                      int32_t* __restrict__ present) {
  const int64_t i = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (i >= N) return;
  int z = Z[i];
  z = z < 0 ? 0 : (z >= S ? S - 1 : z);
  present[z] = 1;
}

// nb_gen = 0: the default table [35 -> 36] (7 Gaussians); nb_gen > 0: generic [5][16] rows of nb_gen coefficients, zero padded
__global__ void __launch_bounds__(256) k_species_build(const int32_t* __restrict__ present, int S,
                                                       const float* __restrict__ emb, int32_t* __restrict__ sp_map,
                                                       int32_t* __restrict__ n_sp_out, float* __restrict__ embc,
                                                       int32_t* __restrict__ sp_z, int nb_gen) {
  __shared__ int zs[128];
  __shared__ int n_sp;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int z = 0; z < S; ++z) {
      if (present[z]) {
        sp_map[z] = n;
        sp_z[n] = z;
        zs[n++] = z;
      } else {
        sp_map[z] = 0;
      }
    }
    n_sp = n;
    *n_sp_out = n;
  }
  __syncthreads();
  if (nb_gen > 0) {
    const int total = n_sp * n_sp * EMB_STRIDE_GEN;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int pair = i / EMB_STRIDE_GEN, k = i % EMB_STRIDE_GEN, rr = k / MAXNB_GEN, bb = k % MAXNB_GEN;
      const int a = pair / n_sp, b = pair % n_sp;
      embc[i] = bb < nb_gen ? emb[(int64_t(zs[a]) * S + zs[b]) * (NR * nb_gen) + rr * nb_gen + bb] : 0.0f;
    }
    return;
  }
  const int total = n_sp * n_sp * EMB_STRIDE;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int pair = i / EMB_STRIDE, k = i % EMB_STRIDE;
    const int a = pair / n_sp, b = pair % n_sp;
    embc[i] = k < NR * NB ? emb[(int64_t(zs[a]) * S + zs[b]) * (NR * NB) + k] : 0.0f;
  }
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_csr_import(const int32_t* __restrict__ rowptr_in, const int32_t* __restrict__ nbr_in,
                                                    int64_t N, int64_t P, int32_t* __restrict__ rowptr,
                                                    int32_t* __restrict__ nbr, int32_t* __restrict__ src,
                                                    int32_t* __restrict__ tcnt) {
  const int64_t s = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (s <= N) rowptr[s] = rowptr_in[s];
  if (s < P && s < rowptr_in[N]) {
    const int j = nbr_in[s];
    nbr[s] = j;
    src[s] = int(s);
    atomicAdd(&tcnt[j], 1);
  }
}
}  // namespace

// Transpose map of a SYMMETRIC list whose rows are sorted by neighbour index (what nl.cu's min-image builder
// produces): the reverse of slot s = (j in row a) is the slot of a in row j, found by binary search -- no
// atomics, no sort.  A slot without a partner (a pair exactly at the cutoff that rounding kept on one side only;
// its force contribution is exactly zero there) maps onto itself with zero weight via tpos = -1.
// The import of the builder's CSR (row pointers, neighbours, identity source map) is folded in: one pass over the list.
__global__ void __launch_bounds__(256) k_transpose_symmetric(const int32_t* __restrict__ rowptr, int64_t N,
                                                             const int32_t* __restrict__ nbr, int32_t* __restrict__ rowptr_out,
                                                             int32_t* __restrict__ nbr_out, int32_t* __restrict__ src_out,
                                                             int32_t* __restrict__ trowptr, int32_t* __restrict__ tpos,
                                                             int64_t n_central) {
  // warp per row a: lanes take the row's slots; each looks its own atom up in the (sorted) row of the neighbour
  const int64_t a = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a > N) return;
  if (lane == 0) { trowptr[a] = rowptr[a]; rowptr_out[a] = rowptr[a]; }
  if (a == N) return;
  const int rb = rowptr[a], re = rowptr[a + 1];
  for (int s = rb + lane; s < re; s += 32) {
    const int j = nbr[s];
    nbr_out[s] = j;
    src_out[s] = s;
    int b = __ldg(rowptr + j), e = __ldg(rowptr + j + 1) - 1, found = -1;
    // Interpolated start: neighbour indices of a row are roughly uniform over [0, N), so a sits near position cnt a / N.
    // One probe there and one at the far end of an 8-slot window narrow the search to the sector(s) around the guess --
    // the plain binary search spent 70 % of the kernel's stall samples on seven dependent probes scattered over the row's
    // three cache lines (6.5 GB of DRAM reads for a 0.36 GB list).  A wrong guess (indices not uniform) only costs the
    // two probes: the binary search below then runs on the remaining range.
    if (e - b >= 16) {
      const int g = b + int((int64_t(e - b + 1) * a) / N);
      const int gc = g > e ? e : g;
      const int v = __ldg(nbr + gc);
      if (v == int(a)) {
        found = gc;
        b = e + 1;
      } else if (v < int(a)) {
        b = gc + 1;
        const int w = gc + 8 < e ? gc + 8 : e;
        if (b <= e) { if (__ldg(nbr + w) >= int(a)) e = w; else b = w + 1; }
      } else {
        e = gc - 1;
        const int w = gc - 8 > b ? gc - 8 : b;
        if (b <= e) { if (__ldg(nbr + w) <= int(a)) b = w; else e = w - 1; }
      }
    }
    while (b <= e) {
      const int mid = (b + e) >> 1;
      const int v = __ldg(nbr + mid);
      if (v == int(a)) { found = mid; break; }
      if (v < int(a)) b = mid + 1; else e = mid - 1;
    }
    // rows of non-central atoms (ghosts of a domain-decomposed box) are never evaluated: their slots carry no contribution
    tpos[s] = j < n_central ? found : -1;
  }
}

// transpose map + compact species table; expects w.rowptr / w.nbr final and w.tcnt = counts by neighbour
static int gm_prepare_tail(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const int32_t* Z) {
  const int64_t N = w.n_atoms, P = w.n_pairs;
  const int S = p->cfg.n_species;
  APAX_TRY(exclusive_scan_i32(stream, w.tcnt, w.trowptr, N, w.scan_scratch));
  APAX_CUDA_TRY(cudaMemsetAsync(w.tcnt, 0, sizeof(int32_t) * (N + 1), stream));
  if (P > 0) {
    APAX_LAUNCH(k_transpose_fill, dim3((unsigned)cdiv(P, 256)), dim3(256), 0, stream, w.rowptr, int(N), w.nbr,
                w.trowptr, w.tcnt, w.tmp_a);
    APAX_TRY(rowsort_i32(stream, w.trowptr, N, w.tmp_a, nullptr, w.tpos, nullptr));
  }
  APAX_CUDA_TRY(cudaMemsetAsync(w.sp_present, 0, sizeof(int32_t) * 128, stream));
  APAX_LAUNCH(k_species_mark, dim3((unsigned)cdiv(N, 128)), dim3(128), 0, stream, Z, N, S, w.sp_present);
  APAX_LAUNCH(k_species_build, dim3(1), dim3(256), 0, stream, w.sp_present, S, p->emb, w.sp_map, w.n_sp, w.embc, w.sp_z, p->generic_basis ? p->cfg.n_basis : 0);
  return APAX_OK;
}

static int gm_prepare_impl(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w,
                           const int32_t* Z, const int32_t* idx) {
  const int64_t N = w.n_atoms, P = w.n_pairs;
  APAX_CUDA_TRY(cudaMemsetAsync(w.cnt, 0, sizeof(int32_t) * (N + 1), stream));
  APAX_CUDA_TRY(cudaMemsetAsync(w.tcnt, 0, sizeof(int32_t) * (N + 1), stream));
  if (P > 0) APAX_LAUNCH(k_coo_count, dim3((unsigned)cdiv(P, 256)), dim3(256), 0, stream, idx, P, int(N), w.cnt, w.tcnt);
  APAX_TRY(exclusive_scan_i32(stream, w.cnt, w.rowptr, N, w.scan_scratch));
  APAX_CUDA_TRY(cudaMemsetAsync(w.cnt, 0, sizeof(int32_t) * (N + 1), stream));
  if (P > 0) {
    APAX_LAUNCH(k_coo_fill, dim3((unsigned)cdiv(P, 256)), dim3(256), 0, stream, idx, P, int(N), w.rowptr, w.cnt,
                w.tmp_a, w.tmp_b);
    // rows ordered by original COO position: deterministic whatever order the atomics landed in
    APAX_TRY(rowsort_i32(stream, w.rowptr, N, w.tmp_b, w.tmp_a, w.src, w.nbr));
  }
  return gm_prepare_tail(stream, p, w, Z);
}

// list already in CSR form (rows = central atoms, e.g. straight from nl.cu): no COO pass, no row sort.
// symmetric_sorted: every pair appears in both directions and rows are sorted by neighbour (min-image builder).
int gm_prepare_from_csr(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const int32_t* Z,
                        const int32_t* rowptr_in, const int32_t* nbr_in, bool symmetric_sorted, int64_t n_central) {
  const int64_t N = w.n_atoms, P = w.n_pairs;
  if (!symmetric_sorted) {
    APAX_CUDA_TRY(cudaMemsetAsync(w.tcnt, 0, sizeof(int32_t) * (N + 1), stream));
    const int64_t n = (P > N + 1 ? P : N + 1);
    APAX_LAUNCH(k_csr_import, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, stream, rowptr_in, nbr_in, N, P, w.rowptr, w.nbr,
                w.src, w.tcnt);
    return gm_prepare_tail(stream, p, w, Z);
  }
  APAX_LAUNCH(k_transpose_symmetric, dim3((unsigned)cdiv((N + 1) * 32, 256)), dim3(256), 0, stream, rowptr_in, N, nbr_in,
              w.rowptr, w.nbr, w.src, w.trowptr, w.tpos, n_central < 0 ? N : n_central);
  const int S = p->cfg.n_species;
  APAX_CUDA_TRY(cudaMemsetAsync(w.sp_present, 0, sizeof(int32_t) * 128, stream));
  APAX_LAUNCH(k_species_mark, dim3((unsigned)cdiv(N, 128)), dim3(128), 0, stream, Z, N, S, w.sp_present);
  APAX_LAUNCH(k_species_build, dim3(1), dim3(256), 0, stream, w.sp_present, S, p->emb, w.sp_map, w.n_sp, w.embc, w.sp_z, p->generic_basis ? p->cfg.n_basis : 0);
  return APAX_OK;
}

// ================================================================================================
// pair kernels
// ================================================================================================
struct DescConst {
  float r_max;
  float beta;       // n_basis^2 / r_max^2            (basis_functions.py:23)
  float norm;       // (2 beta / pi)^(1/4)            (:24)
  float two_beta;
  float pi_f;       // float(np.pi)
  float mu[NB];     // r_min + (r_max-r_min)/n_basis * b (:25-27)
  // generic radial bases (pair kernels with GEN = 1): kind as apax_gm_config.basis_kind, nb <= 16 functions
  int kind, nb;
  float mu_gen[MAXNB_GEN];   // Gaussian centres (in r, or in exp(-r) for the exponential spacing); Bessel: prefactors a_n b_n
};

// Radial functions of the generic bases for one distance: rho_r = f_c(r) sum_b c[r][b] basis_b(r) and (GRAD) d rho_r / dr.
//   kind 0 / 1  GaussianBasis, linear / exponential spacing (basis_functions.py:19-56)
//   kind 2      BesselBasis (:59-79): a_n (sinc((n+1) r/rc) + sinc((n+2) r/rc)) -- the nb + 1 sines and cosines come from ONE
//               sincos(pi r / rc) and the Chebyshev recurrence sin((k+1)t) = 2 cos t sin(kt) - sin((k-1)t) (fp32 error
//               3.4e-8 of max|basis| against 2.7e-8 for the reference's direct evaluation)
// c: this species pair's [5][16] table (shared or global memory).  Returns false beyond the cutoff (everything zero).
template <bool GRAD>
__device__ __forceinline__ bool radial_gen(const DescConst& dc, const float* __restrict__ c, float r, float (&rho)[NR],
                                           float (&drho)[NR]) {
  if (!(r < dc.r_max)) return false;
  float G[MAXNB_GEN], dG[MAXNB_GEN];
  if (dc.kind == 2) {
    const float th = dc.pi_f * r / dc.r_max;
    float s1, c1;
    sincosf(th, &s1, &c1);
    const float two_c = 2.0f * c1;
    float s_prev = 0.0f, c_prev = 1.0f, s_cur = s1, c_cur = c1;   // k = 0, 1
    float sk_prev = 0.0f, dk_prev = 0.0f;
    const float inv_th = th != 0.0f ? 1.0f / th : 0.0f;           // one division for all nb + 1 sincs, one for their slopes
    const float inv_r = r > 0.0f ? 1.0f / r : 0.0f;
#pragma unroll
    for (int k = 1; k <= MAXNB_GEN + 1; ++k) {
      if (k <= dc.nb + 1) {
        const float sk = th != 0.0f ? s_cur * (inv_th * (1.0f / float(k))) : 1.0f;   // jnp.sinc(k r / rc) = sin(k t) / (k t)
        const float dk = (c_cur - sk) * inv_r;                         // d/dr sinc(k r / rc) = (cos(k t) - sinc) / r
        if (k >= 2) {
          G[k - 2] = dc.mu_gen[k - 2] * (sk_prev + sk);
          if (GRAD) dG[k - 2] = dc.mu_gen[k - 2] * (dk_prev + dk);
        }
        sk_prev = sk; dk_prev = dk;
        const float s_next = fmaf(two_c, s_cur, -s_prev), c_next = fmaf(two_c, c_cur, -c_prev);
        s_prev = s_cur; c_prev = c_cur; s_cur = s_next; c_cur = c_next;
      } else if (k >= 2) {
        G[k - 2] = 0.0f;
        if (GRAD) dG[k - 2] = 0.0f;
      }
    }
  } else {
    // exponential spacing: the Gaussians sit on exp(-r), and beta is large (33.6 for the default statics), so an error of
    // exp(-r) is amplified 2 beta |mu - x| ~ 40-fold in the exponent: take it correctly rounded (one fp64 exp per pair)
    const float x = dc.kind == 1 ? float(exp(-double(r))) : r;
    const float dxdr = dc.kind == 1 ? -x : 1.0f;
#pragma unroll
    for (int b = 0; b < MAXNB_GEN; ++b) {
      if (b < dc.nb) {
        const float dm = dc.mu_gen[b] - x;
        G[b] = dc.norm * expf(-dc.beta * (dm * dm));
        if (GRAD) dG[b] = G[b] * dc.two_beta * dm * dxdr;
      } else {
        G[b] = 0.0f;
        if (GRAD) dG[b] = 0.0f;
      }
    }
  }
  float sn, cs;
  sincosf(dc.pi_f * r / dc.r_max, &sn, &cs);
  const float fc = 0.5f * (cs + 1.0f);
  const float fcp = -0.5f * (dc.pi_f / dc.r_max) * sn;
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    float pr = 0.0f, qr = 0.0f;
#pragma unroll
    for (int q = 0; q < MAXNB_GEN / 4; ++q) {
      const float4 cv = *reinterpret_cast<const float4*>(c + rr * MAXNB_GEN + 4 * q);
      pr = fmaf(cv.x, G[4 * q], pr); pr = fmaf(cv.y, G[4 * q + 1], pr); pr = fmaf(cv.z, G[4 * q + 2], pr); pr = fmaf(cv.w, G[4 * q + 3], pr);
      if (GRAD) {
        qr = fmaf(cv.x, dG[4 * q], qr); qr = fmaf(cv.y, dG[4 * q + 1], qr); qr = fmaf(cv.z, dG[4 * q + 2], qr); qr = fmaf(cv.w, dG[4 * q + 3], qr);
      }
    }
    rho[rr] = pr * fc;
    if (GRAD) drho[rr] = fmaf(fcp, pr, fc * qr);
  }
  return true;
}

struct PairGeom {
  const double* R;
  const int32_t* Z;
  const double* offsets;  // [P][3] in COO order, or nullptr
  const double* box;      // [3][3] device, unused for GAS
  int mode;
  int n_species;
};

namespace {

__device__ __forceinline__ double wrap_unit(double x) {
  // jnp.mod(x + 0.5, 1.0) - 0.5  (space.py:146-157): fmod then +1 if negative
  double t = x + 0.5;
  double m = t - trunc(t);
  if (m < 0.0) m += 1.0;
  return m - 0.5;
}

__device__ __forceinline__ void pair_displacement(const PairGeom& g, const double (&box)[9], const double (&Ra)[3],
                                                  int j, int64_t coo_pos, float& dx, float& dy, float& dz) {
  if (g.mode == APAX_DISP_DRVEC) {  // caller supplies dr_vec itself, one row per COO entry
    dx = float(g.R[3 * coo_pos + 0]); dy = float(g.R[3 * coo_pos + 1]); dz = float(g.R[3 * coo_pos + 2]);
    return;
  }
  double v0 = Ra[0] - g.R[3 * int64_t(j) + 0];
  double v1 = Ra[1] - g.R[3 * int64_t(j) + 1];
  double v2 = Ra[2] - g.R[3 * int64_t(j) + 2];
  double d0, d1, d2;
  if (g.mode == APAX_DISP_GAS) {
    d0 = v0; d1 = v1; d2 = v2;
  } else {
    if (g.mode == APAX_DISP_MINIMAGE) {
      v0 = wrap_unit(v0); v1 = wrap_unit(v1); v2 = wrap_unit(v2);
    }
    d0 = box[0] * v0 + box[1] * v1 + box[2] * v2;
    d1 = box[3] * v0 + box[4] * v1 + box[5] * v2;
    d2 = box[6] * v0 + box[7] * v1 + box[8] * v2;
    if (g.offsets) {
      d0 += g.offsets[3 * coo_pos + 0];
      d1 += g.offsets[3 * coo_pos + 1];
      d2 += g.offsets[3 * coo_pos + 2];
    }
  }
  dx = float(d0); dy = float(d1); dz = float(d2);  // dr_vec.astype(float32) (gaussian_moment_descriptor.py:33)
}

__device__ __forceinline__ void load_coeffs(const float* emb, int pt, float (&c)[EMB_STRIDE]) {
  const float4* p = reinterpret_cast<const float4*>(emb + int64_t(pt) * EMB_STRIDE);
#pragma unroll
  for (int q = 0; q < EMB_STRIDE / 4; ++q) {
    const float4 v = p[q];
    c[4 * q + 0] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
  }
}

__device__ __forceinline__ void monomials(float x, float y, float z, float (&mono)[20]) {
  mono[0] = 1.0f; mono[1] = x; mono[2] = y; mono[3] = z;
  const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  mono[4] = xx; mono[5] = xy; mono[6] = xz; mono[7] = yy; mono[8] = yz; mono[9] = zz;
  mono[10] = xx * x; mono[11] = xx * y; mono[12] = xx * z; mono[13] = xy * y; mono[14] = xy * z;
  mono[15] = xz * z; mono[16] = yy * y; mono[17] = yy * z; mono[18] = yz * z; mono[19] = zz * z;
}

__device__ __forceinline__ const float* stage_species_table(const GmWorkspace& ws, float* s_emb, int stride = EMB_STRIDE) {
  const int n_sp = *ws.n_sp;
  const float* emb = ws.embc;
  if (n_sp <= MAX_SP_SMEM) {
    const int total = n_sp * n_sp * stride;
    for (int i = threadIdx.x; i < total; i += blockDim.x) s_emb[i] = ws.embc[i];
    emb = s_emb;
  }
  __syncthreads();
  return emb;
}

// ---- per-atom gather record: position + compact species id in one aligned 32-byte word
__global__ void __launch_bounds__(256) k_atom_pack(GmWorkspace ws, PairGeom geo) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= ws.n_atoms) return;
  int z = geo.Z[i];
  z = z < 0 ? 0 : (z >= geo.n_species ? geo.n_species - 1 : z);
  double4 p;
  if (geo.mode == APAX_DISP_DRVEC) {
    p.x = p.y = p.z = 0.0;
  } else {
    p.x = geo.R[3 * i + 0]; p.y = geo.R[3 * i + 1]; p.z = geo.R[3 * i + 2];
  }
  p.w = __longlong_as_double((long long)ws.sp_map[z]);
  ws.P[i] = p;
}

// ---- forward: packed moments of every atom.  LPA lanes cooperate on one atom; each lane walks a
// strided share of the atom's CSR row with the 100 accumulators in registers; an xor-butterfly over
// the LPA lanes finishes the segmented sum (no atomics, fixed order => bitwise reproducible).
// The neighbour gathers are software-pipelined: indices two pairs ahead, atom records one pair ahead.
// SPEC = 1: the launcher vouches for mode == APAX_DISP_MINIMAGE without offsets (the resident min-image list of the large
// periodic boxes), so the offset registers, the `src` walk and the mode branches fold away.
// `emb` is the staged species table: in shared memory when the launch wrapper says so (a pointer the compiler can follow
// back to the __shared__ array: LDS instead of generic loads), else the compact table in global memory.  `box` lives in
// shared memory, not registers: 18 registers less across the pair loop.
// F2 = 1: the 100 accumulations per pair are issued as 50 packed FFMA2 (fma.rn.f32x2, sm_100): the same roundings element
// by element (results are bitwise those of F2 = 0), half the issue slots -- a scalar three-register FFMA issues at 0.57 per
// clock and sub-partition on a B200 (tools/probes/probe_ffma2.cu: 42 Tflop/s), FFMA2 carries 1.56x the flops.
template <int LPA, int BT, int MINB, int SPEC, int F2, int GEN = 0>
__device__ __forceinline__ void pair_fwd_body(const GmWorkspace& ws, const PairGeom& geo, const DescConst& dc,
                                              const float* __restrict__ emb, const double* box, int n_sp,
                                              double* ra) {
  const int32_t* __restrict__ nbr = ws.nbr;
  const int32_t* __restrict__ src = ws.src;
  const double4* __restrict__ P = ws.P;
  const double* __restrict__ offs = SPEC ? nullptr : (geo.mode == APAX_DISP_DRVEC ? geo.R : geo.offsets);
  float4* __restrict__ G = ws.G;
  const int64_t gid = int64_t(blockIdx.x) * BT + threadIdx.x;
  const int64_t atom = gid / LPA;
  const int sub = int(gid % LPA);
  const bool valid = atom < ws.n_central;
  const bool wrap = SPEC ? true : geo.mode == APAX_DISP_MINIMAGE;
  const bool gas = SPEC ? false : geo.mode == APAX_DISP_GAS;
  int beg = 0, end = 0, zc = 0;
  double* Ra = ra + threadIdx.x;   // the central atom's position waits in shared memory (stride BT): 6 registers less
  Ra[0] = 0.0; Ra[BT] = 0.0; Ra[2 * BT] = 0.0;
  if (valid) {
    beg = ws.rowptr[atom];
    end = ws.rowptr[atom + 1];
    const double4 pa = P[atom];
    Ra[0] = pa.x; Ra[BT] = pa.y; Ra[2 * BT] = pa.z;
    zc = int(__double_as_longlong(pa.w));
  }
  float2 acc2[NPK / 2];              // (acc[2i], acc[2i+1]): adjacent monomials of one radial channel
  float* acc = reinterpret_cast<float*>(acc2);
#pragma unroll
  for (int k = 0; k < NPK / 2; ++k) acc2[k] = make_float2(0.0f, 0.0f);

  int e = beg + sub;
  const int safe = valid ? int(atom) : 0;
  int j_cur = e < end ? __ldg(nbr + e) : safe;
  int j_nxt = e + LPA < end ? __ldg(nbr + e + LPA) : safe;
  double4 p_cur = P[j_cur];
  double o_cur[3] = {0.0, 0.0, 0.0};
  if (offs && e < end) {
    const int64_t sp = __ldg(src + e);
    o_cur[0] = __ldg(offs + 3 * sp); o_cur[1] = __ldg(offs + 3 * sp + 1); o_cur[2] = __ldg(offs + 3 * sp + 2);
  }
  for (; e < end; e += LPA) {
    // ---- issue the gathers of the following pairs before touching this one
    const int j_nn = e + 2 * LPA < end ? __ldg(nbr + e + 2 * LPA) : safe;
    const double4 p_nxt = P[j_nxt];
    double o_nxt[3] = {0.0, 0.0, 0.0};
    if (offs && e + LPA < end) {
      const int64_t sp = __ldg(src + e + LPA);
      o_nxt[0] = __ldg(offs + 3 * sp); o_nxt[1] = __ldg(offs + 3 * sp + 1); o_nxt[2] = __ldg(offs + 3 * sp + 2);
    }
    // ---- displacement in fp64 (distances.py:48-67), cast once to fp32 (gaussian_moment_descriptor.py:33)
    double v0 = Ra[0] - p_cur.x, v1 = Ra[BT] - p_cur.y, v2 = Ra[2 * BT] - p_cur.z;
    if (wrap) { v0 = wrap_unit(v0); v1 = wrap_unit(v1); v2 = wrap_unit(v2); }
    double d0, d1, d2;
    if (gas) {
      d0 = v0; d1 = v1; d2 = v2;
    } else {
      d0 = box[0] * v0 + box[1] * v1 + box[2] * v2 + o_cur[0];
      d1 = box[3] * v0 + box[4] * v1 + box[5] * v2 + o_cur[1];
      d2 = box[6] * v0 + box[7] * v1 + box[8] * v2 + o_cur[2];
    }
    const float dx = float(d0), dy = float(d1), dz = float(d2);
    // embeddings[Z_central, Z_neighbour] (basis_functions.py:139-141).  Both words of the record's last lane are consumed (the
    // upper one is always 0): otherwise the compiler recycles that register for the index prefetch below while the 128-bit
    // gather that targets it is still in flight, and every iteration waits out the gather's latency on the write-after-write
    // (ncu: 28 % of the kernel's stall samples sat on that one SEL).
    const long long wbits = __double_as_longlong(p_cur.w);
    const int pt = zc * n_sp + (int(wbits) | int(wbits >> 32));
    __stcs(G + e, make_float4(dx, dy, dz, __int_as_float(pt)));   // streaming: read once (backward pass); the atom records own the L2
    p_cur = p_nxt; j_cur = j_nxt; j_nxt = j_nn;
    o_cur[0] = o_nxt[0]; o_cur[1] = o_nxt[1]; o_cur[2] = o_nxt[2];
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float r = r2 > 0.0f ? sqrtf(r2) : 0.0f;  // space.distance + safe_mask (space.py:314-323)
    if (r >= dc.r_max) continue;                   // cosine cutoff is exactly 0 there (basis_functions.py:82-85)
    const float s = 1.0f / (r + 1e-5f);            // dn = dr_vec / (dr + 1e-5) (gaussian_moment_descriptor.py:46-48)
    float rho_v[NR], unused_d[NR];
    if (GEN) {   // Bessel / exponential spacing / n_basis != 7: coefficients stay in the [5][16] table, basis by radial_gen
      radial_gen<false>(dc, emb + pt * EMB_STRIDE_GEN, r, rho_v, unused_d);
    } else {
      float c[EMB_STRIDE];
      load_coeffs(emb, pt, c);
      float gb[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float t = dc.mu[b] - r;
        gb[b] = dc.norm * expf(-dc.beta * (t * t));
      }
      const float fc = 0.5f * (cosf(dc.pi_f * r / dc.r_max) + 1.0f);
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        float pr = 0.0f;
#pragma unroll
        for (int b = 0; b < NB; ++b) pr = fmaf(c[rr * NB + b], gb[b], pr);
        rho_v[rr] = pr * fc;
      }
    }
    float mono[20];
    monomials(dx * s, dy * s, dz * s, mono);
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
      const float rho = rho_v[rr];
      if (F2) {
        const float2 rho2 = make_float2(rho, rho);
#pragma unroll
        for (int k = 0; k < 10; ++k)
          acc2[rr * 10 + k] = __ffma2_rn(rho2, make_float2(mono[2 * k], mono[2 * k + 1]), acc2[rr * 10 + k]);
      } else {
#pragma unroll
        for (int k = 0; k < 20; ++k) acc[rr * 20 + k] = fmaf(rho, mono[k], acc[rr * 20 + k]);
      }
    }
  }
#pragma unroll
  for (int o = LPA / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  if (valid) {
#pragma unroll
    for (int k = 0; k < NPK; ++k)
      if ((k % LPA) == sub) ws.MT[int64_t(k) * ws.ld + atom] = acc[k];
  }
}

// ---- backward: per-pair dE/d(dr_vec) from the packed moment adjoints of the central atom.  Pair
// terms are recomputed from the stored fp32 displacement; nothing per-pair other than (dx,dy,dz) was
// kept from the forward pass.  Also leaves the row sum (fp64) of the central atom in Frow.
template <int LPA, int BT, int MINB, int F2, int GEN = 0>
__device__ __forceinline__ void pair_bwd_body(const GmWorkspace& ws, const DescConst& dc, const float* __restrict__ emb) {
  const int64_t gid = int64_t(blockIdx.x) * BT + threadIdx.x;
  const int64_t atom = gid / LPA;
  const int sub = int(gid % LPA);
  const bool valid = atom < ws.n_central;
  int beg = 0, end = 0;
  float2 a2[NPK / 2];                // (a[2i], a[2i+1]) for the packed FFMA2 form
  float* a = reinterpret_cast<float*>(a2);
  if (valid) {
    beg = ws.rowptr[atom];
    end = ws.rowptr[atom + 1];
#pragma unroll
    for (int k = 0; k < NPK; ++k) a[k] = ws.AT[int64_t(k) * ws.ld + atom];
  } else {
#pragma unroll
    for (int k = 0; k < NPK; ++k) a[k] = 0.0f;
  }
  double sx = 0.0, sy = 0.0, sz = 0.0;
  const float4* __restrict__ Gp = ws.G;
  float4* __restrict__ Dp = ws.D;
  float4 g_nxt = beg + sub < end ? __ldg(Gp + beg + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = beg + sub; e < end; e += LPA) {
    const float4 g = g_nxt;
    if (e + LPA < end) g_nxt = __ldcs(Gp + e + LPA);  // next pair's geometry is in flight during this pair's math (streaming load)
    const float dx = g.x, dy = g.y, dz = g.z;
    const int pt = __float_as_int(g.w);
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float r = r2 > 0.0f ? sqrtf(r2) : 0.0f;
    float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (r < dc.r_max) {
      const float s = 1.0f / (r + 1e-5f);
      const float x = dx * s, y = dy * s, z = dz * s;
      float rho_v[NR], drho_v[NR];
      if (GEN) {
        radial_gen<true>(dc, emb + pt * EMB_STRIDE_GEN, r, rho_v, drho_v);
      } else {
        float c[EMB_STRIDE];
        load_coeffs(emb, pt, c);
        float gb[NB], gd[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float t = dc.mu[b] - r;
          gb[b] = dc.norm * expf(-dc.beta * (t * t));
          gd[b] = gb[b] * (dc.two_beta * t);  // dG_b/dr
        }
        float sn, cs;
        const float arg = dc.pi_f * r / dc.r_max;
        sincosf(arg, &sn, &cs);
        const float fc = 0.5f * (cs + 1.0f);
        const float fcp = -0.5f * (dc.pi_f / dc.r_max) * sn;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          float pr = 0.0f, qr = 0.0f;
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            pr = fmaf(c[rr * NB + b], gb[b], pr);
            qr = fmaf(c[rr * NB + b], gd[b], qr);
          }
          rho_v[rr] = pr * fc;
          drho_v[rr] = fmaf(fcp, pr, fc * qr);  // d rho_r / dr
        }
      }
      float mono[20];
      monomials(x, y, z, mono);
      float2 B2[10];
      float* B = reinterpret_cast<float*>(B2);
#pragma unroll
      for (int k = 0; k < 10; ++k) B2[k] = make_float2(0.0f, 0.0f);
      float rbar = 0.0f;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const float rho = rho_v[rr], drho = drho_v[rr];
        float rhobar = 0.0f;
        if (F2) {   // packed: even and odd monomials in the two halves, joined at the end (20 terms: well inside the fp32 noise budget)
          const float2 rho2 = make_float2(rho, rho);
          float2 rb2 = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            rb2 = __ffma2_rn(a2[rr * 10 + k], make_float2(mono[2 * k], mono[2 * k + 1]), rb2);
            B2[k] = __ffma2_rn(rho2, a2[rr * 10 + k], B2[k]);
          }
          rhobar = rb2.x + rb2.y;
        } else {
#pragma unroll
          for (int k = 0; k < 20; ++k) {
            rhobar = fmaf(a[rr * 20 + k], mono[k], rhobar);
            B[k] = fmaf(rho, a[rr * 20 + k], B[k]);
          }
        }
        rbar = fmaf(rhobar, drho, rbar);
      }
      // nbar_i = sum_k B_k d mono_k / d n_i
      const float xx = mono[4], xy = mono[5], xz = mono[6], yy = mono[7], yz = mono[8], zz = mono[9];
      float nx = B[1] + 2.0f * x * B[4] + y * B[5] + z * B[6] + 3.0f * xx * B[10] + 2.0f * xy * B[11] +
                 2.0f * xz * B[12] + yy * B[13] + yz * B[14] + zz * B[15];
      float ny = B[2] + x * B[5] + 2.0f * y * B[7] + z * B[8] + xx * B[11] + 2.0f * xy * B[13] + xz * B[14] +
                 3.0f * yy * B[16] + 2.0f * yz * B[17] + zz * B[18];
      float nz = B[3] + x * B[6] + y * B[8] + 2.0f * z * B[9] + xx * B[12] + xy * B[14] + 2.0f * xz * B[15] +
                 yy * B[17] + 2.0f * yz * B[18] + 3.0f * zz * B[19];
      // n = d*s, s = 1/(r+eps):  dbar = s*nbar + (d/r) * (rbar - s^2 (nbar.d));  dr/dd = 0 at r = 0
      const float sdot = nx * dx + ny * dy + nz * dz;
      const float radial = r > 0.0f ? (rbar - s * s * sdot) / r : 0.0f;
      out.x = fmaf(radial, dx, s * nx);
      out.y = fmaf(radial, dy, s * ny);
      out.z = fmaf(radial, dz, s * nz);
      sx += double(out.x); sy += double(out.y); sz += double(out.z);
    }
    __stcs(Dp + e, out);
  }
#pragma unroll
  for (int o = LPA / 2; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (valid && sub == 0) {
    ws.Frow[3 * atom + 0] = sx; ws.Frow[3 * atom + 1] = sy; ws.Frow[3 * atom + 2] = sz;
  }
}

template <int LPA, int BT, int MINB, int SPEC, int F2 = 0, int GEN = 0>
__global__ void __launch_bounds__(BT, MINB) k_pair_fwd(GmWorkspace ws, PairGeom geo, DescConst dc) {
  __shared__ __align__(16) float s_emb[MAX_SP_SMEM * MAX_SP_SMEM * (GEN ? EMB_STRIDE_GEN : EMB_STRIDE)];
  __shared__ double box[9];
  __shared__ double s_ra[3 * BT];
  const bool plain = SPEC ? false : (geo.mode == APAX_DISP_GAS || geo.mode == APAX_DISP_DRVEC);
  if (threadIdx.x < 9)
    box[threadIdx.x] = plain ? ((threadIdx.x % 4 == 0 && geo.mode == APAX_DISP_GAS) ? 1.0 : 0.0) : __ldg(geo.box + threadIdx.x);
  stage_species_table(ws, s_emb, GEN ? EMB_STRIDE_GEN : EMB_STRIDE);
  const int n_sp = *ws.n_sp;
  if (n_sp <= MAX_SP_SMEM) {
    pair_fwd_body<LPA, BT, MINB, SPEC, F2, GEN>(ws, geo, dc, s_emb, box, n_sp, s_ra);
  } else {
    pair_fwd_body<LPA, BT, MINB, SPEC, F2, GEN>(ws, geo, dc, ws.embc, box, n_sp, s_ra);
  }
}
template <int LPA, int BT, int MINB, int F2 = 0, int GEN = 0>
__global__ void __launch_bounds__(BT, MINB) k_pair_bwd(GmWorkspace ws, DescConst dc) {
  __shared__ __align__(16) float s_emb[MAX_SP_SMEM * MAX_SP_SMEM * (GEN ? EMB_STRIDE_GEN : EMB_STRIDE)];
  stage_species_table(ws, s_emb, GEN ? EMB_STRIDE_GEN : EMB_STRIDE);
  if (*ws.n_sp <= MAX_SP_SMEM) {
    pair_bwd_body<LPA, BT, MINB, F2, GEN>(ws, dc, s_emb);
  } else {
    pair_bwd_body<LPA, BT, MINB, F2, GEN>(ws, dc, ws.embc);
  }
}
// ---- force assembly: F_a = -( sum_{e: central a} dbar_e  -  sum_{e: neighbour a} dbar_e ), both sums
// in fp64 over fp32 per-pair values, i.e. the transpose of the two gathers R[idx[1]], R[idx[0]] in
// distances.py:55-56 (R is fp64 there, :54).  The second sum walks the transpose map.
template <int LPA>
__global__ void __launch_bounds__(256) k_force_gather(GmWorkspace ws, double* __restrict__ F) {
  const int64_t gid = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const int64_t atom = gid / LPA;
  const int sub = int(gid % LPA);
  const bool valid = atom < ws.n_atoms;
  int beg = 0, end = 0;
  if (valid) {
    beg = ws.trowptr[atom];
    end = ws.trowptr[atom + 1];
  }
  double tx = 0.0, ty = 0.0, tz = 0.0;
  for (int k = beg + sub; k < end; k += LPA) {
    const int t = ws.tpos[k];
    if (t < 0) continue;  // unmatched boundary pair of a symmetric list: zero contribution
    const float4 d = ws.D[t];
    tx += double(d.x); ty += double(d.y); tz += double(d.z);
    if (d.w != 0.0f) {   // empirical corrections: the pair's dE_corr/dr / r rides in .w, the vector is rebuilt in fp64
      const float4 g = ws.G[t];
      tx += double(d.w) * double(g.x); ty += double(d.w) * double(g.y); tz += double(d.w) * double(g.z);
    }
  }
#pragma unroll
  for (int o = LPA / 2; o > 0; o >>= 1) {
    tx += __shfl_xor_sync(0xffffffffu, tx, o);
    ty += __shfl_xor_sync(0xffffffffu, ty, o);
    tz += __shfl_xor_sync(0xffffffffu, tz, o);
  }
  if (valid && sub == 0) {
    const bool central = atom < ws.n_central;
    F[3 * atom + 0] = -((central ? ws.Frow[3 * atom + 0] : 0.0) - tx);
    F[3 * atom + 1] = -((central ? ws.Frow[3 * atom + 1] : 0.0) - ty);
    F[3 * atom + 2] = -((central ? ws.Frow[3 * atom + 2] : 0.0) - tz);
  }
}

// ================================================================================================
// contractions (thread per atom, generated straight-line code on packed moments)
// ================================================================================================
// The contraction programs contain block barriers (instruction-cache locality, see contract_gen.cuh), so
// out-of-range threads stay alive on a clamped atom index and only their stores are suppressed.
constexpr int kContractThreads = 128;   // two 128-thread blocks per SM (255 registers): 1.54 -> 1.50 ms (bwd), 0.47 -> 0.45 ms (fwd) against one of 256; 64 x 4 the same

// MODE 0: features -> GT.  1: exact (tf32 hi, residual lo) pairs for the floating-point tensor-core readout.
// 2: GT plus each atom's maximum of |g[k] c[k]| -- the row scale of the integer tensor-core readout's digit planes.
template <int MODE>
__global__ void __launch_bounds__(kContractThreads, 256 / kContractThreads) k_contract_fwd(GmWorkspace ws) {
  const int64_t atom_raw = int64_t(blockIdx.x) * kContractThreads + threadIdx.x;
  const bool live = atom_raw < ws.n_central;
  const int64_t atom = live ? atom_raw : ws.n_central - 1;
  float m[NPK];
#pragma unroll
  for (int k = 0; k < NPK; ++k) m[k] = ws.MT[int64_t(k) * ws.ld + atom];
  float* gt = ws.GT + atom;
  float* glt = ws.GLT + atom;
  const int64_t ld = ws.ld;
  float mx = 0.0f;
  gm_contract_forward(m, [&](int f, float v) {
    if (MODE == 2) mx = fmaxf(mx, fabsf(v * __ldg(ws.q_cs + f)));
    if (!live) return;
    if (MODE == 1) {
      float hi, lo;
      split_tf32(v, hi, lo);
      gt[int64_t(f) * ld] = hi;
      glt[int64_t(f) * ld] = lo;
    } else {
      gt[int64_t(f) * ld] = v;
    }
  });
  if (MODE == 2 && live) ws.q_rowmax[atom] = __float_as_int(mx < 3.0e38f ? mx : 3.4e38f);   // NaN / inf rows quantise to zero
}

// ---- features AND their digit planes in one kernel (the integer tensor-core readout's input): thread == atom in both
// phases.  Phase 1 runs the contraction program, parks the 360 features of the block's 128 atoms in a per-block slot of
// the (idle) feature buffer -- written and re-read by the same thread within microseconds, so the slot lives in L2 -- and
// keeps the row maximum of the scaled features.  Phase 2 re-reads its own column 64 features at a time, converts to
// fixed point at the row's exponent and emits the four base-256 digit planes through a 32 KB shared-memory tile as full
// 64-byte row segments (the layout of k_i8_quantize).  Against k_contract_fwd<2> + k_i8_quantize this drops one launch,
// the fp32 feature matrix' trip through HBM (1.44 KB/atom written, 1.5 KB/atom read) and the row-maximum array.
// Persistent blocks (grid <= 2 per SM), each owning the slot `blockIdx.x` of [NFP][128] floats.
__global__ void __launch_bounds__(kContractThreads, 256 / kContractThreads)
k_contract_fwd_quantize(GmWorkspace ws, const float* __restrict__ col_scale, int8_t* __restrict__ planes,
                        int32_t* __restrict__ row_exp, int64_t n_tiles) {
  static_assert(kContractThreads == 128, "one thread per atom of a 128-atom GEMM tile");
  __shared__ __align__(16) uint8_t tile[I8_DIGITS][128 * 64];
  __shared__ float s_cs[NFP];            // per-feature powers of two (0 for the padding features): broadcast shared-memory reads
  const int tid = threadIdx.x;
  for (int k = tid; k < NFP; k += kContractThreads) s_cs[k] = k < NF ? col_scale[k] : 0.0f;
  __syncthreads();
  float* slot = ws.GT + int64_t(blockIdx.x) * NFP * 128 + tid;     // this thread's column of the block's slot, row stride 128
  const int64_t plane_stride = ws.ld * int64_t(NFP);
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t m0 = t * 128, atom_raw = m0 + tid;
    const bool live = atom_raw < ws.n_central;
    const int64_t atom = live ? atom_raw : ws.n_central - 1;
    float m[NPK];
#pragma unroll
    for (int k = 0; k < NPK; ++k) m[k] = ws.MT[int64_t(k) * ws.ld + atom];
    float mx = 0.0f;
    gm_contract_forward(m, [&](int f, float v) {
      mx = fmaxf(mx, fabsf(v * s_cs[f]));
      slot[f * 128] = v;
    });
    if (!(mx < 3.0e38f)) mx = 3.4e38f;   // NaN / inf rows quantise to zero
    int e;
    float sc, sc2;
    i8_row_scale(live ? mx : 0.0f, I8_DIGITS, e, sc, sc2);          // rows beyond n_central (tile padding): zero digits
    row_exp[atom_raw] = e;
    // Phase 2 walks the 24 groups of 16 features with the NEXT group's slot values already in flight: ncu had 23 % of the
    // kernel's stall samples on the first use of the just-loaded values (L2 latency, 8 warps per SM to hide it).
    float nxt[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) nxt[j] = slot[j * 128];
#pragma unroll 1
    for (int grp = 0; grp < NFP / 16; ++grp) {
      const int k0 = grp * 16, g = grp & 3, kc = k0 & ~63;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = nxt[j];
      if (grp + 1 < NFP / 16) {
        const float* np_ = slot + (k0 + 16) * 128;
#pragma unroll
        for (int j = 0; j < 16; ++j) nxt[j] = (k0 + 16 + j < NF) ? np_[j * 128] : 0.0f;
      }
      uint32_t w[I8_DIGITS][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t ew[4], tb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          ew[j] = i8_digit_word<I8_DIGITS>(sc != 0.0f ? ((v[4 * q + j] * s_cs[k0 + 4 * q + j]) * sc) * sc2 : 0.0f);
        i8_transpose4(ew, tb);
#pragma unroll
        for (int d = 0; d < I8_DIGITS; ++d) w[d][q] = tb[I8_DIGITS - 1 - d];
      }
      const int piece = g ^ ((tid >> 1) & 3);   // 16-byte pieces of a row are xor-swizzled: conflict-free both ways
#pragma unroll
      for (int d = 0; d < I8_DIGITS; ++d)
        *reinterpret_cast<uint4*>(&tile[d][tid * 64 + piece * 16]) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
      if (g == 3) {                              // a 64-feature chunk of the tile is staged: write it out as 64-byte row segments
        __syncthreads();
#pragma unroll
        for (int d = 0; d < I8_DIGITS; ++d) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int idx = tid + 128 * j, r = idx >> 2, p = idx & 3;
            const uint4 val = *reinterpret_cast<const uint4*>(&tile[d][r * 64 + (p ^ ((r >> 1) & 3)) * 16]);
            __stcs(reinterpret_cast<uint4*>(planes + d * plane_stride + (m0 + r) * int64_t(NFP) + kc + p * 16), val);   // streaming: keeps the feature slot in L2
          }
        }
        __syncthreads();
      }
    }
  }
}

__global__ void __launch_bounds__(kContractThreads, 256 / kContractThreads) k_contract_bwd(GmWorkspace ws) {
  const int64_t atom_raw = int64_t(blockIdx.x) * kContractThreads + threadIdx.x;
  const bool live = atom_raw < ws.n_central;
  const int64_t atom = live ? atom_raw : ws.n_central - 1;
  float m[NPK], a[NPK];
#pragma unroll
  for (int k = 0; k < NPK; ++k) {
    m[k] = ws.MT[int64_t(k) * ws.ld + atom];
    a[k] = 0.0f;
  }
  const float* gt = ws.GT + atom;
  const int64_t ld = ws.ld;
  // looped form of the reverse program: one generic block per contraction type (~15 KB of code instead of ~200 KB of
  // straight-line code that ran at 12 % issue utilisation behind instruction-fetch stalls)
  __shared__ float s_ring[GM_RING_FLOATS_PER_THREAD * kContractThreads];   // feature adjoints in flight (cp.async)
  gm_contract_backward_looped<1>(m, a, gt, ld, s_ring + threadIdx.x, kContractThreads);
  if (live) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) ws.AT[int64_t(k) * ws.ld + atom] = a[k];
  }
}

// ---- small systems: one atom on kSplit warps.  Thread-per-atom leaves 12,288 atoms with one warp per scheduler on 96 SMs
// and the kernel's duration is then the latency of ONE warp walking the whole program (55 us backward, 41 us forward).
// The contraction blocks are independent, so here a CTA is 32 atoms x kSplit warps: warp w runs part w of the program
// (forward: whole blocks; backward: every kSplit-th (r, s) instance of each type) on the same 32 atoms.
constexpr int kSplit = GM_FWD_PARTS;
static_assert(kSplit == 4 && kContractThreads == 32 * kSplit, "split contraction kernels: 4 warps per 32-atom group");

// Features + digit planes, the split twin of k_contract_fwd_quantize: groups of 32 atoms (grid-stride), feature slot
// [NFP][32] per block; the row maximum is joined across the warps through shared memory and the 64-feature chunks of
// phase 2 are dealt to the warps (chunk kc/64 = w, w + 4).  n_groups covers the 128-row tile padding (zero digits).
__global__ void __launch_bounds__(kContractThreads, 256 / kContractThreads)
k_contract_fwd_quantize_split(GmWorkspace ws, const float* __restrict__ col_scale, int8_t* __restrict__ planes,
                              int32_t* __restrict__ row_exp, int64_t n_groups) {
  __shared__ __align__(16) uint8_t tile[kSplit][I8_DIGITS][32 * 64];
  __shared__ float s_mx[kSplit][32];
  __shared__ float s_cs[NFP];            // per-feature powers of two (0 for the padding features)
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < NFP; k += kContractThreads) s_cs[k] = k < NF ? col_scale[k] : 0.0f;
  __syncthreads();
  float* slot = ws.GT + int64_t(blockIdx.x) * NFP * 32 + lane;     // this atom's column of the block's slot, row stride 32
  const int64_t plane_stride = ws.ld * int64_t(NFP);
  for (int64_t t = blockIdx.x; t < n_groups; t += gridDim.x) {
    const int64_t m0 = t * 32, atom_raw = m0 + lane;
    const bool live = atom_raw < ws.n_central;
    const int64_t atom = live ? atom_raw : ws.n_central - 1;
    float m[NPK];
#pragma unroll
    for (int k = 0; k < NPK; ++k) m[k] = ws.MT[int64_t(k) * ws.ld + atom];
    float mx = 0.0f;
    gm_contract_forward_part(m, part, [&](int f, float v) {
      mx = fmaxf(mx, fabsf(v * s_cs[f]));
      slot[f * 32] = v;
    });
    s_mx[part][lane] = mx;
    __syncthreads();                       // every warp's features are in the slot, every partial maximum in s_mx
    mx = fmaxf(fmaxf(s_mx[0][lane], s_mx[1][lane]), fmaxf(s_mx[2][lane], s_mx[3][lane]));
    if (!(mx < 3.0e38f)) mx = 3.4e38f;   // NaN / inf rows quantise to zero
    int e;
    float sc, sc2;
    i8_row_scale(live ? mx : 0.0f, I8_DIGITS, e, sc, sc2);          // rows beyond n_central (tile padding): zero digits
    if (part == 0) row_exp[atom_raw] = e;
#pragma unroll 1
    for (int kc = part * 64; kc < NFP; kc += kSplit * 64) {
      float nxt[16];                       // the next group's slot values, in flight while the current group is converted
#pragma unroll
      for (int j = 0; j < 16; ++j) nxt[j] = (kc + j < NF) ? slot[(kc + j) * 32] : 0.0f;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int k0 = kc + g * 16;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = nxt[j];
        if (g < 3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) nxt[j] = (k0 + 16 + j < NF) ? slot[(k0 + 16 + j) * 32] : 0.0f;
        }
        uint32_t w[I8_DIGITS][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t ew[4], tb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ew[j] = i8_digit_word<I8_DIGITS>(sc != 0.0f ? ((v[4 * q + j] * s_cs[k0 + 4 * q + j]) * sc) * sc2 : 0.0f);
          i8_transpose4(ew, tb);
#pragma unroll
          for (int d = 0; d < I8_DIGITS; ++d) w[d][q] = tb[I8_DIGITS - 1 - d];
        }
        const int piece = g ^ ((lane >> 1) & 3);   // 16-byte pieces of a row are xor-swizzled: conflict-free both ways
#pragma unroll
        for (int d = 0; d < I8_DIGITS; ++d)
          *reinterpret_cast<uint4*>(&tile[part][d][lane * 64 + piece * 16]) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
      }
      __syncwarp();
#pragma unroll
      for (int d = 0; d < I8_DIGITS; ++d) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = lane + 32 * j, r = idx >> 2, p = idx & 3;
          const uint4 val = *reinterpret_cast<const uint4*>(&tile[part][d][r * 64 + (p ^ ((r >> 1) & 3)) * 16]);
          *reinterpret_cast<uint4*>(planes + d * plane_stride + (m0 + r) * int64_t(NFP) + kc + p * 16) = val;
        }
      }
      __syncwarp();
    }
    __syncthreads();                       // the slot and s_mx are rewritten by the next group
  }
}

// Reverse mode, split twin of k_contract_bwd: each warp accumulates the adjoint of its instances; the four partial
// adjoints are joined pairwise through shared memory (2 x 100 x 32 floats) in a fixed order.
__global__ void __launch_bounds__(kContractThreads, 256 / kContractThreads) k_contract_bwd_split(GmWorkspace ws) {
  __shared__ float s_ring[GM_RING_FLOATS_PER_THREAD * kContractThreads];   // feature adjoints in flight (cp.async)
  __shared__ float s_red[2][NPK][32];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int64_t atom_raw = int64_t(blockIdx.x) * 32 + lane;
  const bool live = atom_raw < ws.n_central;
  const int64_t atom = live ? atom_raw : ws.n_central - 1;
  float m[NPK], a[NPK];
#pragma unroll
  for (int k = 0; k < NPK; ++k) {
    m[k] = ws.MT[int64_t(k) * ws.ld + atom];
    a[k] = 0.0f;
  }
  gm_contract_backward_looped<kSplit>(m, a, ws.GT + atom, ws.ld, s_ring + threadIdx.x, kContractThreads, part);
  if (part >= 2) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) s_red[part - 2][k][lane] = a[k];
  }
  __syncthreads();
  if (part < 2) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) a[k] += s_red[part][k][lane];        // 0 += 2, 1 += 3
  }
  __syncthreads();
  if (part == 1) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) s_red[0][k][lane] = a[k];
  }
  __syncthreads();
  if (part == 0 && live) {
#pragma unroll
    for (int k = 0; k < NPK; ++k) ws.AT[int64_t(k) * ws.ld + atom] = a[k] + s_red[0][k][lane];
  }
}

// feature-major [360][ld] -> row-major [N][360] (only for the descriptor-only entry point)
__global__ void __launch_bounds__(256) k_features_to_rowmajor(const float* __restrict__ GT, int64_t ld, int64_t N,
                                                              float* __restrict__ g) {
  __shared__ float tile[32][33];
  const int64_t a0 = int64_t(blockIdx.x) * 32;
  const int f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r;
    const int64_t a = a0 + tx;
    tile[r][tx] = (f < NF && a < N) ? GT[int64_t(f) * ld + a] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t a = a0 + r;
    const int f = f0 + tx;
    if (a < N && f < NF) g[a * NF + f] = tile[tx][r];
  }
}

// row-major [N][360] -> feature-major [360][ld] (the feature adjoints handed to apax_gm_descriptor_vjp)
__global__ void __launch_bounds__(256) k_features_from_rowmajor(const float* __restrict__ g, int64_t ld, int64_t N,
                                                                float* __restrict__ GT) {
  __shared__ float tile[32][33];
  const int64_t a0 = int64_t(blockIdx.x) * 32;
  const int f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t a = a0 + r;
    const int f = f0 + tx;
    tile[r][tx] = (a < N && f < NF) ? g[a * NF + f] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r;
    const int64_t a = a0 + tx;
    if (f < NF && a < ld) GT[int64_t(f) * ld + a] = tile[tx][r];
  }
}

// per-slot dE/d(dr_vec) (fp32, CSR order) -> the caller's COO order in fp64; entries the preparation dropped (padding) stay 0
__global__ void __launch_bounds__(256) k_pair_grad_to_coo(const float4* __restrict__ D, const int32_t* __restrict__ src,
                                                          const int32_t* __restrict__ rowptr, int64_t n_central,
                                                          double* __restrict__ out) {
  const int64_t s = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (s >= rowptr[n_central]) return;
  const float4 d = D[s];
  const int64_t e = src[s];
  out[3 * e + 0] = double(d.x); out[3 * e + 1] = double(d.y); out[3 * e + 2] = double(d.z);
}

// ================================================================================================
// readout: fp32 SIMT GEMM on feature-major activations.
//   YT[n][m] = sum_k W[k][n] * X(k, m),  n in [0,NC), m in [0,ld)
// 128(n) x 128(m) x 8(k) tiles, 256 threads, 8x8 micro-tile, register-prefetch double buffering.
// ================================================================================================
constexpr float kSwish = 1.6765324703310907f;  // apax/layers/activation.py:8

enum { EPI_STORE = 0, EPI_ACT = 1, EPI_MULD = 2 };

struct GemmArgs {
  const float* W;     // [K][ldw]
  int ldw;
  const float* XT;    // [K][ld]
  float* YT;          // [NC][ld]   (EPI_ACT: swish(y))
  float* Y2T;         // EPI_ACT: swish'(y) written ; EPI_MULD: multiplier D[n][m] read
  const float* bias;  // EPI_ACT: [NC]
  const float* ebar;  // EPI_MULD: per-atom factor [ld]
  int K;              // multiple of 8
  int64_t ld;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kGemmStages = 4;  // 4 x 8 KB: global->shared copies run three k-tiles ahead of the FMAs

// 128(n) x 128(m) x 8(k) tiles, 256 threads, 8x8 register micro-tile.  Warps tile the block 4(n) x 2(m) and
// lanes tile a warp 4(n) x 8(m), so every fragment load is one shared-memory wavefront; operands arrive by
// cp.async (LDGSTS) through a 4-stage ring, i.e. no global-load latency on the FMA critical path.
template <int EPI>
__global__ void __launch_bounds__(256, 2) k_sgemm(GemmArgs g) {
  __shared__ __align__(16) float Ws[kGemmStages][8][128];
  __shared__ __align__(16) float Xs[kGemmStages][8][128];
  const int tid = threadIdx.x;
  const int64_t m0 = int64_t(blockIdx.x) * 128;
  const int n0 = blockIdx.y * 128;
  const int lrow = tid >> 5;         // 0..7
  const int lcol = (tid & 31) * 4;   // 0..124
  const int warp = tid >> 5, lane = tid & 31;
  const int nb = (warp >> 1) * 32 + (lane >> 3) * 4;  // first n of this thread (second group: +16)
  const int mb = (warp & 1) * 64 + (lane & 7) * 4;    // first m of this thread (second group: +32)

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  const float* wsrc = g.W + int64_t(lrow) * g.ldw + n0 + lcol;
  const float* xsrc = g.XT + int64_t(lrow) * g.ld + m0 + lcol;
  auto issue = [&](int stage, int kt) {
    cp_async16(&Ws[stage][lrow][lcol], wsrc + int64_t(kt) * 8 * g.ldw);
    cp_async16(&Xs[stage][lrow][lcol], xsrc + int64_t(kt) * 8 * g.ld);
  };
  const int nk = g.K / 8;
#pragma unroll
  for (int s = 0; s < kGemmStages - 1; ++s) {
    if (s < nk) issue(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kGemmStages - 2>();
    __syncthreads();  // tile kt has landed for everyone; everyone is done reading the stage refilled below
    {
      const int nxt = kt + kGemmStages - 1;
      if (nxt < nk) issue(nxt % kGemmStages, nxt);
      cp_async_commit();
    }
    const int cur = kt % kGemmStages;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const float4 wa = *reinterpret_cast<const float4*>(&Ws[cur][kk][nb]);
      const float4 wb = *reinterpret_cast<const float4*>(&Ws[cur][kk][nb + 16]);
      const float4 xa = *reinterpret_cast<const float4*>(&Xs[cur][kk][mb]);
      const float4 xb = *reinterpret_cast<const float4*>(&Xs[cur][kk][mb + 32]);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
    }
  }
  cp_async_wait<0>();

  float4 eb[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
  if (EPI == EPI_MULD) {
    eb[0] = *reinterpret_cast<const float4*>(g.ebar + m0 + mb);
    eb[1] = *reinterpret_cast<const float4*>(g.ebar + m0 + mb + 32);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + nb + (i < 4 ? i : 16 + (i - 4));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t m = m0 + mb + 32 * h;
      float4 v = make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      const int64_t o = int64_t(n) * g.ld + m;
      if (EPI == EPI_ACT) {
        const float b = g.bias[n];
        float y[4] = {v.x + b, v.y + b, v.z + b, v.w + b};
        float hh[4], dd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float sg = 1.0f / (1.0f + expf(-y[q]));
          hh[q] = kSwish * y[q] * sg;                        // variance_preserving_swish
          dd[q] = kSwish * sg * (1.0f + y[q] * (1.0f - sg)); // its derivative
        }
        *reinterpret_cast<float4*>(g.YT + o) = make_float4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<float4*>(g.Y2T + o) = make_float4(dd[0], dd[1], dd[2], dd[3]);
      } else if (EPI == EPI_MULD) {
        const float4 d = *reinterpret_cast<const float4*>(g.Y2T + o);
        *reinterpret_cast<float4*>(g.YT + o) =
            make_float4(v.x * d.x * eb[h].x, v.y * d.y * eb[h].y, v.z * d.z * eb[h].z, v.w * d.w * eb[h].w);
      } else {
        *reinterpret_cast<float4*>(g.YT + o) = v;
      }
    }
  }
}

// ---- output layer + PerElementScaleShift + mask_by_atom (readout.py:47-50, scaling.py:40-50,
// models.py:109-112).  Thread per atom; h = H2 . w2 + b2 in fp32, E_i in fp64.
template <int NOUT>
__global__ void __launch_bounds__(128) k_head(const float* __restrict__ H2T, const float* __restrict__ H2LT, int64_t ld, int64_t N,
                                              const float* __restrict__ w2, const float* __restrict__ b2, int n_out,
                                              const int32_t* __restrict__ Z, int S, const double* __restrict__ scale,
                                              const double* __restrict__ shift, int mask_atoms,
                                              double* __restrict__ Eatom, float* __restrict__ ebar) {
  const int64_t m = int64_t(blockIdx.x) * 128 + threadIdx.x;  // grid covers whole 128-atom tiles
  if (m >= ld) return;
  if (m >= N) {
    ebar[m] = 0.0f;
    for (int k = 0; k < n_out; ++k) Eatom[int64_t(k) * ld + m] = 0.0;
    return;
  }
  float acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; ++k) acc[k] = 0.0f;
  for (int n = 0; n < NH; ++n) {
    float h = H2T[int64_t(n) * ld + m];
    if (H2LT) h += H2LT[int64_t(n) * ld + m];  // hi + lo is the exact fp32 activation
#pragma unroll
    for (int k = 0; k < NOUT; ++k)
      if (k < n_out) acc[k] = fmaf(h, w2[n * n_out + k], acc[k]);
  }
  int z = Z[m];
  const double msk = (mask_atoms && z == 0) ? 0.0 : 1.0;
  z = z < 0 ? 0 : (z >= S ? S - 1 : z);
  const double sc = scale[z], sh = shift[z];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
    if (k < n_out) Eatom[int64_t(k) * ld + m] = (sc * double(acc[k] + b2[k]) + sh) * msk;
  ebar[m] = float(sc * msk);  // cotangent of h after the fp64->fp32 cast
}

// ---- the same tail when the output layer was already contracted in the GEMM epilogue (integer tensor-core
// readout): h[k][m] = sum_n swish(y2)[n][m] w2[n][k]
template <int NOUT>
__global__ void __launch_bounds__(128) k_head_scale_shift(const float* __restrict__ hT, int64_t ld, int64_t N,
                                                          const float* __restrict__ b2, int n_out,
                                                          const int32_t* __restrict__ Z, int S,
                                                          const double* __restrict__ scale, const double* __restrict__ shift,
                                                          int mask_atoms, double* __restrict__ Eatom, float* __restrict__ ebar) {
  const int64_t m = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (m >= ld) return;
  if (m >= N) {
    ebar[m] = 0.0f;
    for (int k = 0; k < n_out; ++k) Eatom[int64_t(k) * ld + m] = 0.0;
    return;
  }
  int z = Z[m];
  const double msk = (mask_atoms && z == 0) ? 0.0 : 1.0;
  z = z < 0 ? 0 : (z >= S ? S - 1 : z);
  const double sc = scale[z], sh = shift[z];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
    if (k < n_out) Eatom[int64_t(k) * ld + m] = (sc * double(hT[int64_t(k) * ld + m] + b2[k]) + sh) * msk;
  ebar[m] = float(sc * msk);
}

__device__ __forceinline__ int64_t cdiv_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }

__global__ void __launch_bounds__(256) k_negate_f64(double* __restrict__ x, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) x[i] = -x[i];
}

// ---- dE/dy2 = ebar * w2[:,head] * swish'(y2) as an exact (tf32 hi, lo) pair: input of the backward tensor-core GEMM
__global__ void __launch_bounds__(256) k_ybar2(const float* __restrict__ D2T, const float* __restrict__ ebar,
                                               const float* __restrict__ w2, int n_out, int head, int64_t ld,
                                               int64_t n_cols, float* __restrict__ out_hi, float* __restrict__ out_lo) {
  const int64_t m4 = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (m4 >= n_cols) return;
  const float w = w2[n * n_out + head];
  const float4 d = *reinterpret_cast<const float4*>(D2T + int64_t(n) * ld + m4);
  const float4 e = *reinterpret_cast<const float4*>(ebar + m4);
  const float v[4] = {e.x * w * d.x, e.y * w * d.y, e.z * w * d.z, e.w * w * d.w};
  float hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_tf32(v[q], hi[q], lo[q]);
  *reinterpret_cast<float4*>(out_hi + int64_t(n) * ld + m4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(out_lo + int64_t(n) * ld + m4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- fp64 energy sum, two fixed-shape stages => deterministic (apax/utils/math.py:7-12)
__global__ void __launch_bounds__(256) k_energy_partial(const double* __restrict__ Eatom, int64_t ld, int64_t N,
                                                        double* __restrict__ partial) {
  __shared__ double sm[256];
  const int head = blockIdx.y;
  const int64_t chunk = cdiv_dev(N, int64_t(gridDim.x));
  const int64_t b = int64_t(blockIdx.x) * chunk, e = min(N, b + chunk);
  double s = 0.0;
  for (int64_t i = b + threadIdx.x; i < e; i += 256) s += Eatom[int64_t(head) * ld + i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[head * 256 + blockIdx.x] = sm[0];
}

// per-structure energies of a batch: one warp per (structure, head), fixed summation order
__global__ void __launch_bounds__(32) k_energy_segments(const double* __restrict__ Eatom, int64_t ld,
                                                        const int32_t* __restrict__ seg_ptr, int n_out,
                                                        double* __restrict__ energy) {
  const int b = blockIdx.x, head = blockIdx.y;
  double s = 0.0;
  for (int i = seg_ptr[b] + threadIdx.x; i < seg_ptr[b + 1]; i += 32) s += Eatom[int64_t(head) * ld + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) energy[int64_t(b) * n_out + head] = s;
}

__global__ void k_energy_final(const double* __restrict__ partial, int n_blocks, double* __restrict__ energy) {
  const int head = blockIdx.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += partial[head * 256 + i];
    energy[head] = s;
  }
}

}  // namespace

// ================================================================================================
// host orchestration
// ================================================================================================
// ---- pairwise empirical corrections (apax/layers/empirical.py): ZBLRepulsion (:25-86) and ExponentialRepulsion (:89-126).
// E_corr = w * sum over list entries of E_ij(r) with w = 0.5 softplus(rep_scale) (ZBL) or 1 (exponential); both are
// evaluated in fp32 per entry and summed in fp64, like the reference (fp64_sum).  The distance is clipped to
// [0.02, r_max] (zero slope outside), the cosine cutoff ends at the correction's own r_max.
// The pair functions are evaluated in fp64 from the parameters' fp32 values: only the two or three entries per atom inside
// the corrections' range get here, so the cost is nil, and the steep repulsion (pair forces of 100 eV/A) would otherwise
// carry a few fp32 ulps of the LARGEST pair force into every force component (the reference's fp32 chain has the same
// weakness; the result stays inside its tolerance either way, this one is the closer to the exact derivative).
__device__ __forceinline__ void corr_pair(const CorrConst& cc, double r, int zi, int zj, double& e, double& dedr) {
  e = 0.0; dedr = 0.0;
  const double pi_d = 3.14159265358979323846;
  if (cc.zbl_on) {
    const double rmax = double(cc.zbl_rmax);
    const double dr = fmin(fmax(r, 0.02), rmax);
    const bool inside = r > 0.02 && r < rmax;
    double sn, cs;
    sincos(pi_d * dr / rmax, &sn, &cs);
    const double fc = 0.5 * (cs + 1.0), dfc = -0.5 * (pi_d / rmax) * sn;
    const double fi = double(zi), fj = double(zj);
    const double k = (pow(fi, 0.23) + pow(fj, 0.23)) / 0.46850;
    const double dist = dr * k;
    double f = 0.0, df = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double t = double(cc.zbl_c[q]) * exp(-double(cc.zbl_e[q]) * dist);
      f += t;
      df -= double(cc.zbl_e[q]) * k * t;
    }
    const double pref = fi * fj / dr;
    const double w = 0.5 * double(cc.zbl_scale);
    e += w * (pref * f * fc);
    if (inside) dedr += w * pref * (-f * fc / dr + df * fc + f * dfc);
  }
  if (cc.exp_on) {
    const double rmax = double(cc.exp_rmax);
    const double dr = fmin(fmax(r, 0.02), rmax);
    const bool inside = r > 0.02 && r < rmax;
    double sn, cs;
    sincos(pi_d * dr / rmax, &sn, &cs);
    const double fc = 0.5 * (cs + 1.0), dfc = -0.5 * (pi_d / rmax) * sn;
    const double ri = double(cc.exp_r[zi]), rj = double(cc.exp_r[zj]);
    const double k = (ri + rj) / (ri * rj);
    const double f = double(cc.exp_a[zi]) * double(cc.exp_a[zj]) * exp(-dr * k) / (dr * dr);
    e += f * fc;
    if (inside) dedr += f * (-k - 2.0 / dr) * fc + f * dfc;
  }
}

// warp per central atom.  FORCES = 0: Eatom[h][atom] += correction energy of the atom's row, for every head (before the
// energy sums).  FORCES = 1: D[slot].w = dE_corr/dr / r and Frow[atom] += the row sum of the vectors (after k_pair_bwd, before
// the gather).
template <int FORCES>
__global__ void __launch_bounds__(256) k_pair_corrections(GmWorkspace ws, PairGeom geo, CorrConst cc, int n_out) {
  const int64_t atom = (int64_t(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (atom >= ws.n_central) return;
  const int n_sp = *ws.n_sp;
  const float r_far = 1.0001f * fmaxf(cc.zbl_on ? cc.zbl_rmax : 0.f, cc.exp_on ? cc.exp_rmax : 0.f);
  const bool plain = geo.mode == APAX_DISP_GAS || geo.mode == APAX_DISP_DRVEC;
  double box[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) box[i] = plain ? 0.0 : __ldg(geo.box + i);
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
  for (int s = ws.rowptr[atom] + lane; s < ws.rowptr[atom + 1]; s += 32) {
    const float4 g = ws.G[s];
    const float r2 = g.x * g.x + g.y * g.y + g.z * g.z;
    if (!(r2 < r_far * r_far)) continue;
    // The reference takes the distance of these terms from the fp64 displacement (self.distance(dr_vec).astype(dtype),
    // empirical.py:56) -- the repulsion is steep enough for the fp32 rounding of the displacement to show -- so the few
    // entries in range get their displacement again in fp64 (distances.py:48-67, as in k_pair_fwd).
    double d0, d1, d2;
    const int64_t coo = ws.src[s];
    if (geo.mode == APAX_DISP_DRVEC) {
      d0 = geo.R[3 * coo]; d1 = geo.R[3 * coo + 1]; d2 = geo.R[3 * coo + 2];
    } else {
      const int64_t j = ws.nbr[s];
      double v0 = geo.R[3 * atom] - geo.R[3 * j], v1 = geo.R[3 * atom + 1] - geo.R[3 * j + 1], v2 = geo.R[3 * atom + 2] - geo.R[3 * j + 2];
      if (geo.mode == APAX_DISP_MINIMAGE) { v0 = wrap_unit(v0); v1 = wrap_unit(v1); v2 = wrap_unit(v2); }
      if (geo.mode == APAX_DISP_GAS) {
        d0 = v0; d1 = v1; d2 = v2;
      } else {
        d0 = box[0] * v0 + box[1] * v1 + box[2] * v2;
        d1 = box[3] * v0 + box[4] * v1 + box[5] * v2;
        d2 = box[6] * v0 + box[7] * v1 + box[8] * v2;
        if (geo.offsets) { d0 += geo.offsets[3 * coo]; d1 += geo.offsets[3 * coo + 1]; d2 += geo.offsets[3 * coo + 2]; }
      }
    }
    const double r64sq = d0 * d0 + d1 * d1 + d2 * d2;
    const double r = r64sq > 0.0 ? sqrt(r64sq) : 0.0;
    const int pt = __float_as_int(g.w);
    double e, dedr;
    corr_pair(cc, r, ws.sp_z[pt / n_sp], ws.sp_z[pt % n_sp], e, dedr);
    if (FORCES) {
      // the correction's share of dE/d(dr) is kept apart from the descriptor's fp32 value: its scalar dE_corr/dr / r goes
      // into the spare lane of D and the vector is formed in fp64 wherever it is summed (in the transpose gather and in the
      // virial from the stored fp32 displacement; here, for the central atom's own row, from the fp64 one)
      const double k = r > 0.0 ? dedr / r : 0.0;
      reinterpret_cast<float*>(ws.D + s)[3] = float(k);
      acc0 += k * d0; acc1 += k * d1; acc2 += k * d2;
    } else {
      acc0 += e;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
    if (FORCES) { acc1 += __shfl_xor_sync(0xffffffffu, acc1, o); acc2 += __shfl_xor_sync(0xffffffffu, acc2, o); }
  }
  if (lane != 0) return;
  if (FORCES) {
    ws.Frow[3 * atom + 0] += acc0; ws.Frow[3 * atom + 1] += acc1; ws.Frow[3 * atom + 2] += acc2;
  } else {
    for (int h = 0; h < n_out; ++h) ws.Eatom[int64_t(h) * ws.ld + atom] += acc0;
  }
}

// ---- stress x volume: dU/d(eps) = sum over list entries of dE/d(dr) (x) dr, for the strain applied as dr' = (I + eps) dr
// (apax/layers/properties.py:11-42 with periodic_general's perturbation, space.py:277-278).  Fixed-shape two-stage sum in
// fp64 over the fp32 per-pair values the force pass left in the workspace.
constexpr int kVirialBlocks = 1024;
__global__ void __launch_bounds__(256) k_virial_partial(const float4* __restrict__ D, const float4* __restrict__ G,
                                                        const int32_t* __restrict__ rowptr, int64_t n_central,
                                                        double* __restrict__ partial) {
  __shared__ double red[256][9];
  const int64_t P = rowptr[n_central];
  const int64_t per = (P + gridDim.x - 1) / gridDim.x;
  const int64_t b = int64_t(blockIdx.x) * per, e = b + per < P ? b + per : P;
  double acc[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) acc[q] = 0.0;
  for (int64_t s = b + threadIdx.x; s < e; s += blockDim.x) {
    const float4 d = D[s], g = G[s];
    const double dx = double(d.x) + double(d.w) * double(g.x);   // .w: empirical corrections (0 without them)
    const double dy = double(d.y) + double(d.w) * double(g.y);
    const double dz = double(d.z) + double(d.w) * double(g.z);
    acc[0] += dx * double(g.x); acc[1] += dx * double(g.y); acc[2] += dx * double(g.z);
    acc[3] += dy * double(g.x); acc[4] += dy * double(g.y); acc[5] += dy * double(g.z);
    acc[6] += dz * double(g.x); acc[7] += dz * double(g.y); acc[8] += dz * double(g.z);
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) red[threadIdx.x][q] = acc[q];
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o)
      for (int q = 0; q < 9; ++q) red[threadIdx.x][q] += red[threadIdx.x + o][q];
    __syncthreads();
  }
  if (threadIdx.x < 9) partial[blockIdx.x * 9 + threadIdx.x] = red[0][threadIdx.x];
}

// The OFFSETS branch as the reference computes it: disp_fn (apax/layers/distances.py:12-22) hands back dR + (I + eps) . dR and
// the offsets are added afterwards (:62-66), so dU/d(eps) = sum over entries of dE/d(dr') (x) dR with dR = box . (R[i1] - R[i0])
// in fp64 and dE/d(dr') taken at dr' = 2 dR + offsets (the force pass that ran just before, with the doubled box).
__global__ void __launch_bounds__(256) k_virial_partial_rows(GmWorkspace ws, const double* __restrict__ R,
                                                             const double* __restrict__ box, double* __restrict__ partial) {
  __shared__ double red[8][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double b[9], acc[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) { b[q] = __ldg(box + q); acc[q] = 0.0; }
  for (int64_t atom = int64_t(blockIdx.x) * 8 + warp; atom < ws.n_central; atom += int64_t(gridDim.x) * 8) {
    const double ax = R[3 * atom], ay = R[3 * atom + 1], az = R[3 * atom + 2];
    for (int s = ws.rowptr[atom] + lane; s < ws.rowptr[atom + 1]; s += 32) {
      const int64_t j = ws.nbr[s];
      const double v0 = ax - R[3 * j], v1 = ay - R[3 * j + 1], v2 = az - R[3 * j + 2];
      const double r0 = b[0] * v0 + b[1] * v1 + b[2] * v2;
      const double r1 = b[3] * v0 + b[4] * v1 + b[5] * v2;
      const double r2 = b[6] * v0 + b[7] * v1 + b[8] * v2;
      const float4 d = ws.D[s], g = ws.G[s];
      const double dx = double(d.x) + double(d.w) * double(g.x);   // .w: empirical corrections (0 without them)
      const double dy = double(d.y) + double(d.w) * double(g.y);
      const double dz = double(d.z) + double(d.w) * double(g.z);
      acc[0] += dx * r0; acc[1] += dx * r1; acc[2] += dx * r2;
      acc[3] += dy * r0; acc[4] += dy * r1; acc[5] += dy * r2;
      acc[6] += dz * r0; acc[7] += dz * r1; acc[8] += dz * r2;
    }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    if (lane == 0) red[warp][q] = acc[q];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
    partial[blockIdx.x * 9 + threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32) k_virial_final(const double* __restrict__ partial, int n_blocks, double* __restrict__ out) {
  if (threadIdx.x < 9) {
    double acc = 0.0;
    for (int b = 0; b < n_blocks; ++b) acc += partial[b * 9 + threadIdx.x];
    out[threadIdx.x] = acc;
  }
}

static DescConst make_desc_const(const apax_gm_config& cfg) {
  DescConst dc{};
  dc.r_max = cfg.r_max;
  const double beta = double(cfg.n_basis) * cfg.n_basis / (double(cfg.r_max) * cfg.r_max);
  dc.beta = float(beta);
  dc.two_beta = float(2.0 * beta);
  dc.norm = float(pow(2.0 * beta / M_PI, 0.25));
  dc.pi_f = float(M_PI);
  for (int b = 0; b < NB; ++b)
    dc.mu[b] = float(double(cfg.r_min) + (double(cfg.r_max) - double(cfg.r_min)) / cfg.n_basis * b);
  // generic bases (constants as apax_train_plan_create derives them for the table-driven kernels)
  const int nb = cfg.n_basis;
  const double r_min = cfg.r_min, r_max = cfg.r_max;
  dc.kind = cfg.basis_kind;
  dc.nb = nb;
  for (int b = 0; b < MAXNB_GEN; ++b) dc.mu_gen[b] = 0.0f;
  if (cfg.basis_kind == 0) {                                   // GaussianBasis, linear spacing (basis_functions.py:22-28)
    for (int b = 0; b < nb && b < MAXNB_GEN; ++b) dc.mu_gen[b] = float(r_min + (r_max - r_min) / nb * b);
  } else if (cfg.basis_kind == 1) {                            // exponential spacing (:29-37)
    const double be = pow(2.0 / nb * (exp(-r_min) - exp(-r_max)), -2.0);
    dc.beta = float(be);
    dc.two_beta = float(2.0 * be);
    dc.norm = 1.0f;
    for (int b = 0; b < nb && b < MAXNB_GEN; ++b)
      dc.mu_gen[b] = float(nb > 1 ? exp(-r_max) + (exp(-r_min) - exp(-r_max)) * b / (nb - 1) : exp(-r_max));   // np.linspace
  } else {                                                     // BesselBasis (:72-78; b_n is 1: the square root is of the PRODUCT)
    dc.norm = 1.0f;
    for (int n = 0; n < nb && n < MAXNB_GEN; ++n) {
      const double a = ((n & 1) ? -1.0 : 1.0) * (sqrt(2.0) * M_PI / pow(r_max, 1.5));
      const double bb = double(n + 1) * (n + 2) / sqrt(double(n + 1) * (n + 1) * double(n + 2) * (n + 2));
      dc.mu_gen[n] = float(float(a) * float(bb));
    }
  }
  return dc;
}

constexpr int kPairGeneric = 100;   // launch variant of the generic-basis pair kernels

static int pick_lpa(int64_t n_atoms) {
  if (n_atoms >= 100000) return 4;
  if (n_atoms >= 4096) return 8;
  if (n_atoms >= 1024) return 16;
  return 32;
}

// Block shape of the pair kernels.  Their ~200-225 live registers allow one 256-thread block per SM (8 warps); under a
// register cap the compiler parks a dozen accumulators in local memory (L1-resident) and more warps hide the latency of
// the serial per-pair prelude.  apax_gm_params_set_option(APAX_OPT_PAIR_BLOCK) picks the shape (a measurement switch, results
// are bitwise the same): 0 = 256 x 1 (uncapped), 1 = 128 x 3 (168 registers, default), 2 = 128 x 4 (128 registers).  Measured
// on the 1,000,002-atom box (fwd / bwd ms): 2.96 / 3.01, 2.21 / 2.52, 3.20 / 2.89; 64-thread blocks at 144 registers: 2.88 / 2.94.
// Round 2 (profiles/r02i_pair_ffma2.txt, r02d_pair_pipeline_ab.txt): packed FFMA2 accumulation 1.79 -> 1.73 ms forward
// (bitwise the same, now the default) but 2.33 -> 2.55 ms backward (more spills under the 168-register cap, long-scoreboard
// stalls 12 -> 26 % of the samples: variant 3 only); explicit software pipelining of the per-pair prelude against the
// accumulation of the previous pair: 2.86 / 3.12 ms (spills) -- dropped.  Variant 4 = the round-1 scalar kernels.
template <int LPA, int BT, int MINB, int F2 = 0>
static int launch_pair_fwd_shape(cudaStream_t s, const GmWorkspace& w, const PairGeom& geo, const DescConst& dc) {
  const dim3 grid((unsigned)cdiv(w.n_central * LPA, BT));
  if (geo.mode == APAX_DISP_MINIMAGE && !geo.offsets) {
    APAX_LAUNCH_AS("k_pair_fwd<LPA>", (k_pair_fwd<LPA, BT, MINB, 1, F2>), grid, dim3(BT), 0, s, w, geo, dc);
  } else {
    APAX_LAUNCH_AS("k_pair_fwd<LPA>", (k_pair_fwd<LPA, BT, MINB, 0, F2>), grid, dim3(BT), 0, s, w, geo, dc);
  }
  return APAX_OK;
}
template <int LPA>
static int launch_pair_fwd(cudaStream_t s, const GmWorkspace& w, const PairGeom& geo, const DescConst& dc, int variant) {
  if (variant == kPairGeneric) {   // Bessel / exponential-spacing / n_basis != 7 models
    const dim3 grid((unsigned)cdiv(w.n_central * LPA, 128));
    if (geo.mode == APAX_DISP_MINIMAGE && !geo.offsets) {
      APAX_LAUNCH_AS("k_pair_fwd_gen<LPA>", (k_pair_fwd<LPA, 128, 3, 1, 1, 1>), grid, dim3(128), 0, s, w, geo, dc);
    } else {
      APAX_LAUNCH_AS("k_pair_fwd_gen<LPA>", (k_pair_fwd<LPA, 128, 3, 0, 1, 1>), grid, dim3(128), 0, s, w, geo, dc);
    }
    return APAX_OK;
  }
  switch (variant) {
    case 0: return launch_pair_fwd_shape<LPA, 256, 1>(s, w, geo, dc);
    case 2: return launch_pair_fwd_shape<LPA, 128, 4>(s, w, geo, dc);
    case 4: return launch_pair_fwd_shape<LPA, 128, 3, 0>(s, w, geo, dc);
    default: return launch_pair_fwd_shape<LPA, 128, 3, 1>(s, w, geo, dc);   // 1 (default) and 3: packed FFMA2 accumulation
  }
}
template <int LPA>
static int launch_pair_bwd(cudaStream_t s, const GmWorkspace& w, const DescConst& dc, int variant) {
  if (variant == kPairGeneric) {
    APAX_LAUNCH_AS("k_pair_bwd_gen<LPA>", (k_pair_bwd<LPA, 128, 3, 0, 1>), dim3((unsigned)cdiv(w.n_central * LPA, 128)), dim3(128), 0, s, w, dc);
    return APAX_OK;
  }
  switch (variant) {
    case 0: APAX_LAUNCH_AS("k_pair_bwd<LPA>", (k_pair_bwd<LPA, 256, 1>), dim3((unsigned)cdiv(w.n_central * LPA, 256)), dim3(256), 0, s, w, dc); break;
    case 2: APAX_LAUNCH_AS("k_pair_bwd<LPA>", (k_pair_bwd<LPA, 128, 4>), dim3((unsigned)cdiv(w.n_central * LPA, 128)), dim3(128), 0, s, w, dc); break;
    case 3: APAX_LAUNCH_AS("k_pair_bwd<LPA>", (k_pair_bwd<LPA, 128, 3, 1>), dim3((unsigned)cdiv(w.n_central * LPA, 128)), dim3(128), 0, s, w, dc); break;
    default: APAX_LAUNCH_AS("k_pair_bwd<LPA>", (k_pair_bwd<LPA, 128, 3>), dim3((unsigned)cdiv(w.n_central * LPA, 128)), dim3(128), 0, s, w, dc); break;
  }
  return APAX_OK;
}
template <int LPA>
static int launch_force_gather(cudaStream_t s, const GmWorkspace& w, double* F) {
  APAX_LAUNCH(k_force_gather<LPA>, dim3((unsigned)cdiv(w.n_atoms * LPA, 256)), dim3(256), 0, s, w, F);
  return APAX_OK;
}

#define APAX_DISPATCH_LPA(lpa, fn, ...)            \
  do {                                             \
    switch (lpa) {                                 \
      case 4: APAX_TRY(fn<4>(__VA_ARGS__)); break;   \
      case 8: APAX_TRY(fn<8>(__VA_ARGS__)); break;   \
      case 16: APAX_TRY(fn<16>(__VA_ARGS__)); break; \
      default: APAX_TRY(fn<32>(__VA_ARGS__)); break; \
    }                                              \
  } while (0)

static PairGeom make_geom(const apax_gm_params_impl* p, const double* R, const int32_t* Z, const double* box,
                          const double* offsets, int mode) {
  PairGeom geo{};
  geo.R = R;
  geo.Z = Z;
  geo.offsets = offsets;
  geo.mode = mode;
  geo.n_species = p->cfg.n_species;
  geo.box = box;
  return geo;
}

// moment adjoints from the feature adjoints in GT: thread per atom, or one atom on four warps for small systems
static int launch_contract_bwd(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w) {
  if (w.n_central <= p->contract_split)
    APAX_LAUNCH(k_contract_bwd_split, dim3((unsigned)cdiv(w.n_central, 32)), dim3(kContractThreads), 0, stream, w);
  else
    APAX_LAUNCH(k_contract_bwd, dim3((unsigned)cdiv(w.n_central, kContractThreads)), dim3(kContractThreads), 0, stream, w);
  return APAX_OK;
}

// descriptor forward: moments + contractions -> GT
static int gm_descriptor_impl(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w,
                              const PairGeom& geo, int feature_mode = 0) {
  const DescConst dc = make_desc_const(p->cfg);
  const int lpa = pick_lpa(w.n_central);
  APAX_LAUNCH(k_atom_pack, dim3((unsigned)cdiv(w.n_atoms, 256)), dim3(256), 0, stream, w, geo);
  APAX_DISPATCH_LPA(lpa, launch_pair_fwd, stream, w, geo, dc, (p->generic_basis ? kPairGeneric : p->pair_block));
  const dim3 grid_c((unsigned)cdiv(w.n_central, kContractThreads));
  if (feature_mode == 3) {
    // fused with the layer-0 quantiser of the integer readout: persistent blocks, at most two per SM, one slot each
    const int64_t n_tiles = cdiv(w.n_central, 128);
    if (w.n_central <= p->contract_split) {       // small system: 32-atom groups, one atom on four warps
      const int64_t n_groups = 4 * n_tiles, cap = 8 * int64_t(num_sms());
      APAX_LAUNCH(k_contract_fwd_quantize_split, dim3((unsigned)(n_groups < cap ? n_groups : cap)), dim3(kContractThreads), 0, stream,
                  w, w.q_cs, w.q_planes, w.q_row_exp, n_groups);
    } else {
      const int64_t blocks = n_tiles < 2 * int64_t(num_sms()) ? n_tiles : 2 * int64_t(num_sms());
      APAX_LAUNCH(k_contract_fwd_quantize, dim3((unsigned)blocks), dim3(kContractThreads), 0, stream, w, w.q_cs, w.q_planes,
                  w.q_row_exp, n_tiles);
    }
  } else if (feature_mode == 1) {
    APAX_LAUNCH(k_contract_fwd<1>, grid_c, dim3(kContractThreads), 0, stream, w);
  } else if (feature_mode == 2) {
    APAX_LAUNCH(k_contract_fwd<2>, grid_c, dim3(kContractThreads), 0, stream, w);
  } else {
    APAX_LAUNCH(k_contract_fwd<0>, grid_c, dim3(kContractThreads), 0, stream, w);
  }
  return APAX_OK;
}

// Readout implementation, a field of the parameter handle (apax_gm_params_set_option(APAX_OPT_READOUT)); nothing is read
// from the environment.
//   APAX_READOUT_I8 (default): exact integer tensor-core GEMMs with fused epilogues (i8_gemm.cuh) -- no rounding inside the
//                 reduction at all, every parity test holds, ~2x faster than the CUDA-core path.
//   APAX_READOUT_SIMT: fp32 CUDA-core GEMMs (the first correct path; kept as the cross-check the i8 tests compare with).
//   APAX_READOUT_TC:   tcgen05 3xTF32 GEMMs -- ~1e-6 accurate per output but BIASED (the tensor core's fp32 accumulator
//                 truncates instead of rounding), ~8e-6 relative on total energies: not parity-safe.
enum { READOUT_SIMT = APAX_READOUT_SIMT, READOUT_TC = APAX_READOUT_TC, READOUT_I8 = APAX_READOUT_I8 };

static int energy_sums(cudaStream_t stream, const GmWorkspace& w, int n_out, double* energy, const CorrConst* cc = nullptr,
                       const PairGeom* geo = nullptr) {
  const int64_t N = w.n_central, ld = w.ld;
  if (cc && geo && (cc->zbl_on || cc->exp_on))   // the corrections' energy joins the per-atom energies before they are summed
    APAX_LAUNCH(k_pair_corrections<0>, dim3((unsigned)cdiv(N * 32, 256)), dim3(256), 0, stream, w, *geo, *cc, n_out);
  if (w.n_seg > 0) {
    APAX_LAUNCH(k_energy_segments, dim3((unsigned)w.n_seg, n_out), dim3(32), 0, stream, w.Eatom, ld, w.seg_ptr, n_out, energy);
    return APAX_OK;
  }
  const int n_blk = int(N < 256 * 64 ? 1 : (N / (256 * 64) > 256 ? 256 : N / (256 * 64)));
  APAX_LAUNCH(k_energy_partial, dim3(n_blk, n_out), dim3(256), 0, stream, w.Eatom, ld, N, w.partial);
  APAX_LAUNCH(k_energy_final, dim3(n_out), dim3(32), 0, stream, w.partial, n_blk, energy);
  return APAX_OK;
}

static int readout_head_and_energy(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w,
                                   const PairGeom& geo, const float* h2_lo, double* energy) {
  const int64_t N = w.n_central, ld = w.ld;
  const int n_out = p->cfg.n_out;
  const int64_t m_tiles = cdiv(N, 128);
  if (n_out == 1) {
    APAX_LAUNCH(k_head<1>, dim3((unsigned)m_tiles), dim3(128), 0, stream, w.H2T, h2_lo, ld, N, p->w2, p->b2, n_out,
                geo.Z, p->cfg.n_species, p->scale, p->shift, p->cfg.mask_atoms, w.Eatom, w.ebar);
  } else {
    APAX_LAUNCH(k_head<MAX_OUT>, dim3((unsigned)m_tiles), dim3(128), 0, stream, w.H2T, h2_lo, ld, N, p->w2, p->b2,
                n_out, geo.Z, p->cfg.n_species, p->scale, p->shift, p->cfg.mask_atoms, w.Eatom, w.ebar);
  }
  return energy_sums(stream, w, n_out, energy, &p->corr, &geo);
}

// ---- readout forward + (per head) backward down to dE/dg in GT, on the fp32 CUDA cores
static int readout_simt(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const PairGeom& geo,
                        double* energy, int head /* -1: forward only */) {
  const int64_t N = w.n_central, ld = w.ld;
  const int n_out = p->cfg.n_out;
  const int64_t m_tiles = cdiv(N, 128);
  const dim3 grid_h((unsigned)m_tiles, NH / 128);
  if (head < 0) {
    {  // layer 0: 360 -> 256
      GemmArgs a{};
      a.W = p->w0; a.ldw = NH; a.XT = w.GT; a.YT = w.H1T; a.Y2T = w.D1T; a.bias = p->b0; a.K = NF; a.ld = ld;
      auto k_sgemm_fwd_l0 = k_sgemm<EPI_ACT>;
      APAX_LAUNCH(k_sgemm_fwd_l0, grid_h, dim3(256), 0, stream, a);
    }
    {  // layer 1: 256 -> 256
      GemmArgs a{};
      a.W = p->w1; a.ldw = NH; a.XT = w.H1T; a.YT = w.H2T; a.Y2T = w.D2T; a.bias = p->b1; a.K = NH; a.ld = ld;
      auto k_sgemm_fwd_l1 = k_sgemm<EPI_ACT>;
      APAX_LAUNCH(k_sgemm_fwd_l1, grid_h, dim3(256), 0, stream, a);
    }
    return readout_head_and_energy(stream, p, w, geo, nullptr, energy);
  }
  {  // dE/dy1[n][m] = swish'(y1)[n][m] * ebar[m] * sum_k (W1[n][k] w2[k][head]) swish'(y2)[k][m]: the head's output
     // weights are folded into the transposed layer-1 weights once per parameter set (w1t_head)
    GemmArgs a{};
    a.W = p->w1t_head + size_t(head) * NH * NH; a.ldw = NH; a.XT = w.D2T; a.YT = w.H1T; a.Y2T = w.D1T; a.K = NH; a.ld = ld;
    a.ebar = w.ebar;
    auto k_sgemm_bwd_l1 = k_sgemm<EPI_MULD>;
    APAX_LAUNCH(k_sgemm_bwd_l1, grid_h, dim3(256), 0, stream, a);
  }
  {  // dE/dg = W0 . dE/dy1   (360 rows padded to 384)
    GemmArgs a{};
    a.W = p->w0t; a.ldw = NFP; a.XT = w.H1T; a.YT = w.GT; a.K = NH; a.ld = ld;
    auto k_sgemm_bwd_l0 = k_sgemm<EPI_STORE>;
    APAX_LAUNCH(k_sgemm_bwd_l0, dim3((unsigned)m_tiles, NFP / 128), dim3(256), 0, stream, a);
  }
  return APAX_OK;
}

// ---- the same on the tensor cores: every activation travels as an exact (tf32 hi, lo) pair
static int readout_tc(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const PairGeom& geo,
                      double* energy, int head) {
  const int64_t N = w.n_central, ld = w.ld;
  const int n_out = p->cfg.n_out;
  if (head < 0) {
    {  // layer 0: 360 -> 256, swish + swish' epilogue
      TcOperand act{w.GT, w.GLT, NF, ld}, wgt{p->w0_hi, p->w0_lo, NF, NH};
      TcEpilogue e{};
      e.mode = TC_EPI_ACT; e.ld = ld; e.out_hi = w.H1T; e.out_lo = w.H1LT; e.aux = w.D1T; e.bias = p->b0;
      APAX_TRY(tc_gemm(stream, act, wgt, N, e, p->tc_err));
    }
    {  // layer 1: 256 -> 256
      TcOperand act{w.H1T, w.H1LT, NH, ld}, wgt{p->w1_hi, p->w1_lo, NH, NH};
      TcEpilogue e{};
      e.mode = TC_EPI_ACT; e.ld = ld; e.out_hi = w.H2T; e.out_lo = w.H2LT; e.aux = w.D2T; e.bias = p->b1;
      APAX_TRY(tc_gemm(stream, act, wgt, N, e, p->tc_err));
    }
    return readout_head_and_energy(stream, p, w, geo, w.H2LT, energy);
  }
  const int64_t n_cols = round_up(N, 128);
  APAX_LAUNCH(k_ybar2, dim3((unsigned)cdiv(n_cols, 1024), NH), dim3(256), 0, stream, w.D2T, w.ebar, p->w2, n_out, head, ld,
              n_cols, w.H2T, w.H2LT);  // H2 was consumed by the head kernel: its buffers now carry dE/dy2
  {  // dE/dy1 = (W1 . dE/dy2) * swish'(y1)
    TcOperand act{w.H2T, w.H2LT, NH, ld}, wgt{p->w1t_hi, p->w1t_lo, NH, NH};
    TcEpilogue e{};
    e.mode = TC_EPI_MULD; e.ld = ld; e.out_hi = w.H1T; e.out_lo = w.H1LT; e.aux = w.D1T;
    APAX_TRY(tc_gemm(stream, act, wgt, N, e, p->tc_err));
  }
  {  // dE/dg = W0 . dE/dy1 (384 padded rows)
    TcOperand act{w.H1T, w.H1LT, NH, ld}, wgt{p->w0t_hi, p->w0t_lo, NH, NFP};
    TcEpilogue e{};
    e.mode = TC_EPI_STORE; e.ld = ld; e.out_hi = w.GT;
    APAX_TRY(tc_gemm(stream, act, wgt, N, e, p->tc_err));
  }
  return APAX_OK;
}

// ---- the same on the integer tensor cores (i8_gemm.cuh): activations travel between the four GEMMs as base-256
// digit planes written by the epilogues; fp32 copies exist only of swish'(y1) (D1T) and of dE/dg (GT).
// Buffer reuse: descriptor planes in GLT, swish(y1) / dE/dy1 planes in H1LT, swish'(y2) planes in H2LT, the epilogue
// staging in H2T, and head sums + row exponents in the first rows of H1T (none of these hold fp32 data here).
static int32_t* i8_rowmax_buffer(const GmWorkspace& w) {
  return reinterpret_cast<int32_t*>(w.H1T + int64_t(MAX_OUT) * w.ld) + 3 * w.ld;
}

static int readout_i8(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const PairGeom& geo,
                      double* energy, int head) {
  const int64_t N = w.n_central, ld = w.ld;
  // rows the GEMMs sweep: the evaluated (central) atoms only, in whole 128-atom tiles.  The plane arrays keep the stride of
  // ALL local atoms (ld), but ghost rows of a domain-decomposed box are never quantised, multiplied or read.
  const int64_t rows = round_up(N, 128);
  const int n_out = p->cfg.n_out;
  int8_t* gq = reinterpret_cast<int8_t*>(w.GLT);
  int8_t* h1q = reinterpret_cast<int8_t*>(w.H1LT);
  int8_t* d2q = reinterpret_cast<int8_t*>(w.H2LT);
  float* h_out = w.H1T;                                                   // [n_out <= 16][ld]
  int32_t* e_g = reinterpret_cast<int32_t*>(w.H1T + int64_t(MAX_OUT) * ld);
  int32_t* e_h1 = e_g + ld;
  int32_t* e_x1 = e_h1 + ld;
  int32_t* rowmax = i8_rowmax_buffer(w);   // filled by k_contract_fwd<2>
  const I8Planes w0{p->q_w0, NH, NFP, int64_t(NH) * NFP, 4}, w1{p->q_w1, NH, NH, int64_t(NH) * NH, 4};
  if (head < 0) {
    if (!w.q_planes)   // otherwise the contraction kernel has emitted the descriptor's digit planes and exponents itself
      APAX_TRY(i8_quantize_rows(stream, w.GT, ld, N, NF, NFP, 4, p->cs_g, gq, e_g, rowmax, /*have_rowmax=*/true, rows));
    {  // layer 0: 360 -> 256; swish' kept in fp32 for the backward pass, swish goes on as digit planes
      const I8Planes act{gq, rows, NFP, ld * int64_t(NFP), 4};
      I8Epilogue e{};
      e.mode = I8_EPI_ACT_Q; e.ld = ld; e.n_valid = NH; e.row_exp = e_g; e.bias = p->b0; e.aux = w.D1T;
      e.q_planes = h1q; e.q_kp = NH; e.q_plane_stride = ld * int64_t(NH); e.q_col_scale = p->cs_h1; e.q_row_exp = e_h1;
      e.scratch = w.H2T;
      APAX_TRY(i8_gemm(stream, act, w0, e, p->tc_err));
    }
    {  // layer 1: 256 -> 256, output layer contracted in the epilogue; swish' goes on as fixed-scale digit planes
      const I8Planes act{h1q, rows, NH, ld * int64_t(NH), 4};
      I8Epilogue e{};
      e.mode = I8_EPI_ACT_HEAD; e.ld = ld; e.n_valid = NH; e.row_exp = e_h1; e.bias = p->b1;
      e.q_planes = d2q; e.q_kp = NH; e.q_plane_stride = ld * int64_t(NH); e.q_col_scale = p->cs_d2; e.q_fixed_exp = p->d2_exp;
      e.w_head = p->w2; e.n_out = n_out; e.h_out = h_out;
      APAX_TRY(i8_gemm(stream, act, w1, e, p->tc_err));
    }
    const int64_t m_tiles = cdiv(N, 128);
    if (n_out == 1) {
      APAX_LAUNCH(k_head_scale_shift<1>, dim3((unsigned)m_tiles), dim3(128), 0, stream, h_out, ld, N, p->b2, n_out, geo.Z,
                  p->cfg.n_species, p->scale, p->shift, p->cfg.mask_atoms, w.Eatom, w.ebar);
    } else {
      APAX_LAUNCH(k_head_scale_shift<MAX_OUT>, dim3((unsigned)m_tiles), dim3(128), 0, stream, h_out, ld, N, p->b2, n_out,
                  geo.Z, p->cfg.n_species, p->scale, p->shift, p->cfg.mask_atoms, w.Eatom, w.ebar);
    }
    return energy_sums(stream, w, n_out, energy, &p->corr, &geo);
  }
  {  // dE/dy1 = swish'(y1) * ebar * (W1 w2[head]) . swish'(y2)  -> digit planes
    const I8Planes act{d2q, rows, NH, ld * int64_t(NH), 4};
    const I8Planes wth{p->q_w1th + size_t(head) * 4 * NH * NH, NH, NH, int64_t(NH) * NH, 4};
    I8Epilogue e{};
    e.mode = I8_EPI_MULD_Q; e.ld = ld; e.n_valid = NH; e.row_exp = nullptr; e.fixed_exp = p->d2_exp; e.aux = w.D1T;
    e.row_mul = w.ebar;
    e.q_planes = h1q; e.q_kp = NH; e.q_plane_stride = ld * int64_t(NH); e.q_col_scale = p->cs_x1; e.q_row_exp = e_x1;
    e.scratch = w.H2T;
    APAX_TRY(i8_gemm(stream, act, wth, e, p->tc_err));
  }
  {  // dE/dg = W0 . dE/dy1 (384 padded rows, fp32 for the contraction backward)
    const I8Planes act{h1q, rows, NH, ld * int64_t(NH), 4};
    const I8Planes w0t{p->q_w0t, NFP, NH, int64_t(NFP) * NH, 4};
    I8Epilogue e{};
    e.mode = I8_EPI_STORE; e.ld = ld; e.n_valid = NFP; e.row_exp = e_x1; e.out = w.GT;
    APAX_TRY(i8_gemm(stream, act, w0t, e, p->tc_err));
  }
  return APAX_OK;
}

static int gm_compute_impl(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w,
                           const PairGeom& geo, double* energy, double* forces) {
  const int64_t N = w.n_central;  // per-central work only; ghosts enter as neighbours
  const int n_out = p->cfg.n_out;
  const DescConst dc = make_desc_const(p->cfg);
  const int lpa = pick_lpa(N);
  const int mode = p->readout_mode;
  GmWorkspace wq = w;
  auto readout = [&](int head) {
    return mode == READOUT_I8   ? readout_i8(stream, p, wq, geo, energy, head)
           : mode == READOUT_TC ? readout_tc(stream, p, w, geo, energy, head)
                                : readout_simt(stream, p, w, geo, energy, head);
  };
  if (mode == READOUT_I8) {
    // the contraction kernel emits the layer-0 digit planes and row exponents itself (k_contract_fwd_quantize)
    wq.q_cs = p->cs_g;
    wq.q_rowmax = i8_rowmax_buffer(w);
    wq.q_planes = reinterpret_cast<int8_t*>(w.GLT);
    wq.q_row_exp = reinterpret_cast<int32_t*>(w.H1T + int64_t(MAX_OUT) * w.ld);     // e_g of readout_i8
    APAX_TRY(gm_descriptor_impl(stream, p, wq, geo, 3));
  } else {
    APAX_TRY(gm_descriptor_impl(stream, p, w, geo, mode == READOUT_TC ? 1 : 0));
  }
  APAX_TRY(readout(-1));
  if (!forces) return APAX_OK;
  // mean_forces (apax_gm_params_set_option): ONE backward pass with the mean head's output weights (slot n_out of the folded
  // backward weights) gives -d(mean_h E_h)/dR, instead of one pass per head
  const bool mean_only = p->mean_forces && n_out > 1;
  if (mean_only && mode == READOUT_TC) return APAX_ERR_UNSUPPORTED;
  const int n_passes = mean_only ? 1 : n_out;
  for (int pass = 0; pass < n_passes; ++pass) {
    const int head = mean_only ? n_out : pass;
    // every backward pass reads only what the forward pass left behind (D1, D2 / its planes, ebar): nothing to redo
    APAX_TRY(readout(head));
    APAX_TRY(launch_contract_bwd(stream, p, w));
    APAX_DISPATCH_LPA(lpa, launch_pair_bwd, stream, w, dc, (p->generic_basis ? kPairGeneric : p->pair_block));
    if (p->corr.zbl_on || p->corr.exp_on)
      APAX_LAUNCH(k_pair_corrections<1>, dim3((unsigned)cdiv(N * 32, 256)), dim3(256), 0, stream, w, geo, p->corr, n_out);
    APAX_DISPATCH_LPA(pick_lpa(w.n_atoms), launch_force_gather, stream, w, forces + int64_t(pass) * w.n_atoms * 3);
  }
  return APAX_OK;
}

int gm_compute_device(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const double* R,
                      const int32_t* Z, const double* box, const double* offsets, int mode, double* energy,
                      double* forces, int64_t n_central) {
  const PairGeom geo = make_geom(p, R, Z, box, offsets, mode);
  GmWorkspace wc = w;
  if (n_central >= 0) wc.n_central = n_central;
  return gm_compute_impl(stream, p, wc, geo, energy, forces);
}

}  // namespace apax

// ================================================================================================
// C ABI
// ================================================================================================
using namespace apax;

static int validate_cfg(const apax_gm_config* c) {
  if (!c) return APAX_ERR_INVALID_ARG;
  if (c->n_radial != NR || c->n_basis < 1 || c->n_basis > MAXNB_GEN || c->n_contr < 1 || c->n_contr > 8 || c->n_hidden1 != NH ||
      c->n_hidden2 != NH)
    return APAX_ERR_UNSUPPORTED;
  if (c->n_out < 1 || c->n_out > MAX_OUT || c->n_species < 1 || c->n_species > 128) return APAX_ERR_UNSUPPORTED;
  if (!(c->r_max > c->r_min) || !(c->r_min >= 0.f)) return APAX_ERR_INVALID_ARG;
  if (c->basis_kind < 0 || c->basis_kind > 2) return APAX_ERR_UNSUPPORTED;
  return APAX_OK;
}

struct apax_gm_params : apax_gm_params_impl {};

template <class T>
static int upload(T** dst, const T* src, size_t n) {
  APAX_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T)));
  APAX_CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return APAX_OK;
}

extern "C" int apax_gm_params_create(const apax_gm_config* cfg, const float* emb, const float* w0, const float* b0,
                                     const float* w1, const float* b1, const float* w2, const float* b2,
                                     const double* scale, const double* shift, apax_gm_params** out) {
  APAX_TRY(validate_cfg(cfg));
  if (!emb || !w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !scale || !shift || !out) return APAX_ERR_INVALID_ARG;
  apax_gm_params* p = new apax_gm_params();
  p->cfg = *cfg;
  p->readout_mode = APAX_READOUT_I8;
  p->generic_basis = (cfg->basis_kind != 0 || cfg->n_basis != NB) ? 1 : 0;   // pair kernels with the [5][16] coefficient table
  p->pair_block = 1;
  p->mean_forces = 0;
  p->contract_split = 8192;    // one wave of 32-atom groups on 148 SMs x 2 CTAs; tools/split_ab.py: -16 % at 1,026 atoms, break-even at 12,288
  const int S = cfg->n_species, n_out = cfg->n_out;
  // n_contr < 8 truncates the feature list c0|c1|...|c7 (gaussian_moment_descriptor.py:101): the caller's first dense layer
  // has n_feat rows; the kernels always produce all 360 features, the dropped ones meet zero weight rows (and receive a
  // zero adjoint), which is the same function.
  static const int kFeatCount[8] = {5, 20, 35, 50, 85, 160, 235, 360};
  const int n_feat = kFeatCount[cfg->n_contr - 1];
  std::vector<float> w0_padded(size_t(NF) * NH, 0.0f);
  std::copy(w0, w0 + size_t(n_feat) * NH, w0_padded.begin());
  w0 = w0_padded.data();
  // NTKLinear (ntk_linear.py:42-45): y = sqrt(1/in) * x.w + 0.1 * b, folded into the stored weights
  auto wf = [&](int fan_in) { return cfg->use_ntk ? float(sqrt(1.0 / fan_in)) : 1.0f; };
  const float bf = cfg->use_ntk ? 0.1f : 1.0f;
  {
    const size_t n = size_t(S) * S * NR * cfg->n_basis;
    float* h = new float[n];
    const float sc = cfg->use_embed_norm ? float(1.0 / sqrt(double(cfg->n_basis))) : 1.0f;
    for (size_t i = 0; i < n; ++i) h[i] = sc * emb[i];  // embed_norm * coeffs (basis_functions.py:145-146)
    int st = upload(&p->emb, h, n);
    delete[] h;
    if (st) return st;
  }
  {
    float* h = new float[size_t(NF) * NH];
    float* ht = new float[size_t(NH) * NFP]();
    for (int k = 0; k < NF; ++k)
      for (int n = 0; n < NH; ++n) {
        h[size_t(k) * NH + n] = wf(n_feat) * w0[size_t(k) * NH + n];
        ht[size_t(n) * NFP + k] = h[size_t(k) * NH + n];
      }
    int st = upload(&p->w0, h, size_t(NF) * NH);
    if (!st) st = upload(&p->w0t, ht, size_t(NH) * NFP);
    delete[] h;
    delete[] ht;
    if (st) return st;
  }
  {
    float* h = new float[size_t(NH) * NH];
    float* ht = new float[size_t(NH) * NH];
    for (int k = 0; k < NH; ++k)
      for (int n = 0; n < NH; ++n) {
        h[size_t(k) * NH + n] = wf(NH) * w1[size_t(k) * NH + n];
        ht[size_t(n) * NH + k] = h[size_t(k) * NH + n];
      }
    int st = upload(&p->w1, h, size_t(NH) * NH);
    if (!st) st = upload(&p->w1t, ht, size_t(NH) * NH);
    delete[] h;
    delete[] ht;
    if (st) return st;
  }
  {
    float* h = new float[size_t(NH) * n_out];
    for (size_t i = 0; i < size_t(NH) * n_out; ++i) h[i] = wf(NH) * w2[i];
    int st = upload(&p->w2, h, size_t(NH) * n_out);
    // w1t_head[head][k][n] = W1[n][k] * w2[k][head]  (reduction index k = layer-1 output feature); slot n_out holds the
    // mean head w2_mean[k] = mean_h w2[k][h]: the gradient of the MEAN energy (ShallowEnsembleModel with force_variance=False,
    // apax/nn/models.py:265-267, and the MD energy_fn of a shallow ensemble, apax/md/simulate.py:55-58) in one backward pass
    float* hh = new float[size_t(n_out + 1) * NH * NH];
    for (int hd = 0; hd <= n_out; ++hd)
      for (int k = 0; k < NH; ++k) {
        float w2k = 0.0f;
        if (hd < n_out) {
          w2k = h[size_t(k) * n_out + hd];
        } else {
          double acc = 0.0;
          for (int q = 0; q < n_out; ++q) acc += double(h[size_t(k) * n_out + q]);
          w2k = float(acc / n_out);
        }
        for (int n = 0; n < NH; ++n) hh[(size_t(hd) * NH + k) * NH + n] = (wf(NH) * w1[size_t(n) * NH + k]) * w2k;
      }
    if (!st) st = upload(&p->w1t_head, hh, size_t(n_out + 1) * NH * NH);
    delete[] hh;
    delete[] h;
    if (st) return st;
  }
  {
    float hb0[NH], hb1[NH], hb2[MAX_OUT];
    for (int i = 0; i < NH; ++i) { hb0[i] = bf * b0[i]; hb1[i] = bf * b1[i]; }
    for (int i = 0; i < n_out; ++i) hb2[i] = bf * b2[i];
    APAX_TRY(upload(&p->b0, hb0, NH));
    APAX_TRY(upload(&p->b1, hb1, NH));
    APAX_TRY(upload(&p->b2, hb2, n_out));
  }
  APAX_TRY(upload(&p->scale, scale, S));
  APAX_TRY(upload(&p->shift, shift, S));
  {  // exact (tf32 hi, lo) pairs of the (NTK-folded) weights for the tensor-core readout
    auto split_upload = [&](float** dhi, float** dlo, int rows, int cols, auto get) -> int {
      const size_t n = size_t(rows) * cols;
      float* hi = new float[n];
      float* lo = new float[n];
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) split_tf32(get(r, c), hi[size_t(r) * cols + c], lo[size_t(r) * cols + c]);
      int st = upload(dhi, hi, n);
      if (!st) st = upload(dlo, lo, n);
      delete[] hi;
      delete[] lo;
      return st;
    };
    const float f0 = wf(n_feat), f1 = wf(NH);
    APAX_TRY(split_upload(&p->w0_hi, &p->w0_lo, NF, NH, [&](int k, int n) { return f0 * w0[size_t(k) * NH + n]; }));
    APAX_TRY(split_upload(&p->w1_hi, &p->w1_lo, NH, NH, [&](int k, int n) { return f1 * w1[size_t(k) * NH + n]; }));
    APAX_TRY(split_upload(&p->w1t_hi, &p->w1t_lo, NH, NH, [&](int n, int k) { return f1 * w1[size_t(k) * NH + n]; }));
    APAX_TRY(split_upload(&p->w0t_hi, &p->w0t_lo, NH, NFP,
                          [&](int n, int k) { return k < NF ? f0 * w0[size_t(k) * NH + n] : 0.0f; }));
    {  // digit planes of the same weights for the integer tensor-core readout
      std::vector<float> W;
      std::vector<int8_t> planes;
      std::vector<float> cs;
      W.assign(size_t(NF) * NH, 0.0f);
      for (int k = 0; k < NF; ++k)
        for (int n = 0; n < NH; ++n) W[size_t(k) * NH + n] = f0 * w0[size_t(k) * NH + n];
      i8_quantize_weights(W.data(), NF, NH, NH, NFP, NH, planes, cs);
      APAX_TRY(upload(&p->q_w0, planes.data(), planes.size()));
      APAX_TRY(upload(&p->cs_g, cs.data(), cs.size()));
      W.assign(size_t(NH) * NH, 0.0f);
      for (int k = 0; k < NH; ++k)
        for (int n = 0; n < NH; ++n) W[size_t(k) * NH + n] = f1 * w1[size_t(k) * NH + n];
      i8_quantize_weights(W.data(), NH, NH, NH, NH, NH, planes, cs);
      APAX_TRY(upload(&p->q_w1, planes.data(), planes.size()));
      APAX_TRY(upload(&p->cs_h1, cs.data(), cs.size()));
      // backward layer 1: one weight matrix per head (+ the mean head in slot n_out), one scale vector for all of them
      std::vector<float> Wh(size_t(n_out + 1) * NH * NH);
      for (int hd = 0; hd <= n_out; ++hd)
        for (int k = 0; k < NH; ++k) {
          float w2k = 0.0f;
          if (hd < n_out) {
            w2k = f1 * w2[size_t(k) * n_out + hd];
          } else {
            double acc = 0.0;
            for (int q = 0; q < n_out; ++q) acc += double(f1 * w2[size_t(k) * n_out + q]);
            w2k = float(acc / n_out);
          }
          for (int n = 0; n < NH; ++n) Wh[(size_t(hd) * NH + k) * NH + n] = (f1 * w1[size_t(n) * NH + k]) * w2k;
        }
      for (int hd = 0; hd <= n_out; ++hd) i8_weight_scales(Wh.data() + size_t(hd) * NH * NH, NH, NH, NH, NH, cs, hd > 0);
      std::vector<int8_t> all;
      for (int hd = 0; hd <= n_out; ++hd) {
        i8_quantize_weights(Wh.data() + size_t(hd) * NH * NH, NH, NH, NH, NH, NH, planes, cs, true);
        all.insert(all.end(), planes.begin(), planes.end());
      }
      APAX_TRY(upload(&p->q_w1th, all.data(), all.size()));
      APAX_TRY(upload(&p->cs_d2, cs.data(), cs.size()));
      float cmax = 0.0f;
      for (int k = 0; k < NH; ++k) cmax = fmaxf(cmax, cs[k]);
      p->d2_exp = int(ceil(log2(1.85 * double(cmax) / 63.0)));   // |swish'| <= 1.845: every scaled value stays below 64
      W.assign(size_t(NH) * NFP, 0.0f);
      for (int n = 0; n < NH; ++n)
        for (int k = 0; k < NF; ++k) W[size_t(n) * NFP + k] = f0 * w0[size_t(k) * NH + n];
      i8_quantize_weights(W.data(), NH, NFP, NFP, NH, NFP, planes, cs);
      APAX_TRY(upload(&p->q_w0t, planes.data(), planes.size()));
      APAX_TRY(upload(&p->cs_x1, cs.data(), cs.size()));
    }
    APAX_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&p->tc_err), sizeof(int32_t)));
    APAX_CUDA_TRY(cudaMemset(p->tc_err, 0, sizeof(int32_t)));
    const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    APAX_TRY(upload(&p->eye, eye, 9));
  }
  *out = p;
  return APAX_OK;
}

extern "C" void apax_gm_params_destroy(apax_gm_params* p) {
  if (!p) return;
  cudaFree(p->emb); cudaFree(p->w0); cudaFree(p->w0t); cudaFree(p->b0); cudaFree(p->w1); cudaFree(p->w1t);
  cudaFree(p->b1); cudaFree(p->w2); cudaFree(p->b2); cudaFree(p->scale); cudaFree(p->shift);
  cudaFree(p->w1t_head);
  cudaFree(p->eye);
  if (p->corr.exp_r) cudaFree(const_cast<float*>(p->corr.exp_r));
  if (p->corr.exp_a) cudaFree(const_cast<float*>(p->corr.exp_a));
  cudaFree(p->w0_hi); cudaFree(p->w0_lo); cudaFree(p->w1_hi); cudaFree(p->w1_lo); cudaFree(p->w1t_hi);
  cudaFree(p->w1t_lo); cudaFree(p->w0t_hi); cudaFree(p->w0t_lo); cudaFree(p->tc_err);
  cudaFree(p->q_w0); cudaFree(p->q_w1); cudaFree(p->q_w1th); cudaFree(p->q_w0t);
  cudaFree(p->cs_g); cudaFree(p->cs_h1); cudaFree(p->cs_d2); cudaFree(p->cs_x1);
  delete p;
}

extern "C" int apax_gm_params_set_option(apax_gm_params* p, int32_t option, int32_t value) {
  if (!p) return APAX_ERR_INVALID_ARG;
  switch (option) {
    case APAX_OPT_READOUT:
      if (value != APAX_READOUT_I8 && value != APAX_READOUT_SIMT && value != APAX_READOUT_TC) return APAX_ERR_INVALID_ARG;
      p->readout_mode = value;
      return APAX_OK;
    case APAX_OPT_MEAN_FORCES:
      if (value != 0 && value != 1) return APAX_ERR_INVALID_ARG;
      p->mean_forces = value;
      return APAX_OK;
    case APAX_OPT_CONTRACT_SPLIT:
      if (value < 0) return APAX_ERR_INVALID_ARG;
      p->contract_split = value;
      return APAX_OK;
    case APAX_OPT_PAIR_BLOCK:
      if (value < 0 || value > 4) return APAX_ERR_INVALID_ARG;
      p->pair_block = value;
      return APAX_OK;
    default:
      return APAX_ERR_INVALID_ARG;
  }
}

extern "C" int apax_gm_params_status(const apax_gm_params* p) {
  // synchronising query of the tensor-core pipeline's time-out flag (0 = healthy)
  if (!p) return APAX_ERR_INVALID_ARG;
  int32_t flag = 0;
  APAX_CUDA_TRY(cudaMemcpy(&flag, p->tc_err, sizeof(flag), cudaMemcpyDeviceToHost));
  return flag == 0 ? APAX_OK : APAX_ERR_CUDA;
}

extern "C" size_t apax_gm_workspace_bytes(int64_t n_atoms, int64_t n_pairs, int32_t n_out) {
  if (n_atoms < 0 || n_pairs < 0 || n_out < 1) return 0;
  return carve_gm_workspace(nullptr, n_atoms, n_pairs, n_out, 128, nullptr);
}

static int get_ws(const apax_gm_params* p, int64_t n_atoms, int64_t n_pairs, void* workspace, size_t bytes,
                  GmWorkspace* w) {
  if (!p || !workspace || n_atoms < 1 || n_pairs < 0) return APAX_ERR_INVALID_ARG;
  if (n_atoms >= (int64_t(1) << 31) - 256 || n_pairs >= (int64_t(1) << 31) - 256) return APAX_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return APAX_ERR_INVALID_ARG;
  const size_t need = carve_gm_workspace(workspace, n_atoms, n_pairs, p->cfg.n_out, 128, w);
  if (need > bytes) return APAX_ERR_WORKSPACE;
  return APAX_OK;
}

extern "C" int apax_gm_prepare(apax_stream_t stream, const apax_gm_params* params, const int32_t* Z,
                               const int32_t* idx, int64_t n_atoms, int64_t n_pairs, void* workspace,
                               size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!Z || (n_pairs > 0 && !idx)) return APAX_ERR_INVALID_ARG;
  return gm_prepare_impl(stream, params, w, Z, idx);
}

extern "C" int apax_gm_compute(apax_stream_t stream, const apax_gm_params* params, const double* R,
                               const int32_t* Z, const double* box, const double* offsets, int64_t n_atoms,
                               int64_t n_pairs, int32_t disp_mode, double* energy, double* forces, void* workspace,
                               size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!R || !Z || !energy) return APAX_ERR_INVALID_ARG;
  if (disp_mode < APAX_DISP_GAS || disp_mode > APAX_DISP_MINIMAGE) return APAX_ERR_INVALID_ARG;
  if (disp_mode != APAX_DISP_GAS && !box) return APAX_ERR_INVALID_ARG;
  const PairGeom geo = make_geom(params, R, Z, box, offsets, disp_mode);
  return gm_compute_impl(stream, params, w, geo, energy, forces);
}

extern "C" int apax_gm_compute_partial(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                       const int32_t* Z, const double* box, const double* offsets, int64_t n_atoms,
                                       int64_t n_pairs, int64_t n_central, int32_t disp_mode, double* energy,
                                       double* forces, void* workspace, size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!R || !Z || !energy || n_central < 1 || n_central > n_atoms) return APAX_ERR_INVALID_ARG;
  if (disp_mode < APAX_DISP_GAS || disp_mode > APAX_DISP_MINIMAGE) return APAX_ERR_INVALID_ARG;
  if (disp_mode != APAX_DISP_GAS && !box) return APAX_ERR_INVALID_ARG;
  return gm_compute_device(stream, params, w, R, Z, box, offsets, disp_mode, energy, forces, n_central);
}

extern "C" int apax_gm_energy_forces(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                     const int32_t* Z, const int32_t* idx, const double* box, const double* offsets,
                                     int64_t n_atoms, int64_t n_pairs, int32_t disp_mode, double* energy,
                                     double* forces, void* workspace, size_t workspace_bytes) {
  APAX_TRY(apax_gm_prepare(stream, params, Z, idx, n_atoms, n_pairs, workspace, workspace_bytes));
  return apax_gm_compute(stream, params, R, Z, box, offsets, n_atoms, n_pairs, disp_mode, energy, forces, workspace,
                         workspace_bytes);
}

extern "C" int apax_gm_energy_forces_batched(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                             const int32_t* Z, const int32_t* idx, const double* offsets,
                                             const int32_t* struct_ptr, int64_t n_structs, int64_t n_atoms,
                                             int64_t n_pairs, double* energy, double* forces, void* workspace,
                                             size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!R || !Z || !energy || !struct_ptr || n_structs < 1 || (n_pairs > 0 && !idx)) return APAX_ERR_INVALID_ARG;
  APAX_TRY(gm_prepare_impl(stream, params, w, Z, idx));
  w.seg_ptr = struct_ptr;
  w.n_seg = n_structs;
  const PairGeom geo = make_geom(params, R, Z, params->eye, offsets, APAX_DISP_OFFSETS);
  return gm_compute_impl(stream, params, w, geo, energy, forces);
}

extern "C" int apax_gm_descriptor(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                  const int32_t* Z, const int32_t* idx, const double* box, const double* offsets,
                                  int64_t n_atoms, int64_t n_pairs, int32_t disp_mode, float* g, void* workspace,
                                  size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!R || !Z || !g || (n_pairs > 0 && !idx)) return APAX_ERR_INVALID_ARG;
  if (disp_mode < APAX_DISP_GAS || disp_mode > APAX_DISP_DRVEC) return APAX_ERR_INVALID_ARG;
  APAX_TRY(gm_prepare_impl(stream, params, w, Z, idx));
  if ((disp_mode == APAX_DISP_OFFSETS || disp_mode == APAX_DISP_MINIMAGE) && !box) return APAX_ERR_INVALID_ARG;
  const PairGeom geo = make_geom(params, R, Z, box, offsets, disp_mode);
  APAX_TRY(gm_descriptor_impl(stream, params, w, geo));
  APAX_LAUNCH(k_features_to_rowmajor, dim3((unsigned)cdiv(n_atoms, 32), cdiv(NF, 32)), dim3(256), 0, stream, w.GT,
              w.ld, n_atoms, g);
  return APAX_OK;
}

extern "C" int apax_gm_descriptor_vjp(apax_stream_t stream_, const apax_gm_params* params, const double* R, const int32_t* Z,
                                      const int32_t* idx, const double* box, const double* offsets, int64_t n_atoms,
                                      int64_t n_pairs, int32_t disp_mode, const float* g_bar, double* dr_vec_bar, double* R_bar,
                                      void* workspace, size_t workspace_bytes) {
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  if (!R || !Z || !g_bar || (!dr_vec_bar && !R_bar) || (n_pairs > 0 && !idx)) return APAX_ERR_INVALID_ARG;
  if (disp_mode < APAX_DISP_GAS || disp_mode > APAX_DISP_DRVEC) return APAX_ERR_INVALID_ARG;
  if ((disp_mode == APAX_DISP_OFFSETS || disp_mode == APAX_DISP_MINIMAGE) && !box) return APAX_ERR_INVALID_ARG;
  if (disp_mode == APAX_DISP_DRVEC && R_bar) return APAX_ERR_INVALID_ARG;   // no positions in that mode
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const apax_gm_params_impl* p = params;
  APAX_TRY(gm_prepare_impl(stream, p, w, Z, idx));
  const PairGeom geo = make_geom(p, R, Z, box, offsets, disp_mode);
  const DescConst dc = make_desc_const(p->cfg);
  const int lpa = pick_lpa(w.n_central);
  // forward pair pass: the packed moments (the contractions are polynomial in them) and the per-pair displacements
  APAX_LAUNCH(k_atom_pack, dim3((unsigned)cdiv(w.n_atoms, 256)), dim3(256), 0, stream, w, geo);
  APAX_DISPATCH_LPA(lpa, launch_pair_fwd, stream, w, geo, dc, (p->generic_basis ? kPairGeneric : p->pair_block));
  // reverse: g_bar -> moment adjoints -> per-pair dE/d(dr_vec)
  APAX_LAUNCH(k_features_from_rowmajor, dim3((unsigned)cdiv(w.ld, 32), cdiv(NF, 32)), dim3(256), 0, stream, g_bar, w.ld, n_atoms,
              w.GT);
  APAX_TRY(launch_contract_bwd(stream, p, w));
  APAX_DISPATCH_LPA(lpa, launch_pair_bwd, stream, w, dc, (p->generic_basis ? kPairGeneric : p->pair_block));
  if (dr_vec_bar) {
    APAX_CUDA_TRY(cudaMemsetAsync(dr_vec_bar, 0, sizeof(double) * 3 * size_t(n_pairs), stream));
    if (n_pairs > 0)
      APAX_LAUNCH(k_pair_grad_to_coo, dim3((unsigned)cdiv(n_pairs, 256)), dim3(256), 0, stream, w.D, w.src, w.rowptr, w.n_central,
                  dr_vec_bar);
  }
  if (R_bar) {
    // dr_vec = T(R[idx[1]] - R[idx[0]]) with the identity Jacobian of space.transform (space.py:93-97): the transpose of the two
    // gathers, exactly what the force assembly computes with the opposite sign
    APAX_DISPATCH_LPA(pick_lpa(w.n_atoms), launch_force_gather, stream, w, R_bar);
    APAX_LAUNCH(k_negate_f64, dim3((unsigned)cdiv(3 * n_atoms, 256)), dim3(256), 0, stream, R_bar, 3 * n_atoms);
  }
  return APAX_OK;
}

extern "C" int apax_gm_atomic_energies(apax_stream_t stream, const apax_gm_params* params, int64_t n_atoms, int64_t n_pairs,
                                       double* e_atoms, void* workspace, size_t workspace_bytes) {
  if (!params || !e_atoms || !workspace) return APAX_ERR_INVALID_ARG;
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  // Eatom is [n_out][ld] (ld = n_atoms rounded up to 128): a strided copy to the caller's dense [n_out][n_atoms]
  APAX_CUDA_TRY(cudaMemcpy2DAsync(e_atoms, sizeof(double) * n_atoms, w.Eatom, sizeof(double) * w.ld, sizeof(double) * n_atoms,
                                  size_t(params->cfg.n_out), cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return APAX_OK;
}

extern "C" int apax_gm_stress_times_vol(apax_stream_t stream, const apax_gm_params* params, int64_t n_atoms, int64_t n_pairs,
                                        double* stress, void* workspace, size_t workspace_bytes) {
  if (!params || !stress || !workspace) return APAX_ERR_INVALID_ARG;
  if (params->cfg.n_out != 1) return APAX_ERR_UNSUPPORTED;
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // scratch for the per-block partial sums: the int32 [n_pairs] array of the list preparation (idle after prepare); lists
  // too short to hold 1024 x 9 doubles are summed by one block into the row-force scratch (idle after the force assembly)
  double* partial = reinterpret_cast<double*>(w.tmp_a);
  if (size_t(w.n_pairs) * sizeof(int32_t) < size_t(kVirialBlocks) * 9 * sizeof(double)) {
    APAX_LAUNCH(k_virial_partial, dim3(1), dim3(256), 0, s, w.D, w.G, w.rowptr, w.n_atoms, w.Frow);
    APAX_LAUNCH(k_virial_final, dim3(1), dim3(32), 0, s, w.Frow, 1, stress);
    return APAX_OK;
  }
  APAX_LAUNCH(k_virial_partial, dim3(kVirialBlocks), dim3(256), 0, s, w.D, w.G, w.rowptr, w.n_atoms, partial);
  APAX_LAUNCH(k_virial_final, dim3(1), dim3(32), 0, s, partial, kVirialBlocks, stress);
  return APAX_OK;
}

extern "C" int apax_gm_stress_times_vol_offsets(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                                const double* box, int64_t n_atoms, int64_t n_pairs, double* stress,
                                                void* workspace, size_t workspace_bytes) {
  if (!params || !stress || !workspace || !R || !box) return APAX_ERR_INVALID_ARG;
  if (params->cfg.n_out != 1) return APAX_ERR_UNSUPPORTED;
  GmWorkspace w;
  APAX_TRY(get_ws(params, n_atoms, n_pairs, workspace, workspace_bytes, &w));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* partial = reinterpret_cast<double*>(w.tmp_a);   // as in apax_gm_stress_times_vol
  if (size_t(w.n_pairs) * sizeof(int32_t) < size_t(kVirialBlocks) * 9 * sizeof(double)) {
    APAX_LAUNCH(k_virial_partial_rows, dim3(1), dim3(256), 0, s, w, R, box, w.Frow);
    APAX_LAUNCH(k_virial_final, dim3(1), dim3(32), 0, s, w.Frow, 1, stress);
    return APAX_OK;
  }
  APAX_LAUNCH(k_virial_partial_rows, dim3(kVirialBlocks), dim3(256), 0, s, w, R, box, partial);
  APAX_LAUNCH(k_virial_final, dim3(1), dim3(32), 0, s, partial, kVirialBlocks, stress);
  return APAX_OK;
}

extern "C" int apax_gm_params_set_corrections(apax_gm_params* p, const apax_gm_corrections* c) {
  if (!p || !c) return APAX_ERR_INVALID_ARG;
  auto softplus = [](double x) { return x > 30.0 ? x : log1p(exp(x)); };
  CorrConst cc{};
  if (c->zbl_on) {
    if (!(c->zbl_r_max > 0.02f)) return APAX_ERR_INVALID_ARG;
    static const float kExp[4] = {3.19980f, 0.94229f, 0.4029f, 0.20162f};   // empirical.py:45
    cc.zbl_on = 1;
    cc.zbl_rmax = c->zbl_r_max;
    cc.zbl_scale = float(softplus(double(c->zbl_rep_scale)));                // jax.nn.softplus(rep_scale) (:70)
    for (int q = 0; q < 4; ++q) { cc.zbl_c[q] = float(softplus(double(c->zbl_coefficients[q]))); cc.zbl_e[q] = kExp[q]; }
  }
  if (c->exp_on) {
    if (!(c->exp_r_max > 0.02f) || !c->exp_rep_scale_host || !c->exp_rep_prefactor_host) return APAX_ERR_INVALID_ARG;
    const int S = p->cfg.n_species;
    std::vector<float> r(S), a(S);
    for (int z = 0; z < S; ++z) { r[z] = fabsf(c->exp_rep_scale_host[z]); a[z] = fabsf(c->exp_rep_prefactor_host[z]); }   // :113-117
    float *dr = nullptr, *da = nullptr;
    APAX_TRY(upload(&dr, r.data(), size_t(S)));
    APAX_TRY(upload(&da, a.data(), size_t(S)));
    cc.exp_on = 1; cc.exp_rmax = c->exp_r_max; cc.exp_r = dr; cc.exp_a = da;
  }
  if (p->corr.exp_r) cudaFree(const_cast<float*>(p->corr.exp_r));
  if (p->corr.exp_a) cudaFree(const_cast<float*>(p->corr.exp_a));
  p->corr = cc;
  return APAX_OK;
}
