// NVT Nose-Hoover-chain molecular dynamics around the GMNN force kernels, entirely on the device.
//
// Replaces the body of the jitted fori_loop of apax/md/simulate.py:313-354 for the ensemble apax builds at
// simulate.py:100-119, i.e. jax_md's nvt_nose_hoover (apax/utils/jax_md_reduced/simulate.py:536-639):
//   chain half step (simulate.py:436-487) -> velocity Verlet (:216-231) with the fractional, wrapped shift of
//   periodic_general (space.py:284-309) -> kinetic energy -> chain half step.
// The chain is a handful of scalars: one thread integrates it on the device (no host round trip inside the
// MD loop -- the reference hops to the host every step through io_callback, simulate.py:343); the O(N) parts
// are elementwise kernels and a fixed-shape fp64 reduction.  dtypes follow the reference: fp64 state, time
// steps and chain masses rounded to fp32 where it casts with f32().
#include <math.h>
#include <stddef.h>

#include "common.cuh"

namespace apax {

constexpr int kMaxChain = 16;

struct MdChainState {       // lives in device memory (apax_md_nvt_state_bytes)
  double xi[kMaxChain];
  double p_xi[kMaxChain];
  double Q[kMaxChain];      // fp32-rounded values
  double KE;
  double scale;             // momentum scale of the last chain half step
  double inv_box[9];
  double scale_prev;        // 1, or the scale of the post-force half when both halves of a step boundary ran in one launch
  double partial[256];
  int dof;
  int chain_length;
};

namespace {

__constant__ double kSY3[3] = {0.828981543588751, -0.657963087177502, 0.828981543588751};
__constant__ double kSY5[5] = {0.2967324292201065, 0.2967324292201065, -0.186929716880426, 0.2967324292201065,
                               0.2967324292201065};
__constant__ double kSY7[7] = {0.784513610477560, 0.235573213359357, -1.17767998417887, 1.31518632068391,
                               -1.17767998417887, 0.235573213359357, 0.784513610477560};

__device__ __forceinline__ double f32r(double x) { return double(float(x)); }

// exp(x) of the chain's arguments (dt/8 p_xi/Q: 1e-3 at the default tau = 100 dt).  One thread runs ~50 of these per half
// step back to back and a dependent fp64 operation costs ~40 cycles here, so the DEPTH of the evaluation is the kernel:
// Taylor polynomials in Estrin form -- degree 5 below 2^-8 (truncation x^6/720 < 5e-18, depth 3), degree 9 below 2^-4
// (x^10/10! < 3e-19, depth 4) -- and libm's exp beyond that.
__device__ __forceinline__ double exp_chain(double x) {
  const double ax = fabs(x);
  if (ax > 0.0625) return exp(x);
  const double x2 = x * x;
  const double a01 = 1.0 + x, a23 = fma(x, 1.0 / 6.0, 0.5), a45 = fma(x, 1.0 / 120.0, 1.0 / 24.0);
  const double x4 = x2 * x2;
  const double lo = fma(x2, a23, a01);
  if (ax < 0.00390625) return fma(x4, a45, lo);
  const double a67 = fma(x, 1.0 / 5040.0, 1.0 / 720.0), a89 = fma(x, 1.0 / 362880.0, 1.0 / 40320.0);
  const double hi = fma(x2, a67, a45);                 // a45 + x^2 a67
  const double hi2 = fma(x4, a89, hi);                 // ... + x^4 a89
  return fma(x4, hi2, lo);
}

// The chain of one launch in registers: CL > 0 fixes the chain length at compile time (5 is apax's and jax_md's default), so
// that every loop below unrolls and xi / p_xi / 1/Q never touch local or shared memory; CL == 0 is the run-time-length
// fallback (2..16).
template <int CL>
struct ChainRegs {
  static constexpr int kCap = CL ? CL : kMaxChain;
  double xi[kCap], p[kCap], iq[kCap];
  double KE, dofkT, kT;
  int len;
  __device__ __forceinline__ int M() const { return (CL ? CL : len) - 1; }
};

// substep_fn (simulate.py:436-473).  One thread integrates the chain, so the serial fp64 latencies are what this costs: the
// chain masses are constant within a call, their reciprocals `iq` are taken once per launch (x / Q -> x * (1/Q): one ulp).
// The exponentials of the second sweep depend on the momenta BEFORE the sweep only, so they are independent of its carry.
template <int CL>
__device__ __forceinline__ double chain_substep(ChainRegs<CL>& c, double delta) {
  constexpr int kCap = ChainRegs<CL>::kCap;
  const int M = c.M();
  const double kT = c.kT;
  const double delta_2 = f32r(delta / 2.0), delta_4 = f32r(delta_2 / 2.0), delta_8 = f32r(delta_4 / 2.0);
  double p_old[kCap];
  double G = c.p[M - 1] * c.p[M - 1] * c.iq[M - 1] - kT;
  c.p[M] += delta_4 * G;
#pragma unroll
  for (int m = 0; m < kCap; ++m) p_old[m] = c.p[m];
  double carry = c.p[M];
#pragma unroll
  for (int m = kCap - 2; m >= 1; --m) {
    if (m <= M - 1) {
      G = p_old[m - 1] * p_old[m - 1] * c.iq[m - 1] - kT;
      const double s = exp_chain(carry * (-delta_8 * c.iq[m + 1]));
      carry = s * (s * p_old[m] + delta_4 * G);
      c.p[m] = carry;
    }
  }
  G = 2.0 * c.KE - c.dofkT;
  double s = exp_chain(c.p[1] * (-delta_8 * c.iq[1]));
  c.p[0] = s * (s * c.p[0] + delta_4 * G);
  s = exp_chain(c.p[0] * (-delta_2 * c.iq[0]));
  c.KE *= s * s;
  const double p_scale = s;
#pragma unroll
  for (int m = 0; m < kCap; ++m)
    if (m <= M) c.xi[m] += delta_2 * c.p[m] * c.iq[m];
  G = 2.0 * c.KE - c.dofkT;
#pragma unroll
  for (int m = 0; m < kCap; ++m) p_old[m] = c.p[m];
#pragma unroll
  for (int m = 0; m < kCap - 1; ++m) {
    if (m < M) {
      const double sc = exp_chain(p_old[m + 1] * (-delta_8 * c.iq[m + 1]));
      const double upd = sc * (sc * p_old[m] + delta_4 * G);
      G = upd * upd * c.iq[m] - kT;
      c.p[m] = upd;
    }
  }
  c.p[M] += delta_4 * G;
  return p_scale;
}

template <int CL>
__device__ __forceinline__ double chain_half(ChainRegs<CL>& c, double dt, int chain_steps, int sy_steps) {
  if (chain_steps == 1 && sy_steps == 1) return chain_substep<CL>(c, dt);
  double total = 1.0;
  const double delta = f32r(dt / chain_steps);
  const double* ws = sy_steps == 3 ? kSY3 : (sy_steps == 5 ? kSY5 : kSY7);
  for (int i = 0; i < chain_steps * sy_steps; ++i) {
    const double w = sy_steps == 1 ? 1.0 : ws[i % sy_steps];
    total *= chain_substep<CL>(c, f32r(delta * w));
  }
  return total;
}

// update_chain_mass_fn + half_step_chain_fn (simulate.py:475-495) on the state staged in shared memory `sc`; `both` runs the
// post-force half of one step and the pre-force half of the next (the second depends on the chain only: the first already
// rescaled KE) and leaves the two momentum factors in scale_prev / scale.
template <int CL>
__device__ void chain_run(MdChainState& sc, double dt, double kT, double tau, int chain_steps, int sy_steps, int both) {
  constexpr int kCap = ChainRegs<CL>::kCap;
  ChainRegs<CL> c;
  c.len = sc.chain_length;
  c.KE = sc.KE;
  c.kT = kT;
  c.dofkT = sc.dof * kT;
  const float q = float(kT) * (float(tau) * float(tau));
#pragma unroll
  for (int m = 0; m < kCap; ++m) {
    const double Qm = double(m == 0 ? q * float(sc.dof) : q);
    if (m < c.len) sc.Q[m] = Qm;
    c.iq[m] = 1.0 / Qm;
    c.xi[m] = sc.xi[m];
    c.p[m] = sc.p_xi[m];
  }
  sc.scale_prev = both ? chain_half<CL>(c, dt, chain_steps, sy_steps) : 1.0;
  sc.scale = chain_half<CL>(c, dt, chain_steps, sy_steps);
#pragma unroll
  for (int m = 0; m < kCap; ++m) {
    if (m < c.len) {
      sc.xi[m] = c.xi[m];
      sc.p_xi[m] = c.p[m];
    }
  }
  sc.KE = c.KE;
}

// The chain is a few dozen scalars integrated by ONE thread; the state's head (everything before the 2 KB of reduction
// partials) travels through shared memory, loaded and stored by the whole warp -- a single thread copying the full struct
// from and to global memory cost 37 us per launch, 13 % of a 12,288-atom MD step (tools/md_step_profile.py).
constexpr int kChainHeadDoubles = int(offsetof(MdChainState, partial) / sizeof(double));
// n_partial > 0: the kinetic energy is first summed from the per-block partials of k_md_kick_ke (fixed order: lane-strided
// partial sums joined by a butterfly), which saves the separate k_md_ke_final launch of the post-force half.
__global__ void __launch_bounds__(32) k_md_chain_half(MdChainState* st, double dt, double kT, double tau, int chain_steps, int sy_steps,
                                                      const double* __restrict__ box, int n_partial = 0, int both = 0) {
  if (blockIdx.x != 0) return;
  __shared__ MdChainState sc;
  double* sd = reinterpret_cast<double*>(&sc);
  const double* gd = reinterpret_cast<const double*>(st);
  for (int i = threadIdx.x; i < kChainHeadDoubles; i += 32) sd[i] = gd[i];
  if (threadIdx.x == 0) { sc.dof = st->dof; sc.chain_length = st->chain_length; }
  if (n_partial > 0) {
    double ke = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) ke += st->partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ke += __shfl_xor_sync(0xffffffffu, ke, o);
    __syncwarp();
    if (threadIdx.x == 0) sc.KE = ke;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    MdChainState& c = sc;
    if (sc.chain_length == 5) chain_run<5>(sc, dt, kT, tau, chain_steps, sy_steps, both);
    else chain_run<0>(sc, dt, kT, tau, chain_steps, sy_steps, both);
    if (box) {  // inverse of the (possibly time-dependent) box for the fractional position update
      const double* a = box;
      const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
      const double id = 1.0 / det;
      c.inv_box[0] = (a[4] * a[8] - a[5] * a[7]) * id; c.inv_box[1] = (a[2] * a[7] - a[1] * a[8]) * id; c.inv_box[2] = (a[1] * a[5] - a[2] * a[4]) * id;
      c.inv_box[3] = (a[5] * a[6] - a[3] * a[8]) * id; c.inv_box[4] = (a[0] * a[8] - a[2] * a[6]) * id; c.inv_box[5] = (a[2] * a[3] - a[0] * a[5]) * id;
      c.inv_box[6] = (a[3] * a[7] - a[4] * a[6]) * id; c.inv_box[7] = (a[1] * a[6] - a[0] * a[7]) * id; c.inv_box[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    }
  }
  __syncwarp();
  double* go = reinterpret_cast<double*>(st);
  for (int i = threadIdx.x; i < kChainHeadDoubles; i += 32) go[i] = sd[i];
}

// P <- (P*scale_prev)*scale + dt/2 F ; R <- mod(R + inv_box.(dt P / m), 1)      (momentum_step, position_step, periodic_shift)
__global__ void __launch_bounds__(256) k_md_kick_drift(double* __restrict__ R, double* __restrict__ P,
                                                       const double* __restrict__ F, const double* __restrict__ mass,
                                                       const MdChainState* __restrict__ st, double dt, double dt_2,
                                                       int64_t N) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= N) return;
  const double s = st->scale, sp = st->scale_prev;   // sp == 1 unless both chain halves ran in one launch
  double p[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double pk = P[3 * i + k] * sp;               // rounded like k_md_scale's store
    p[k] = pk * s + dt_2 * F[3 * i + k];
    P[3 * i + k] = p[k];
    d[k] = dt * p[k] / mass[i];
  }
  const double* ib = st->inv_box;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double t = R[3 * i + k] + (ib[3 * k] * d[0] + ib[3 * k + 1] * d[1] + ib[3 * k + 2] * d[2]);
    double m = fmod(t, 1.0);
    if (m < 0.0) m += 1.0;   // jnp.mod
    R[3 * i + k] = m;
  }
}

// P <- P + dt/2 F, and the per-block partial sums of 0.5 p^2/m
__global__ void __launch_bounds__(256) k_md_kick_ke(double* __restrict__ P, const double* __restrict__ F,
                                                    const double* __restrict__ mass, MdChainState* __restrict__ st,
                                                    double dt_2, int64_t N, int apply_kick) {
  __shared__ double sm[256];
  const int64_t chunk = (N + gridDim.x - 1) / gridDim.x;
  const int64_t b = int64_t(blockIdx.x) * chunk, e = min(N, b + chunk);
  double acc = 0.0;
  for (int64_t i = b + threadIdx.x; i < e; i += 256) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double p = P[3 * i + k];
      if (apply_kick) {
        p += dt_2 * F[3 * i + k];
        P[3 * i + k] = p;
      }
      s += p * p;
    }
    acc += 0.5 * s / mass[i];
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) st->partial[blockIdx.x] = sm[0];
}

__global__ void k_md_ke_final(MdChainState* st, int n_blocks, double* ke_out = nullptr) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += st->partial[i];
    if (ke_out) *ke_out = s; else st->KE = s;
  }
}

__global__ void k_md_set_global(MdChainState* st, const double* __restrict__ ke_global, int dof) {
  if (threadIdx.x == 0) {
    st->KE = *ke_global;
    st->dof = dof;
  }
}

__global__ void __launch_bounds__(256) k_md_scale(double* __restrict__ P, const MdChainState* __restrict__ st, int64_t n3) {
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n3) P[i] *= st->scale;
}

__global__ void k_md_chain_init(MdChainState* st, int chain_length, int dof) {
  if (threadIdx.x != 0) return;
  for (int m = 0; m < kMaxChain; ++m) { st->xi[m] = 0.0; st->p_xi[m] = 0.0; st->Q[m] = 1.0; }
  st->chain_length = chain_length;
  st->dof = dof;
  st->scale = 1.0;
  st->scale_prev = 1.0;
}

// blocks of the kinetic-energy reduction: one atom per thread up to the 256 partials the state holds
int ke_blocks(int64_t N) { return int(N <= 256 ? 1 : (cdiv(N, 256) > 256 ? 256 : cdiv(N, 256))); }

}  // namespace
}  // namespace apax

using namespace apax;

extern "C" size_t apax_md_nvt_state_bytes(void) { return sizeof(MdChainState); }

static int md_cfg_ok(const apax_md_nvt_config* c) {
  if (!c || c->chain_length < 2 || c->chain_length > kMaxChain || c->chain_steps < 1) return APAX_ERR_INVALID_ARG;
  if (c->sy_steps != 1 && c->sy_steps != 3 && c->sy_steps != 5 && c->sy_steps != 7) return APAX_ERR_INVALID_ARG;
  if (!(c->dt > 0.0) || !(c->kT > 0.0) || !(c->tau > 0.0)) return APAX_ERR_INVALID_ARG;
  return APAX_OK;
}

extern "C" int apax_md_nvt_init(apax_stream_t stream, const apax_md_nvt_config* cfg, const double* P, const double* mass,
                                int64_t n_atoms, void* state) {
  APAX_TRY(md_cfg_ok(cfg));
  if (!P || !mass || !state || n_atoms < 1) return APAX_ERR_INVALID_ARG;
  MdChainState* st = static_cast<MdChainState*>(state);
  APAX_LAUNCH(k_md_chain_init, dim3(1), dim3(32), 0, stream, st, cfg->chain_length, int(3 * n_atoms));
  const int nb = ke_blocks(n_atoms);
  APAX_LAUNCH(k_md_kick_ke, dim3(nb), dim3(256), 0, stream, const_cast<double*>(P), (const double*)nullptr, mass, st, 0.0,
              n_atoms, 0);
  APAX_LAUNCH(k_md_ke_final, dim3(1), dim3(32), 0, stream, st, nb);
  return APAX_OK;
}

extern "C" int apax_md_nvt_pre_force(apax_stream_t stream, const apax_md_nvt_config* cfg, double* R_frac, double* P,
                                     const double* F, const double* mass, const double* box, int64_t n_atoms,
                                     void* state) {
  APAX_TRY(md_cfg_ok(cfg));
  if (!R_frac || !P || !F || !mass || !box || !state) return APAX_ERR_INVALID_ARG;
  MdChainState* st = static_cast<MdChainState*>(state);
  const double dt = double(float(cfg->dt)), dt_2 = double(float(dt / 2));   // f32(dt), f32(dt / 2) (simulate.py:219-220)
  APAX_LAUNCH(k_md_chain_half, dim3(1), dim3(32), 0, stream, st, dt, cfg->kT, double(float(cfg->tau)), cfg->chain_steps,
              cfg->sy_steps, box);
  APAX_LAUNCH(k_md_kick_drift, dim3((unsigned)cdiv(n_atoms, 256)), dim3(256), 0, stream, R_frac, P, F, mass, st, dt, dt_2,
              n_atoms);
  return APAX_OK;
}

extern "C" int apax_md_nvt_post_force(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, const double* F,
                                      const double* mass, int64_t n_atoms, void* state) {
  APAX_TRY(md_cfg_ok(cfg));
  if (!P || !F || !mass || !state) return APAX_ERR_INVALID_ARG;
  MdChainState* st = static_cast<MdChainState*>(state);
  const double dt = double(float(cfg->dt)), dt_2 = double(float(dt / 2));
  const int nb = ke_blocks(n_atoms);
  APAX_LAUNCH(k_md_kick_ke, dim3(nb), dim3(256), 0, stream, P, F, mass, st, dt_2, n_atoms, 1);
  APAX_LAUNCH(k_md_chain_half, dim3(1), dim3(32), 0, stream, st, dt, cfg->kT, double(float(cfg->tau)), cfg->chain_steps,
              cfg->sy_steps, (const double*)nullptr, nb);
  APAX_LAUNCH(k_md_scale, dim3((unsigned)cdiv(3 * n_atoms, 256)), dim3(256), 0, stream, P, st, 3 * n_atoms);
  return APAX_OK;
}

extern "C" int apax_md_nvt_kick_ke(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, const double* F,
                                   const double* mass, int64_t n_atoms, int32_t apply_kick, void* state, double* ke_local) {
  APAX_TRY(md_cfg_ok(cfg));
  if (!P || !mass || !state || !ke_local || (apply_kick && !F) || n_atoms < 0) return APAX_ERR_INVALID_ARG;
  MdChainState* st = static_cast<MdChainState*>(state);
  const double dt = double(float(cfg->dt)), dt_2 = double(float(dt / 2));
  const int nb = ke_blocks(n_atoms);
  APAX_LAUNCH(k_md_kick_ke, dim3(nb), dim3(256), 0, stream, P, F, mass, st, dt_2, n_atoms, apply_kick ? 1 : 0);
  APAX_LAUNCH(k_md_ke_final, dim3(1), dim3(32), 0, stream, st, nb, ke_local);
  return APAX_OK;
}

extern "C" int apax_md_nvt_set_global(apax_stream_t stream, void* state, const double* ke_global, int64_t n_atoms_global) {
  if (!state || !ke_global || n_atoms_global < 1 || 3 * n_atoms_global > 2147483647LL) return APAX_ERR_INVALID_ARG;
  APAX_LAUNCH(k_md_set_global, dim3(1), dim3(32), 0, stream, static_cast<MdChainState*>(state), ke_global, int(3 * n_atoms_global));
  return APAX_OK;
}

extern "C" int apax_md_nvt_chain_finish(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, int64_t n_atoms, void* state,
                                        const double* ke_global, int64_t n_atoms_global) {
  APAX_TRY(md_cfg_ok(cfg));
  if (!P || !state || n_atoms < 0) return APAX_ERR_INVALID_ARG;
  APAX_TRY(apax_md_nvt_set_global(stream, state, ke_global, n_atoms_global));
  MdChainState* st = static_cast<MdChainState*>(state);
  const double dt = double(float(cfg->dt));
  APAX_LAUNCH(k_md_chain_half, dim3(1), dim3(32), 0, stream, st, dt, cfg->kT, double(float(cfg->tau)), cfg->chain_steps,
              cfg->sy_steps, (const double*)nullptr);
  if (n_atoms > 0) APAX_LAUNCH(k_md_scale, dim3((unsigned)cdiv(3 * n_atoms, 256)), dim3(256), 0, stream, P, st, 3 * n_atoms);
  return APAX_OK;
}

// Several MD steps between two neighbour-list decisions in ONE call: the inner loop of apax's MD driver
// (apax/md/simulate.py:313-354; jax.lax.fori_loop over `apply_fn`) without returning to the host language between steps.
// Each step is apax_md_nvt_pre_force -> apax_gm_compute (MINIMAGE, the list prepared in `workspace`) -> apax_md_nvt_post_force,
// all asynchronous on `stream`; at 10^4 atoms a step is ~26 short kernels, so the host-side cost per launch is what counts.
extern "C" int apax_md_nvt_run(apax_stream_t stream, const apax_md_nvt_config* cfg, const apax_gm_params* params, double* R_frac,
                               double* P, double* F, const double* mass, const int32_t* Z, const double* box, int64_t n_atoms,
                               int64_t n_pairs, void* state, double* energy, void* workspace, size_t workspace_bytes,
                               int32_t n_steps) {
  if (n_steps < 0) return APAX_ERR_INVALID_ARG;
  if (n_steps == 0) return APAX_OK;
  APAX_TRY(md_cfg_ok(cfg));
  if (!R_frac || !P || !F || !mass || !box || !state) return APAX_ERR_INVALID_ARG;
  MdChainState* st = static_cast<MdChainState*>(state);
  const double dt = double(float(cfg->dt)), dt_2 = double(float(dt / 2));
  const int nb = ke_blocks(n_atoms);
  APAX_TRY(apax_md_nvt_pre_force(stream, cfg, R_frac, P, F, mass, box, n_atoms, state));
  for (int32_t i = 0; i < n_steps; ++i) {
    APAX_TRY(apax_gm_compute(stream, params, R_frac, Z, box, nullptr, n_atoms, n_pairs, APAX_DISP_MINIMAGE, energy, F, workspace,
                             workspace_bytes));
    if (i + 1 == n_steps) {
      APAX_TRY(apax_md_nvt_post_force(stream, cfg, P, F, mass, n_atoms, state));
      break;
    }
    // step boundary inside the call: kick -> both chain halves in one launch -> rescale + kick + drift in one launch
    // (the same arithmetic as post_force + pre_force: k_md_scale's multiplication moves into k_md_kick_drift)
    APAX_LAUNCH(k_md_kick_ke, dim3(nb), dim3(256), 0, stream, P, F, mass, st, dt_2, n_atoms, 1);
    APAX_LAUNCH(k_md_chain_half, dim3(1), dim3(32), 0, stream, st, dt, cfg->kT, double(float(cfg->tau)), cfg->chain_steps,
                cfg->sy_steps, box, nb, 1);
    APAX_LAUNCH(k_md_kick_drift, dim3((unsigned)cdiv(n_atoms, 256)), dim3(256), 0, stream, R_frac, P, F, mass, st, dt, dt_2,
                n_atoms);
  }
  return APAX_OK;
}
