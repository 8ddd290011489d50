// Shallow-ensemble statistics on the device.  Replaces the tail of ShallowEnsembleModel.__call__
// (apax/nn/models.py:221-263): mean / unbiased variance over the readout heads of the per-structure energies and of the
// per-atom forces (fp64, divisor 1/(n_ens-1)), and the [n_atoms][3][n_ens] layout of `forces_ensemble`.
#include "common.cuh"

namespace apax {
namespace {

// thread per force component: F [n_ens][n_atoms*3] -> mean, sqrt(variance), transposed ensemble
__global__ void __launch_bounds__(256) k_ens_force_stats(const double* __restrict__ F, int64_t n3, int n_ens,
                                                         double* __restrict__ f_mean, double* __restrict__ f_unc,
                                                         double* __restrict__ f_ens) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  double v[16];
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    v[k] = k < n_ens ? F[int64_t(k) * n3 + i] : 0.0;
    s += v[k];
  }
  const double mean = s / double(n_ens);                    // jnp.mean(forces_ens, axis=0)
  double var = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < n_ens) var += (v[k] - mean) * (v[k] - mean);    // fp64_sum((f - mean)^2, axis=0)
  f_mean[i] = mean;
  f_unc[i] = sqrt(var * (1.0 / double(n_ens - 1)));
  if (f_ens)
    for (int k = 0; k < n_ens; ++k) f_ens[i * n_ens + k] = v[k];   // jnp.transpose(forces_ens, (1, 2, 0))
}

__global__ void __launch_bounds__(256) k_ens_energy_stats(const double* __restrict__ E, int64_t n_structs, int n_ens,
                                                          double* __restrict__ e_mean, double* __restrict__ e_unc) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_structs) return;
  double sum = 0.0;
  for (int k = 0; k < n_ens; ++k) sum += E[s * n_ens + k];
  const double mean = sum / double(n_ens);
  double var = 0.0;
  for (int k = 0; k < n_ens; ++k) var += (E[s * n_ens + k] - mean) * (E[s * n_ens + k] - mean);
  e_mean[s] = mean;
  e_unc[s] = sqrt(var * (1.0 / double(n_ens - 1)));
}

}  // namespace
}  // namespace apax

using namespace apax;

extern "C" int apax_ensemble_stats(apax_stream_t stream_, const double* energy_ens, const double* forces_ens, int64_t n_structs,
                                   int64_t n_atoms, int32_t n_ens, double* e_mean, double* e_unc, double* f_mean, double* f_unc,
                                   double* f_ens_t) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!energy_ens || !forces_ens || !e_mean || !e_unc || !f_mean || !f_unc || n_structs < 1 || n_atoms < 1) return APAX_ERR_INVALID_ARG;
  if (n_ens < 2 || n_ens > 16) return APAX_ERR_UNSUPPORTED;
  APAX_LAUNCH(k_ens_energy_stats, dim3((unsigned)cdiv(n_structs, 256)), dim3(256), 0, s, energy_ens, n_structs, n_ens, e_mean, e_unc);
  APAX_LAUNCH(k_ens_force_stats, dim3((unsigned)cdiv(n_atoms * 3, 256)), dim3(256), 0, s, forces_ens, n_atoms * 3, n_ens, f_mean,
              f_unc, f_ens_t);
  return APAX_OK;
}
