// Host-buffer context: the per-call flow of ASECalculator.calculate (apax/md/ase_calc.py:357-393)
// behind one C call -- H2D of positions, neighbour list (min-image with skin when
// min cell height / 2 > r_max, image list otherwise, :501-522), energy+forces, D2H of the results.
#include <math.h>
#include <string.h>

#include "gm_internal.cuh"
#include "nl_internal.cuh"

using namespace apax;

struct apax_ctx {
  const apax_gm_params_impl* params;
  int64_t max_atoms, max_pairs;
  cudaStream_t stream;
  // device
  double* pos;       // [N][3] Cartesian
  double* frac;      // [N][3]
  double* ref_frac;  // [N][3] positions at the last list build
  int32_t* Z;
  double* mats;      // cell[9] | box = cell^T [9] | inverse box [9] | identity [9] | 2 x identity [9] | stress out [9]
  int32_t* idx;      // [2][cap] (only when a COO copy is wanted; unused on the fast path)
  double* offsets;   // [cap][3]
  int32_t* counters; // n_found, rebuild flag
  uint8_t* overflow;
  double* energy;
  double* forces;
  void* gm_ws;
  size_t gm_ws_bytes;
  void* nl_ws;
  size_t nl_ws_bytes;
  // pinned host staging
  double* h_pos;
  int32_t* h_Z;
  double* h_mats;
  int32_t* h_counters;
  uint8_t* h_overflow;
  double* h_out;  // energy[n_out] then forces
  int32_t* list_Z; // host copy of the atomic numbers the list (and its compact species table) was prepared for
  // state
  int64_t list_atoms;
  int list_minimage;
  double list_cutoff;
  double list_cell[9];
  int64_t n_pairs;
  // CUDA graph of the image-list path (small cells are launch-latency bound: ~45 dependent launches on 96 atoms)
  cudaGraphExec_t graph_exec;
  int64_t graph_N;
  int graph_forces, graph_tag;
  int64_t graph_launches;   // kernels inside the graph: added to the library's launch counter on every replay
  int64_t image_shape_N;
  int image_shape_calls;
  int evaluated;     // the workspace holds the per-pair values of a completed energy+forces evaluation (for apax_ctx_stress)
  double* scratch_ef; // [1 + 3 max_atoms] second evaluation of the offsets-branch stress (allocated on first use)
};

namespace {
__global__ void __launch_bounds__(256) k_to_frac(const double* __restrict__ pos, const double* __restrict__ binv,
                                                 int64_t N, double* __restrict__ frac) {
  // positions = space.transform(inv_box, positions) (ase_calc.py:534-535): f = inv(cell^T) . r
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= N) return;
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  frac[3 * i + 0] = binv[0] * x + binv[1] * y + binv[2] * z;
  frac[3 * i + 1] = binv[3] * x + binv[4] * y + binv[5] * z;
  frac[3 * i + 2] = binv[6] * x + binv[7] * y + binv[8] * z;
}

void host_invert3(const double* a, double* inv) {
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                     a[2] * (a[3] * a[7] - a[4] * a[6]);
  const double id = 1.0 / det;
  inv[0] = (a[4] * a[8] - a[5] * a[7]) * id; inv[1] = (a[2] * a[7] - a[1] * a[8]) * id; inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  inv[3] = (a[5] * a[6] - a[3] * a[8]) * id; inv[4] = (a[0] * a[8] - a[2] * a[6]) * id; inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = (a[3] * a[7] - a[4] * a[6]) * id; inv[7] = (a[1] * a[6] - a[0] * a[7]) * id; inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// neighbor_calculable_with_jax (ase_calc.py:501-522): min plane distance / 2 > r_max
bool minimage_ok(const double* cell, double r_max) {
  double minh = 1e300;
  for (int k = 0; k < 3; ++k) {
    const double* a = cell + 3 * ((k + 1) % 3);
    const double* b = cell + 3 * ((k + 2) % 3);
    const double* c = cell + 3 * k;
    const double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double h = fabs(n[0] * c[0] + n[1] * c[1] + n[2] * c[2]) / nn;
    if (h < minh) minh = h;
  }
  return minh / 2.0 > r_max;
}
}  // namespace

extern "C" int apax_ctx_create(const apax_gm_params* params, int64_t max_atoms, int64_t max_pairs, apax_ctx** out) {
  if (!params || !out || max_atoms < 1 || max_pairs < 1) return APAX_ERR_INVALID_ARG;
  const apax_gm_params_impl* p = reinterpret_cast<const apax_gm_params_impl*>(params);
  apax_ctx* c = new apax_ctx();
  memset(c, 0, sizeof(*c));
  c->params = p;
  c->max_atoms = max_atoms;
  c->max_pairs = max_pairs;
  const int n_out = p->cfg.n_out;
  APAX_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  APAX_CUDA_TRY(cudaMalloc(&c->pos, sizeof(double) * 3 * max_atoms));
  APAX_CUDA_TRY(cudaMalloc(&c->frac, sizeof(double) * 3 * max_atoms));
  APAX_CUDA_TRY(cudaMalloc(&c->ref_frac, sizeof(double) * 3 * max_atoms));
  APAX_CUDA_TRY(cudaMalloc(&c->Z, sizeof(int32_t) * max_atoms));
  APAX_CUDA_TRY(cudaMalloc(&c->mats, sizeof(double) * 54));
  APAX_CUDA_TRY(cudaMalloc(&c->offsets, sizeof(double) * 3 * max_pairs));
  APAX_CUDA_TRY(cudaMalloc(&c->counters, sizeof(int32_t) * 4));
  APAX_CUDA_TRY(cudaMalloc(&c->overflow, 4));
  APAX_CUDA_TRY(cudaMalloc(&c->energy, sizeof(double) * n_out));
  APAX_CUDA_TRY(cudaMalloc(&c->forces, sizeof(double) * 3 * max_atoms * n_out));
  c->gm_ws_bytes = apax_gm_workspace_bytes(max_atoms, max_pairs, n_out);
  c->nl_ws_bytes = apax_nl_workspace_bytes(max_atoms, max_pairs);
  APAX_CUDA_TRY(cudaMalloc(&c->gm_ws, c->gm_ws_bytes));
  APAX_CUDA_TRY(cudaMalloc(&c->nl_ws, c->nl_ws_bytes));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_pos, sizeof(double) * 3 * max_atoms));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_Z, sizeof(int32_t) * max_atoms));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_mats, sizeof(double) * 54));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_counters, sizeof(int32_t) * 4));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_overflow, 4));
  APAX_CUDA_TRY(cudaMallocHost(&c->h_out, sizeof(double) * (n_out + 3 * max_atoms * n_out)));
  c->list_Z = new int32_t[max_atoms];
  c->list_atoms = -1;
  *out = c;
  return APAX_OK;
}

extern "C" void apax_ctx_destroy(apax_ctx* c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  cudaFree(c->pos); cudaFree(c->frac); cudaFree(c->ref_frac); cudaFree(c->Z); cudaFree(c->mats);
  cudaFree(c->offsets); cudaFree(c->counters); cudaFree(c->overflow); cudaFree(c->energy); cudaFree(c->forces);
  cudaFree(c->gm_ws); cudaFree(c->nl_ws); cudaFree(c->scratch_ef);
  cudaFreeHost(c->h_pos); cudaFreeHost(c->h_Z); cudaFreeHost(c->h_mats); cudaFreeHost(c->h_counters);
  cudaFreeHost(c->h_overflow); cudaFreeHost(c->h_out);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  cudaStreamDestroy(c->stream);
  delete[] c->list_Z;
  delete c;
}

// Everything the image-list (vesin-replacement) path does between the host staging buffers, asynchronous on the context's
// stream and free of host synchronisation, so that it can be captured once in a CUDA graph and replayed: H2D of positions /
// numbers / matrices, cell-list image search, offsets, list preparation, E+F, D2H of the counters and the results.  The
// overflow flag is checked by the caller AFTER the results have arrived (every kernel respects the capacity).
static int enqueue_image_path(apax_ctx* c, int64_t N, bool want_forces) {
  const apax_gm_params_impl* p = c->params;
  const int n_out = p->cfg.n_out;
  const int n_f = (p->mean_forces && n_out > 1) ? 1 : n_out;
  const int64_t cap = c->max_pairs;
  cudaStream_t s = c->stream;
  APAX_CUDA_TRY(cudaMemcpyAsync(c->pos, c->h_pos, sizeof(double) * 3 * N, cudaMemcpyHostToDevice, s));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->Z, c->h_Z, sizeof(int32_t) * N, cudaMemcpyHostToDevice, s));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->mats, c->h_mats, sizeof(double) * 45, cudaMemcpyHostToDevice, s));
  GmWorkspace gw;
  NlWorkspace nw;
  carve_gm_workspace(c->gm_ws, N, cap, n_out, 128, &gw);
  carve_nl_workspace(c->nl_ws, N, cap, &nw);
  APAX_CUDA_TRY(cudaMemsetAsync(c->overflow, 0, 4, s));
  APAX_TRY(nl_build_images_csr(s, c->pos, c->mats, 1, N, double(p->cfg.r_max), cap, nw, nullptr, nullptr, nullptr, c->counters,
                               c->overflow));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->h_overflow, c->overflow, 1, cudaMemcpyDeviceToHost, s));
  APAX_TRY(nl_emit_offsets(s, nw, N, c->offsets));
  APAX_TRY(gm_prepare_from_csr(s, p, gw, c->Z, nw.rowptr_c, nw.nbr, /*symmetric_sorted=*/false));
  APAX_TRY(gm_compute_device(s, p, gw, c->pos, c->Z, c->mats + 27, c->offsets, APAX_DISP_OFFSETS, c->energy,
                             want_forces ? c->forces : nullptr));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->h_out, c->energy, sizeof(double) * n_out, cudaMemcpyDeviceToHost, s));
  if (want_forces)
    APAX_CUDA_TRY(cudaMemcpyAsync(c->h_out + n_out, c->forces, sizeof(double) * 3 * N * n_f, cudaMemcpyDeviceToHost, s));
  return APAX_OK;
}

extern "C" int apax_ctx_calculate(apax_ctx* c, const double* positions_host, const int32_t* numbers_host,
                                  const double* cell_host, int64_t n_atoms, double dr_threshold, int32_t rebuild,
                                  double* energy_host, double* forces_host, int64_t* n_pairs_host) {
  if (!c || !positions_host || !numbers_host || !cell_host || !energy_host || n_atoms < 1) return APAX_ERR_INVALID_ARG;
  if (n_atoms > c->max_atoms) return APAX_ERR_OVERFLOW;
  const apax_gm_params_impl* p = c->params;
  const int n_out = p->cfg.n_out;
  const int n_f = (p->mean_forces && n_out > 1) ? 1 : n_out;   // rows of the force output (apax_gm_params_set_option)
  const int64_t N = n_atoms, cap = c->max_pairs;
  cudaStream_t s = c->stream;
  c->evaluated = 0;
  double det = cell_host[0] * (cell_host[4] * cell_host[8] - cell_host[5] * cell_host[7]) -
               cell_host[1] * (cell_host[3] * cell_host[8] - cell_host[5] * cell_host[6]) +
               cell_host[2] * (cell_host[3] * cell_host[7] - cell_host[4] * cell_host[6]);
  if (!(fabs(det) > 1e-12)) return APAX_ERR_UNSUPPORTED;  // gas phase: use apax_gm_energy_forces with a list

  // ---- H2D: positions, numbers, cell / box / inverse / identity.  Page-locked caller buffers are copied from
  // directly; pageable ones go through the context's pinned staging area.
  auto is_pinned = [](const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool pos_pinned = is_pinned(positions_host), z_pinned = is_pinned(numbers_host);
  if (!pos_pinned) memcpy(c->h_pos, positions_host, sizeof(double) * 3 * N);
  if (!z_pinned) memcpy(c->h_Z, numbers_host, sizeof(int32_t) * N);
  double* m = c->h_mats;
  for (int i = 0; i < 9; ++i) m[i] = cell_host[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[9 + 3 * i + j] = cell_host[3 * j + i];  // box = cell.T (ase_calc.py:42-43)
  host_invert3(m + 9, m + 18);
  for (int i = 0; i < 9; ++i) m[27 + i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 9; ++i) m[36 + i] = (i % 4 == 0) ? 2.0 : 0.0;
  const bool minimage = minimage_ok(cell_host, p->cfg.r_max);
  if (!minimage && !g_profile_on) {
    // ---- image-list path (recomputed on every call, apax/md/ase_calc.py:386-387), replayed from a CUDA graph: the first call
    // at a shape runs eagerly (one-time initialisations), the second is captured, later ones are ONE graph launch.  Inputs go
    // through the context's pinned staging buffers (fixed addresses for the graph's copy nodes).
    memcpy(c->h_pos, positions_host, sizeof(double) * 3 * N);
    memcpy(c->h_Z, numbers_host, sizeof(int32_t) * N);
    const bool want_forces = forces_host != nullptr;
    const int tag = p->readout_mode * 1000 + p->pair_block * 10 + p->mean_forces + (p->corr.zbl_on ? 100000 : 0) + (p->corr.exp_on ? 200000 : 0);
    if (c->image_shape_N != N) { c->image_shape_N = N; c->image_shape_calls = 0; }
    const bool have_graph = c->graph_exec && c->graph_N == N && c->graph_forces == int(want_forces) && c->graph_tag == tag;
    if (!have_graph && c->image_shape_calls >= 1) {
      if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
      cudaGraph_t graph = nullptr;
      APAX_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
      const int64_t launches_before = g_launch_count.load();
      const int rc_cap = enqueue_image_path(c, N, want_forces);
      c->graph_launches = g_launch_count.load() - launches_before;
      g_launch_count.fetch_sub(c->graph_launches);   // captured, not executed: counted when the graph is launched
      const cudaError_t e_end = cudaStreamEndCapture(s, &graph);
      if (rc_cap != APAX_OK || e_end != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->image_shape_calls = -1000000;   // capture is not possible here: stay on the eager path
      } else {
        const cudaError_t e_inst = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e_inst != cudaSuccess) { c->graph_exec = nullptr; cudaGetLastError(); c->image_shape_calls = -1000000; }
        else { c->graph_N = N; c->graph_forces = int(want_forces); c->graph_tag = tag; }
      }
    }
    if (c->graph_exec && c->graph_N == N && c->graph_forces == int(want_forces) && c->graph_tag == tag) {
      APAX_CUDA_TRY(cudaGraphLaunch(c->graph_exec, s));
      g_launch_count.fetch_add(c->graph_launches);
    } else {
      APAX_TRY(enqueue_image_path(c, N, want_forces));
    }
    ++c->image_shape_calls;
    APAX_CUDA_TRY(cudaStreamSynchronize(s));
    c->n_pairs = c->h_counters[0];
    if (n_pairs_host) *n_pairs_host = c->n_pairs;
    c->list_atoms = -1;                       // the min-image bookkeeping does not apply to this path
    if (c->h_overflow[0] || c->n_pairs > cap) return APAX_ERR_OVERFLOW;   // caller re-creates the context with more capacity
    memcpy(energy_host, c->h_out, sizeof(double) * n_out);
    if (forces_host) memcpy(forces_host, c->h_out + n_out, sizeof(double) * 3 * N * n_f);
    c->list_atoms = N;
    memcpy(c->list_Z, numbers_host, sizeof(int32_t) * N);
    c->list_minimage = 0;
    c->list_cutoff = double(p->cfg.r_max);
    for (int i = 0; i < 9; ++i) c->list_cell[i] = cell_host[i];
    c->evaluated = want_forces;
    return APAX_OK;
  }
  APAX_CUDA_TRY(cudaMemcpyAsync(c->pos, pos_pinned ? positions_host : c->h_pos, sizeof(double) * 3 * N,
                                cudaMemcpyHostToDevice, s));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->Z, z_pinned ? numbers_host : c->h_Z, sizeof(int32_t) * N, cudaMemcpyHostToDevice, s));
  APAX_CUDA_TRY(cudaMemcpyAsync(c->mats, c->h_mats, sizeof(double) * 45, cudaMemcpyHostToDevice, s));
  const double* d_cell = c->mats;
  const double* d_box = c->mats + 9;
  const double* d_binv = c->mats + 18;
  const double* d_eye = c->mats + 27;

  GmWorkspace gw;
  NlWorkspace nw;
  carve_gm_workspace(c->gm_ws, N, cap, n_out, 128, &gw);
  carve_nl_workspace(c->nl_ws, N, cap, &nw);

  bool same_cell = true;
  for (int i = 0; i < 9; ++i) same_cell = same_cell && c->list_cell[i] == cell_host[i];
  const double cutoff = double(p->cfg.r_max) + (minimage ? dr_threshold : 0.0);
  // The prepared list carries the compact species table (sp_map / embc): a change of atomic numbers at fixed N, cell and
  // within-skin positions must re-run the preparation, as the reference re-initialises on any change of `numbers`
  // (apax/md/ase_calc.py:366-370).
  const bool same_numbers = c->list_atoms == N && memcmp(c->list_Z, numbers_host, sizeof(int32_t) * N) == 0;
  bool need = rebuild || c->list_atoms != N || c->list_minimage != int(minimage) || !same_cell || !same_numbers ||
              c->list_cutoff != cutoff || !minimage;  // the vesin path recomputes every call (ase_calc.py:386-387)
  if (minimage) APAX_LAUNCH(k_to_frac, dim3((unsigned)cdiv(N, 256)), dim3(256), 0, s, c->pos, d_binv, N, c->frac);
  if (minimage && !need) {
    APAX_TRY(apax_nl_needs_rebuild(s, c->frac, c->ref_frac, d_box, N, dr_threshold, c->counters + 1));
    APAX_CUDA_TRY(cudaMemcpyAsync(c->h_counters + 1, c->counters + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    APAX_CUDA_TRY(cudaStreamSynchronize(s));
    need = c->h_counters[1] != 0;
  }
  if (need) {
    APAX_CUDA_TRY(cudaMemsetAsync(c->overflow, 0, 4, s));
    if (minimage) {
      APAX_TRY(nl_build_minimage_csr(s, c->frac, d_box, N, cutoff, cap, nw, nullptr, c->counters, c->overflow));
      APAX_CUDA_TRY(cudaMemcpyAsync(c->ref_frac, c->frac, sizeof(double) * 3 * N, cudaMemcpyDeviceToDevice, s));
    } else {
      // emit only the Cartesian offsets (needed by the displacement); the CSR itself is reused directly
      APAX_TRY(nl_build_images_csr(s, c->pos, d_cell, 1, N, cutoff, cap, nw, nullptr, nullptr, nullptr, c->counters,
                                   c->overflow));
    }
    APAX_CUDA_TRY(cudaMemcpyAsync(c->h_counters, c->counters, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    APAX_CUDA_TRY(cudaMemcpyAsync(c->h_overflow, c->overflow, 1, cudaMemcpyDeviceToHost, s));
    APAX_CUDA_TRY(cudaStreamSynchronize(s));
    c->n_pairs = c->h_counters[0];
    if (c->h_overflow[0] || c->n_pairs > cap) {
      c->list_atoms = -1;
      if (n_pairs_host) *n_pairs_host = c->n_pairs;
      return APAX_ERR_OVERFLOW;  // caller re-creates the context with more capacity (ase_calc.py:381-384)
    }
    if (!minimage) {
      APAX_TRY(nl_emit_offsets(s, nw, N, c->offsets));
    }
    APAX_TRY(gm_prepare_from_csr(s, p, gw, c->Z, nw.rowptr_c, nw.nbr, /*symmetric_sorted=*/minimage));
    c->list_atoms = N;
    memcpy(c->list_Z, numbers_host, sizeof(int32_t) * N);
    c->list_minimage = int(minimage);
    c->list_cutoff = cutoff;
    for (int i = 0; i < 9; ++i) c->list_cell[i] = cell_host[i];
  }
  if (minimage) {
    APAX_TRY(gm_compute_device(s, p, gw, c->frac, c->Z, d_box, nullptr, APAX_DISP_MINIMAGE, c->energy,
                               forces_host ? c->forces : nullptr));
  } else {
    // dr = I.(r_c - r_n) + S@cell: the vesin path of the reference passes inv(cell)-transformed positions and
    // box = cell (ase_calc.py:388-390), whose product is the same Cartesian difference up to fp64 rounding
    APAX_TRY(gm_compute_device(s, p, gw, c->pos, c->Z, d_eye, c->offsets, APAX_DISP_OFFSETS, c->energy,
                               forces_host ? c->forces : nullptr));
  }
  // ---- D2H
  APAX_CUDA_TRY(cudaMemcpyAsync(c->h_out, c->energy, sizeof(double) * n_out, cudaMemcpyDeviceToHost, s));
  const bool f_pinned = forces_host && is_pinned(forces_host);
  if (forces_host)
    APAX_CUDA_TRY(cudaMemcpyAsync(f_pinned ? forces_host : c->h_out + n_out, c->forces, sizeof(double) * 3 * N * n_f,
                                  cudaMemcpyDeviceToHost, s));
  APAX_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(energy_host, c->h_out, sizeof(double) * n_out);
  if (forces_host && !f_pinned) memcpy(forces_host, c->h_out + n_out, sizeof(double) * 3 * N * n_f);
  if (n_pairs_host) *n_pairs_host = c->n_pairs;
  c->evaluated = forces_host != nullptr;
  return APAX_OK;
}

// The `stress` entry of the calculator's results for the structure of the last apax_ctx_calculate (with forces, n_out == 1):
// EnergyDerivativeModel(calc_stress=True) (apax/nn/models.py:160-169, stress_times_vol, apax/layers/properties.py:11-42) followed
// by ProcessStress (apax/md/function_transformations.py:140-159): stress = (dU/d eps)^T / det(box).  On the min-image path the
// per-pair values of the force pass are reduced; on the image-list (vesin) path the reference's disp_fn differentiates at
// doubled displacements (apax/layers/distances.py:16-17), which costs a second force pass -- see apax_gm_stress_times_vol_offsets.
extern "C" int apax_ctx_stress(apax_ctx* c, double* stress_host) {
  if (!c || !stress_host) return APAX_ERR_INVALID_ARG;
  const apax_gm_params_impl* p = c->params;
  if (p->cfg.n_out != 1) return APAX_ERR_UNSUPPORTED;
  if (!c->evaluated || c->list_atoms < 1) return APAX_ERR_INVALID_ARG;   // nothing evaluated (or no forces were asked for)
  const int64_t N = c->list_atoms, cap = c->max_pairs;
  cudaStream_t s = c->stream;
  double* d_stress = c->mats + 45;
  const apax_gm_params* pub = reinterpret_cast<const apax_gm_params*>(p);
  if (c->list_minimage) {
    APAX_TRY(apax_gm_stress_times_vol(reinterpret_cast<apax_stream_t>(s), pub, N, cap, d_stress, c->gm_ws, c->gm_ws_bytes));
  } else {
    if (!c->scratch_ef) APAX_CUDA_TRY(cudaMalloc(&c->scratch_ef, sizeof(double) * (1 + 3 * c->max_atoms)));
    GmWorkspace gw;
    carve_gm_workspace(c->gm_ws, N, cap, 1, 128, &gw);
    // evaluation at dr' = 2 (r_c - r_n) + S.cell, then the virial against the undoubled differences
    APAX_TRY(gm_compute_device(s, p, gw, c->pos, c->Z, c->mats + 36, c->offsets, APAX_DISP_OFFSETS, c->scratch_ef, c->scratch_ef + 1));
    APAX_TRY(apax_gm_stress_times_vol_offsets(reinterpret_cast<apax_stream_t>(s), pub, c->pos, c->mats + 27, N, cap, d_stress,
                                              c->gm_ws, c->gm_ws_bytes));
    c->evaluated = 0;   // the workspace now holds the doubled evaluation: a further stress call needs a fresh calculate
  }
  double st[9];
  APAX_CUDA_TRY(cudaMemcpyAsync(c->h_mats + 45, d_stress, sizeof(double) * 9, cudaMemcpyDeviceToHost, s));
  APAX_CUDA_TRY(cudaStreamSynchronize(s));
  memcpy(st, c->h_mats + 45, sizeof(st));
  const double* a = c->list_cell;
  const double vol = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) stress_host[3 * i + j] = st[3 * j + i] / vol;   // val.T / V, V = det(box) = det(cell)
  return APAX_OK;
}

// Neighbour list + list preparation in one asynchronous call, all on the device: the min-image cell list of nl.cu feeds the
// fused CSR import (binary-search transpose, no COO round trip, no row sort) -- what apax_ctx_calculate does internally, for
// callers that keep positions on the device (MD between host callbacks, the slab-decomposed evaluation).
extern "C" int apax_gm_prepare_minimage(apax_stream_t stream_, const apax_gm_params* params, const double* frac, const int32_t* Z,
                                        const double* box, int64_t n_atoms, int64_t n_central, double cutoff, int64_t capacity,
                                        int32_t* n_found, uint8_t* overflow, void* gm_workspace, size_t gm_workspace_bytes,
                                        void* nl_workspace, size_t nl_workspace_bytes) {
  if (!params || !frac || !Z || !box || !n_found || !overflow || !gm_workspace || !nl_workspace) return APAX_ERR_INVALID_ARG;
  if (n_atoms < 1 || n_central < 1 || n_central > n_atoms || capacity < 1 || !(cutoff > 0.0)) return APAX_ERR_INVALID_ARG;
  if (n_atoms >= (int64_t(1) << 31) - 256 || capacity >= (int64_t(1) << 31) - 256) return APAX_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(gm_workspace) | reinterpret_cast<uintptr_t>(nl_workspace)) & 255) != 0) return APAX_ERR_INVALID_ARG;
  const apax_gm_params_impl* p = reinterpret_cast<const apax_gm_params_impl*>(params);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  GmWorkspace gw;
  NlWorkspace nw;
  if (carve_gm_workspace(gm_workspace, n_atoms, capacity, p->cfg.n_out, 128, &gw) > gm_workspace_bytes) return APAX_ERR_WORKSPACE;
  if (carve_nl_workspace(nl_workspace, n_atoms, capacity, &nw) > nl_workspace_bytes) return APAX_ERR_WORKSPACE;
  APAX_CUDA_TRY(cudaMemsetAsync(overflow, 0, 1, s));
  APAX_TRY(nl_build_minimage_csr(s, frac, box, n_atoms, cutoff, capacity, nw, nullptr, n_found, overflow));
  return gm_prepare_from_csr(s, p, gw, Z, nw.rowptr_c, nw.nbr, /*symmetric_sorted=*/true, n_central);
}
