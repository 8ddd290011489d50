// Halo exchange of the slab-decomposed energy+forces evaluation (BASELINE configs[3]) as two kernels over
// NVLink peer memory instead of library collectives.  The reference has no counterpart (single device,
// apax/md/simulate.py:313-354).
//
// Every rank keeps its local position rows R[n_local][3] and force rows F[n_local][3] (owned atoms first, then ghosts
// grouped by owner rank) in CUDA-IPC memory that its peers have mapped (apax_peer_alloc / apax_peer_open), behind a
// 512-byte header: u32 fwd_flags[8] | u32 rev_flags[8] | u32 ticket | u32 error (sticky: a barrier timed out) | pad |
// f64 energy[16] at byte 128 | u32 red_flags[8] at byte 256 | f64 red_val[2][16] at byte 320 (small all-reduce, two parities).
//
//   forward  (apax_halo_push_rows)     : pack-and-store -- each boundary atom's position goes straight into the ghost row
//                                        of the neighbouring rank's R (remote stores); the last block to finish publishes
//                                        "my rows of step t have landed" (release store into every peer's flag array) and
//                                        waits for the same from every peer, so the kernel's completion IS the barrier.
//   reverse  (apax_halo_pull_add_rows) : publish "my forces of step t are complete", wait for the peers, then every owner
//                                        PULLS the contributions its boundary atoms received as ghosts elsewhere (remote
//                                        loads) and adds them in rank order -- deterministic, no remote atomics; the
//                                        8-byte energy all-reduce rides along (each rank sums the headers in rank order).
// Why one buffer is enough: a rank pushes step t+1 only after its pull of step t, which needed every peer's "forces
// complete" flag, i.e. the peer's compute of step t had finished reading its R; and a peer overwrites its F in step t+1 only
// after the forward barrier of t+1, which needs this rank's push of t+1, issued after this rank's pull of step t.
#include <algorithm>

#include "common.cuh"

namespace apax {
namespace {

constexpr int HDR_FWD = 0, HDR_REV = 8, HDR_TICKET = 16, HDR_ERR = 17;   // u32 slots of the header
constexpr int HDR_ENERGY_BYTE = 128;
constexpr int HDR_RED = 64;            // u32 slot of the small all-reduce's flags (byte 256)
constexpr int HDR_REDVAL_BYTE = 320;   // f64 [2][16]

struct HaloArgs {
  double* rows[8];       // per rank r: first row (in r's block, as mapped here) that belongs to THIS rank's atoms
  uint32_t* hdr[8];      // per rank: its header as mapped here
  int64_t send_ptr[9];   // my send list, grouped by destination rank
  int world, rank;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// every peer's flag in MY header has reached `epoch`, or the wait timed out (error word set, see peer_spin_until)
__device__ __forceinline__ void wait_flags(uint32_t* my_hdr, int first_slot, int world, uint32_t epoch) {
  if (threadIdx.x < world) peer_spin_until(my_hdr + first_slot + threadIdx.x, epoch, my_hdr + HDR_ERR);
  __syncthreads();
}

// thread per (send entry): row send_local[i] of my R -> row (i - send_ptr[r]) of my region in rank r's R
__global__ void __launch_bounds__(256) k_halo_push(HaloArgs h, const double* __restrict__ R, const int64_t* __restrict__ send_local) {
  const int64_t n_send = h.send_ptr[h.world];
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_send; i += int64_t(gridDim.x) * blockDim.x) {
    int r = 0;
    while (i >= h.send_ptr[r + 1]) ++r;
    const int64_t src = send_local[i], dst = i - h.send_ptr[r];
    double* out = h.rows[r] + 3 * dst;
    out[0] = R[3 * src + 0]; out[1] = R[3 * src + 1]; out[2] = R[3 * src + 2];
  }
  // completion = barrier: the last block publishes and waits
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* ticket = h.hdr[h.rank] + HDR_TICKET;
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last) *ticket = 0;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < h.world) {
    __threadfence_system();
    st_release_sys(h.hdr[threadIdx.x] + HDR_FWD + h.rank, h.epoch);
  }
  wait_flags(h.hdr[h.rank], HDR_FWD, h.world, h.epoch);
}

// phase 0: publish + wait (all blocks), then pull from rank `src` only; later launches (phase 1) just pull from their rank.
__global__ void __launch_bounds__(256) k_halo_pull(HaloArgs h, double* __restrict__ F, const int64_t* __restrict__ send_local,
                                                   int src, int phase, int n_out, double* __restrict__ energy) {
  if (phase == 0) {
    if (blockIdx.x == 0 && threadIdx.x < h.world) {
      __threadfence_system();
      st_release_sys(h.hdr[threadIdx.x] + HDR_REV + h.rank, h.epoch);
    }
    wait_flags(h.hdr[h.rank], HDR_REV, h.world, h.epoch);
    if (blockIdx.x == 0 && threadIdx.x < n_out) {   // energy: sum of the ranks' partial energies in rank order
      double e = 0.0;
      for (int r = 0; r < h.world; ++r)
        e += __ldcg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(h.hdr[r]) + HDR_ENERGY_BYTE) + threadIdx.x);
      energy[threadIdx.x] = e;
    }
  }
  if (src < 0) return;
  const int64_t b = h.send_ptr[src], e = h.send_ptr[src + 1];
  const double* in = h.rows[src];
  for (int64_t i = b + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < e; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t a = send_local[i];
    const double* q = in + 3 * (i - b);
    F[3 * a + 0] += __ldcg(q + 0); F[3 * a + 1] += __ldcg(q + 1); F[3 * a + 2] += __ldcg(q + 2);
  }
}

__global__ void k_halo_store_energy(const double* __restrict__ e_local, int n_out, uint32_t* hdr) {
  if (threadIdx.x < n_out)
    reinterpret_cast<double*>(reinterpret_cast<char*>(hdr) + HDR_ENERGY_BYTE)[threadIdx.x] = e_local[threadIdx.x];
}

// Small all-reduce (sum) of up to 16 doubles over the peer headers: every rank stores its values in its own header (parity
// epoch & 1), publishes the epoch to every peer, waits for all, and adds the ranks' values in rank order -- every rank gets
// the bitwise identical sum.  Parities make one call safe against the next (a rank can be at most one call ahead: it needs
// every peer's flag of call t before it can issue call t+1).
__global__ void __launch_bounds__(32) k_peer_allreduce(HaloArgs h, const double* __restrict__ local, double* __restrict__ out, int n) {
  const int par = int(h.epoch & 1u);
  double* mine = reinterpret_cast<double*>(reinterpret_cast<char*>(h.hdr[h.rank]) + HDR_REDVAL_BYTE) + par * 16;
  if (threadIdx.x < n) mine[threadIdx.x] = local[threadIdx.x];
  __syncwarp();
  if (threadIdx.x < h.world) {
    __threadfence_system();
    st_release_sys(h.hdr[threadIdx.x] + HDR_RED + h.rank, h.epoch);
    peer_spin_until(h.hdr[h.rank] + HDR_RED + threadIdx.x, h.epoch, h.hdr[h.rank] + HDR_ERR);
  }
  __syncwarp();
  if (threadIdx.x < n) {
    double s = 0.0;
    for (int r = 0; r < h.world; ++r)
      s += __ldcg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(h.hdr[r]) + HDR_REDVAL_BYTE) + par * 16 + threadIdx.x);
    out[threadIdx.x] = s;
  }
}

int fill(HaloArgs& h, const apax_halo_peers* p, const int64_t* send_ptr_host, uint32_t epoch) {
  if (!p || !send_ptr_host || p->world < 1 || p->world > 8 || p->rank < 0 || p->rank >= p->world) return APAX_ERR_INVALID_ARG;
  for (int r = 0; r < p->world; ++r) {
    if (!p->header[r]) return APAX_ERR_INVALID_ARG;
    h.rows[r] = p->rows[r];
    h.hdr[r] = p->header[r];
    if (send_ptr_host[r + 1] < send_ptr_host[r] || (send_ptr_host[r + 1] > send_ptr_host[r] && !p->rows[r])) return APAX_ERR_INVALID_ARG;
  }
  for (int r = 0; r <= p->world; ++r) h.send_ptr[r] = send_ptr_host[r];
  h.world = p->world; h.rank = p->rank; h.epoch = epoch;
  return APAX_OK;
}

}  // namespace
}  // namespace apax

using namespace apax;

extern "C" int apax_halo_push_rows(apax_stream_t stream_, const double* local_rows, const int64_t* send_local,
                                   const int64_t* send_ptr_host, const apax_halo_peers* peers, uint32_t epoch) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!local_rows) return APAX_ERR_INVALID_ARG;
  HaloArgs h{};
  APAX_TRY(fill(h, peers, send_ptr_host, epoch));
  const int64_t n_send = h.send_ptr[h.world];
  if (n_send > 0 && !send_local) return APAX_ERR_INVALID_ARG;
  const int64_t blocks = std::min<int64_t>(std::max<int64_t>(cdiv(n_send, 256), 1), 4 * int64_t(num_sms()));
  APAX_LAUNCH(k_halo_push, dim3((unsigned)blocks), dim3(256), 0, s, h, local_rows, send_local);
  return APAX_OK;
}

extern "C" int apax_halo_pull_add_rows(apax_stream_t stream_, double* local_rows, const int64_t* send_local,
                                       const int64_t* send_ptr_host, const apax_halo_peers* peers, uint32_t epoch,
                                       const double* energy_local, double* energy_global, int32_t n_out) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!local_rows || !energy_local || !energy_global || n_out < 1 || n_out > 16) return APAX_ERR_INVALID_ARG;
  HaloArgs h{};
  APAX_TRY(fill(h, peers, send_ptr_host, epoch));
  APAX_LAUNCH(k_halo_store_energy, dim3(1), dim3(32), 0, s, energy_local, n_out, h.hdr[h.rank]);
  int phase = 0;
  for (int r = 0; r < h.world; ++r) {   // one launch per source rank, in rank order: contributions to one atom never race
    const int64_t cnt = h.send_ptr[r + 1] - h.send_ptr[r];
    if (cnt == 0) continue;
    const int64_t blocks = std::min<int64_t>(cdiv(cnt, 256), 4 * int64_t(num_sms()));
    APAX_LAUNCH(k_halo_pull, dim3((unsigned)blocks), dim3(256), 0, s, h, local_rows, send_local, r, phase, n_out, energy_global);
    phase = 1;
  }
  if (phase == 0) APAX_LAUNCH(k_halo_pull, dim3(1), dim3(256), 0, s, h, local_rows, send_local, -1, 0, n_out, energy_global);
  return APAX_OK;
}

extern "C" int apax_peer_allreduce_sum(apax_stream_t stream_, const apax_halo_peers* peers, uint32_t epoch, const double* local,
                                       double* global_out, int32_t n) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (!local || !global_out || n < 1 || n > 16) return APAX_ERR_INVALID_ARG;
  HaloArgs h{};
  const int64_t none[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  APAX_TRY(fill(h, peers, none, epoch));
  APAX_LAUNCH(k_peer_allreduce, dim3(1), dim3(32), 0, s, h, local, global_out, int(n));
  return APAX_OK;
}

extern "C" int apax_halo_status(const apax_halo_peers* peers) {
  // synchronising query of this rank's sticky barrier-time-out word (0 = healthy)
  if (!peers || peers->rank < 0 || peers->rank >= 8 || !peers->header[peers->rank]) return APAX_ERR_INVALID_ARG;
  uint32_t err = 0;
  APAX_CUDA_TRY(cudaMemcpy(&err, peers->header[peers->rank] + HDR_ERR, sizeof(err), cudaMemcpyDeviceToHost));
  return err == 0 ? APAX_OK : APAX_ERR_CUDA;
}
