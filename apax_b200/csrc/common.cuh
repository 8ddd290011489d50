// Shared helpers for the apax_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/apax_b200.h"

namespace apax {

extern std::atomic<int64_t> g_launch_count;
extern thread_local cudaError_t g_last_error;

inline int check_launch() {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    g_last_error = cudaGetLastError();
    return APAX_ERR_CUDA;
  }
  return APAX_OK;
}

#define APAX_CUDA_TRY(expr)                 \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) {                \
      ::apax::g_last_error = _e;            \
      return APAX_ERR_CUDA;                 \
    }                                       \
  } while (0)

#define APAX_TRY(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s != APAX_OK) return _s; \
  } while (0)

// optional per-launch CUDA-event timing (apax_profile_enable): events are recorded on the launching
// stream around every kernel of this library, so bench.py can time each kernel live
extern bool g_profile_on;
void profile_before(const char* name, cudaStream_t stream);
void profile_after(cudaStream_t stream);

// launch + count + error check
#define APAX_LAUNCH(kernel, grid, block, smem, stream, ...)        \
  do {                                                             \
    if (::apax::g_profile_on) ::apax::profile_before(#kernel, (stream)); \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
    if (::apax::g_profile_on) ::apax::profile_after((stream));     \
    APAX_TRY(::apax::check_launch());                              \
  } while (0)

// the same with an explicit profile label (template instantiations whose arguments are template parameters at the launch site)
#define APAX_LAUNCH_AS(label, kernel, grid, block, smem, stream, ...) \
  do {                                                                \
    if (::apax::g_profile_on) ::apax::profile_before(label, (stream)); \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
    if (::apax::g_profile_on) ::apax::profile_after((stream));        \
    APAX_TRY(::apax::check_launch());                                 \
  } while (0)

// number of SMs of the current device (queried once per device; 148 on a B200)
int num_sms();

#ifdef __CUDACC__
// Bounded cross-GPU spin: waits until *flag has reached `epoch` (wrap-safe), at most kPeerSpinTimeoutNs.  A peer that died,
// threw on the host or issued a different number of exchanges would otherwise hang this GPU for good; on a time-out the
// sticky error word `err` (in this rank's own header) is set -- the host surfaces it as APAX_ERR_CUDA (apax_halo_status,
// apax_train_peers_status) -- and the kernel carries on with whatever data is there (the results of that call are invalid).
constexpr unsigned long long kPeerSpinTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool peer_spin_until(const uint32_t* flag, uint32_t epoch, uint32_t* err) {
  auto reached = [&]() {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return int32_t(v - epoch) >= 0;
  };
  if (reached()) return true;
  const unsigned long long t0 = global_timer_ns();
  while (!reached()) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > kPeerSpinTimeoutNs) {
      atomicExch(err, 1u);
      return false;
    }
  }
  return true;
}
#endif

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t cdiv(int64_t x, int64_t m) { return (x + m - 1) / m; }

// Bump allocator over a caller-provided workspace; every block is 256-byte aligned.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <class T>
  T* take(int64_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += size_t(count) * sizeof(T);
    return p;
  }
  size_t bytes() const { return (off + 255) & ~size_t(255); }
};

// ---------------------------------------------------------------- exclusive scan (int32)
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries (out[n] = total).  `in` may alias
// `out` (in has n entries).  scratch: cdiv(n+1, 1024) + 1 ints.
int exclusive_scan_i32(cudaStream_t stream, const int32_t* in, int32_t* out, int64_t n, int32_t* scratch);
inline int64_t scan_scratch_ints(int64_t n) { return cdiv(n + 1, 1024) + 2; }

// Sort the entries of every CSR row by ascending key (keys distinct within a row), carrying one
// payload.  Warp per row, rank sort (rows here are ~100 long).  key_out/val_out must not alias inputs.
int rowsort_i32(cudaStream_t stream, const int32_t* rowptr, int64_t n_rows, const int32_t* key_in,
                const int32_t* val_in, int32_t* key_out, int32_t* val_out);

}  // namespace apax
