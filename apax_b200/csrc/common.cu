// Scan / row-sort primitives and ABI diagnostics.
#include "common.cuh"

#include <string.h>

#include <mutex>
#include <string>
#include <vector>

namespace apax {

std::atomic<int64_t> g_launch_count{0};
thread_local cudaError_t g_last_error = cudaSuccess;

int num_sms() {
  static std::atomic<int> cached[64];   // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Per-launch event timing is a measurement aid (bench.py, tools/): one global switch and one record list behind a mutex,
// so enabling it while several host threads launch is safe (their records interleave in launch order).
bool g_profile_on = false;
namespace {
std::mutex g_prof_mutex;
struct ProfRec {
  std::string name;
  cudaEvent_t a, b;
};
std::vector<ProfRec> g_prof;
}  // namespace

void profile_before(const char* name, cudaStream_t stream) {
  ProfRec r;
  r.name = name;
  if (r.name.size() > 2 && r.name.front() == '(' && r.name.back() == ')') r.name = r.name.substr(1, r.name.size() - 2);  // (k<a, b>) at the launch site
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  g_prof_mutex.lock();   // held across the launch so that profile_after finds its own record last
  cudaEventRecord(r.a, stream);
  g_prof.push_back(r);
}
void profile_after(cudaStream_t stream) {
  cudaEventRecord(g_prof.back().b, stream);
  g_prof_mutex.unlock();
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan: 3 launches (block sums, scan of block sums, apply).  1024 elements per block.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 4;
constexpr int kScanTile = kScanThreads * kScanPerThread;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // exclusive scan of one int per thread across a 256-thread block
  __shared__ int warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;  // exclusive
    if (lane == kScanThreads / 32 - 1) *total = wi;
  }
  __syncthreads();
  return incl - v + warp_sums[wid];
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                                 int32_t* __restrict__ tile_sums) {
  __shared__ int total;
  const int64_t base = int64_t(blockIdx.x) * kScanTile + threadIdx.x * kScanPerThread;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k)
    if (base + k < n) s += in[base + k];
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_offsets(int32_t* tile_sums, int64_t n_tiles) {
  // single block; in-place exclusive scan of the tile sums, entry n_tiles receives the grand total
  __shared__ int total;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t start = 0; start < n_tiles; start += kScanThreads) {
    const int64_t i = start + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    const int ex = block_exclusive_scan(v, &total);
    const int c = carry;
    if (i < n_tiles) tile_sums[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const int32_t* in, int32_t* out, int64_t n,
                                                             const int32_t* __restrict__ tile_sums, int64_t n_tiles) {
  __shared__ int total;
  const int64_t base = int64_t(blockIdx.x) * kScanTile + threadIdx.x * kScanPerThread;
  int v[kScanPerThread];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int ex = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    if (base + k <= n) out[base + k] = (base + k < n) ? ex : tile_sums[n_tiles];
    ex += v[k];
  }
}
}  // namespace

int exclusive_scan_i32(cudaStream_t stream, const int32_t* in, int32_t* out, int64_t n, int32_t* scratch) {
  // tiles cover indices [0, n] so that out[n] is written by the apply pass
  const int64_t n_tiles = cdiv(n + 1, kScanTile);
  APAX_LAUNCH(k_scan_tile_sums, dim3((unsigned)n_tiles), dim3(kScanThreads), 0, stream, in, n, scratch);
  APAX_LAUNCH(k_scan_tile_offsets, dim3(1), dim3(kScanThreads), 0, stream, scratch, n_tiles);
  APAX_LAUNCH(k_scan_apply, dim3((unsigned)n_tiles), dim3(kScanThreads), 0, stream, in, out, n, scratch, n_tiles);
  return APAX_OK;
}

// ------------------------------------------------------------------------------------------------
// Row sort: warp per row, rank = number of keys in the row smaller than mine.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_rowsort(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                                 const int32_t* __restrict__ key_in, const int32_t* __restrict__ val_in,
                                                 int32_t* __restrict__ key_out, int32_t* __restrict__ val_out) {
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int b = rowptr[row], e = rowptr[row + 1];
  for (int i = b + lane; i < e; i += 32) {
    const int k = key_in[i];
    int rank = 0;
    for (int j = b; j < e; ++j) rank += (key_in[j] < k);
    key_out[b + rank] = k;
    if (val_in) val_out[b + rank] = val_in[i];
  }
}
}  // namespace

int rowsort_i32(cudaStream_t stream, const int32_t* rowptr, int64_t n_rows, const int32_t* key_in,
                const int32_t* val_in, int32_t* key_out, int32_t* val_out) {
  if (n_rows == 0) return APAX_OK;
  const int64_t blocks = cdiv(n_rows * 32, 256);
  APAX_LAUNCH(k_rowsort, dim3((unsigned)blocks), dim3(256), 0, stream, rowptr, n_rows, key_in, val_in, key_out,
              val_out);
  return APAX_OK;
}

}  // namespace apax

extern "C" {
int apax_abi_version(void) { return APAX_B200_ABI_VERSION; }
const char* apax_last_cuda_error(void) { return cudaGetErrorString(apax::g_last_error); }
int64_t apax_kernel_launch_count(void) { return apax::g_launch_count.load(); }

int apax_profile_enable(int on) {
  std::lock_guard<std::mutex> guard(apax::g_prof_mutex);
  for (auto& r : apax::g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  apax::g_prof.clear();
  apax::g_profile_on = on != 0;
  return APAX_OK;
}

int apax_profile_collect(char* names, int name_stride, float* ms, int max_records) {
  std::lock_guard<std::mutex> guard(apax::g_prof_mutex);
  int n = 0;
  for (auto& r : apax::g_prof) {
    if (n >= max_records) break;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return APAX_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return APAX_ERR_CUDA;
    ms[n] = t;
    strncpy(names + size_t(n) * name_stride, r.name.c_str(), name_stride - 1);
    names[size_t(n) * name_stride + name_stride - 1] = 0;
    ++n;
  }
  return n;
}
}
