// glf_p2p.cu — gradient all-reduce of the data-parallel fusion path as ONE kernel over NVLink / NVSwitch peer memory.
//
// The reference trains under nn.DataParallel (R/main.py:155): every step it reduces all gradients onto GPU 0.  Here each
// GPU is its own process; the only exchange of the fusion path is the average of the 2 x 12 fusion-weight gradients
// (0.35 M floats at C = 256).  An NCCL all-reduce of that size costs ~245 us per step on this box between the graph
// replay and the next step (profiles/README.md); this kernel does it in one launch that is CUDA-graph capturable:
//
//   every rank's flat gradient bucket lives in an IPC-shared allocation, preceded by a signal pad;
//   block b of rank r:  barrier with block b of every peer  (all gradients written)
//                       sum the block's elements over all ranks, in rank order (bitwise identical on every rank)
//                       barrier again                        (every peer has finished reading my bucket)
//                       write scale * sum into my own bucket, in place
//
// Block-level barriers only: block b touches the same element range on every rank, so no grid-wide synchronisation is
// needed.  Signals are self-resetting (signal = CAS 0 -> 1 on the peer's pad, wait = CAS 1 -> 0 on my own pad), so
// the kernel has no epoch argument and replays unchanged from a CUDA graph.
#include <cstring>

#include <cstdlib>

#include "glf_internal.h"

namespace glf {

namespace {

constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_THREADS = 256;
constexpr int P2P_VPT = 8;   // float4 vectors per thread, held in registers across the second barrier

struct P2PParams {
  float* bufs[P2P_MAX_WORLD];
  uint32_t* sigs[P2P_MAX_WORLD];
  int rank, world;
  long long n4;     // float4 vectors
  float scale;
  unsigned long long timeout_ns;   // 0 = wait for the peers as long as it takes (the default, like an NCCL collective)
};

// Ranks of a training job drift apart by far more than any fixed bound (rank-0 checkpoints, validation, data-loader
// stalls), so by default the spins below wait indefinitely, exactly as an NCCL collective would; a hung peer is the host
// watchdog's business.  GLF_P2P_TIMEOUT_S=<seconds> (debugging aid, read at launch) bounds them: on expiry the kernel
// traps, which is sticky for the CUDA context -- use it to find mismatched call counts, not in production.
// Co-residency: block b of rank A waits for block b of rank B, so every block of the grid must be resident on every
// rank at the same time; p2p_allreduce() checks grid <= SMs x blocks-per-SM before launching, and callers must not
// run another kernel that holds all SMs indefinitely on the same device.
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void signal_peer(uint32_t* addr, unsigned long long P2P_TIMEOUT_NS) {
  uint32_t old;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 0u && ++spins > 4096u) {
      const unsigned long long now = p2p_now_ns();
      if (t0 == 0) t0 = now;
      else if (P2P_TIMEOUT_NS != 0 && now - t0 > P2P_TIMEOUT_NS) __trap();
    }
  } while (old != 0u);
}
__device__ __forceinline__ void wait_own(uint32_t* addr, unsigned long long P2P_TIMEOUT_NS) {
  uint32_t old;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 1u && ++spins > 4096u) {
      const unsigned long long now = p2p_now_ns();
      if (t0 == 0) t0 = now;
      else if (P2P_TIMEOUT_NS != 0 && now - t0 > P2P_TIMEOUT_NS) __trap();
    }
  } while (old != 1u);
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// pad layout per rank: [phase 0 | phase 1][gridDim.x blocks][world] uint32
__device__ __forceinline__ void block_barrier(const P2PParams& p, int phase) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < p.world && t != p.rank) {
    __threadfence_system();
    const long long slot = (static_cast<long long>(phase) * gridDim.x + blockIdx.x) * p.world;
    signal_peer(p.sigs[t] + slot + p.rank, p.timeout_ns);
    wait_own(p.sigs[p.rank] + slot + t, p.timeout_ns);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_kernel(const P2PParams p) {
  block_barrier(p, 0);
  float4 acc[P2P_VPT];
  const long long stride = static_cast<long long>(gridDim.x) * P2P_THREADS;
  const long long i0 = static_cast<long long>(blockIdx.x) * P2P_THREADS + threadIdx.x;
#pragma unroll
  for (int it = 0; it < P2P_VPT; ++it) {
    const long long i = i0 + it * stride;
    acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < p.n4) {
      for (int r = 0; r < p.world; ++r) {
        const float4 v = ld_sys_v4(p.bufs[r] + 4 * i);
        acc[it].x += v.x; acc[it].y += v.y; acc[it].z += v.z; acc[it].w += v.w;
      }
    }
  }
  block_barrier(p, 1);
  float* mine = p.bufs[p.rank];
#pragma unroll
  for (int it = 0; it < P2P_VPT; ++it) {
    const long long i = i0 + it * stride;
    if (i < p.n4)
      *reinterpret_cast<float4*>(mine + 4 * i) =
          make_float4(acc[it].x * p.scale, acc[it].y * p.scale, acc[it].z * p.scale, acc[it].w * p.scale);
  }
}

}  // namespace

int p2p_grid(long long n_floats) {
  const long long n4 = (n_floats + 3) / 4;
  long long g = (n4 + P2P_THREADS - 1) / P2P_THREADS;
  if (g > 148) g = 148;
  return static_cast<int>(g < 1 ? 1 : g);
}
long long p2p_max_floats() { return 148LL * P2P_THREADS * P2P_VPT * 4; }
size_t p2p_signal_bytes(int world) { return static_cast<size_t>(2) * 148 * world * sizeof(uint32_t); }

int p2p_allreduce(void* const* bufs, void* const* sigs, int rank, int world, long long n_floats, float scale,
                  cudaStream_t stream) {
  if (world < 2 || world > P2P_MAX_WORLD || rank < 0 || rank >= world)
    return set_error(GLF_ERR_INVALID, "p2p all-reduce: world must be 2..%d", P2P_MAX_WORLD);
  if (n_floats <= 0 || n_floats % 4 != 0 || n_floats > p2p_max_floats())
    return set_error(GLF_ERR_INVALID, "p2p all-reduce: element count must be a multiple of 4 and <= %lld", p2p_max_floats());
  P2PParams p;
  for (int r = 0; r < world; ++r) {
    if (bufs[r] == nullptr || sigs[r] == nullptr || (reinterpret_cast<uintptr_t>(bufs[r]) & 15) != 0)
      return set_error(GLF_ERR_INVALID, "p2p all-reduce: NULL or misaligned peer pointer");
    p.bufs[r] = reinterpret_cast<float*>(bufs[r]);
    p.sigs[r] = reinterpret_cast<uint32_t*>(sigs[r]);
  }
  p.rank = rank; p.world = world;
  p.n4 = n_floats / 4;
  p.scale = scale;
  p.timeout_ns = 0;
  if (const char* e = getenv("GLF_P2P_TIMEOUT_S")) {
    const double sec = atof(e);
    if (sec > 0) p.timeout_ns = static_cast<unsigned long long>(sec * 1e9);
  }
  // every block spins on its peers: the whole grid must be co-resident
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p2p_allreduce_kernel, P2P_THREADS, 0) != cudaSuccess)
    return set_error(GLF_ERR_DEVICE, "p2p all-reduce: cannot query occupancy");
  if (static_cast<long long>(sms) * per_sm < p2p_grid(n_floats))
    return set_error(GLF_ERR_DEVICE, "p2p all-reduce: %d blocks cannot be co-resident on %d SMs", p2p_grid(n_floats), sms);
  p2p_allreduce_kernel<<<p2p_grid(n_floats), P2P_THREADS, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "p2p all-reduce launch");
}

// ---- IPC plumbing (setup time, not on the hot path): export / open a device allocation across processes -------------
int p2p_export(const void* ptr, unsigned char handle[64], unsigned long long* offset) {
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || f == nullptr)
    return set_error(GLF_ERR_DEVICE, "cuMemGetAddressRange entry point unavailable");
  if (reinterpret_cast<RangeFn>(f)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS)
    return set_error(GLF_ERR_DEVICE, "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  int rc = check_cuda(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)), "cudaIpcGetMemHandle");
  if (rc) return rc;
  memcpy(handle, &h, 64);
  *offset = static_cast<unsigned long long>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}

int p2p_open(const unsigned char handle[64], unsigned long long offset, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  int rc = check_cuda(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  if (rc) return rc;
  *out = reinterpret_cast<unsigned char*>(base) + offset;
  return 0;
}

int p2p_close(void* ptr, unsigned long long offset) {
  return check_cuda(cudaIpcCloseMemHandle(reinterpret_cast<unsigned char*>(ptr) - offset), "cudaIpcCloseMemHandle");
}

}  // namespace glf
