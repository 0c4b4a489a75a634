// NCCL-free synchronisation and dense-gradient all-reduce over peer-mapped memory (CUDA IPC /
// NVLink), so that the whole sharded training step -- including its cross-GPU steps -- is a
// sequence of our own kernels and can be captured in ONE CUDA graph per rank.
//
//   p2p_barrier   : epoch barrier.  sig[r] of every rank is written by rank r (system-scope store
//                   after a system fence) and polled locally.  The epoch lives on the device, so
//                   a graph replay advances it by itself.  Replaces dist.barrier()/all_reduce(1).
//   p2p_allreduce : two-shot mean all-reduce of the dense-gradient arena (DDP's bucketed
//                   all-reduce, extend_distributed.py:14 / dlrm_s_pytorch.py:1329-1336): rank r
//                   reduces slice r reading all peers (fixed order -> identical result on every rank)
//                   and stores the mean back into every peer's slice r.
#include <string.h>

#include "common.cuh"

namespace dlrm {

struct PeerPtrs {
  void* p[DLRM_B200_MAX_PEERS];
};

__global__ void __launch_bounds__(32) p2p_barrier_kernel(PeerPtrs sig, int rank, int world, int* epoch) {
  __shared__ int e;
  if (threadIdx.x == 0) {
    e = *epoch + 1;
    *epoch = e;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    volatile int* remote = static_cast<volatile int*>(sig.p[t]) + rank;   // my slot on rank t
    *remote = e;
    volatile int* mine = static_cast<volatile int*>(sig.p[rank]) + t;     // rank t's slot on me
    // a lost peer must trap, not hang the GPU forever -- but ordinary host skew between ranks (seconds while
    // one of them builds a plan or captures a graph) must not: the limit is 60 s of wall clock, not a spin count
    unsigned long long t0 = 0;
    while (*mine < e) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 60000000000ull) __trap();
    }
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(256) p2p_allreduce_mean_kernel(PeerPtrs grad, int rank, int world, long long n) {
  // slice of this rank, in float4 units where possible
  const long long per = (n + world - 1) / world;
  const long long lo = (long long)rank * per;
  const long long hi = lo + per < n ? lo + per : n;
  const float inv = 1.0f / (float)world;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < world; ++r) acc += static_cast<const float*>(grad.p[r])[i];   // fixed order
    acc *= inv;
    for (int r = 0; r < world; ++r) static_cast<float*>(grad.p[r])[i] = acc;
  }
}

}  // namespace dlrm

extern "C" int dlrm_b200_p2p_barrier(void* const* peer_sig, int rank, int world, int32_t* epoch, void* stream) {
  using namespace dlrm;
  if (!peer_sig || !epoch || world < 1 || world > DLRM_B200_MAX_PEERS || rank < 0 || rank >= world)
    return set_error("p2p_barrier: bad arguments (world=%d rank=%d)", world, rank);
  PeerPtrs s{};
  for (int r = 0; r < world; ++r) {
    if (!peer_sig[r]) return set_error("p2p_barrier: peer %d pointer is NULL", r);
    s.p[r] = peer_sig[r];
  }
  p2p_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(s, rank, world, epoch);
  DLRM_CHECK_LAUNCH("p2p_barrier_kernel");
  return 0;
}

extern "C" int dlrm_b200_p2p_allreduce_mean(void* const* peer_grad, int rank, int world, int64_t n, void* stream) {
  using namespace dlrm;
  if (!peer_grad || world < 1 || world > DLRM_B200_MAX_PEERS || rank < 0 || rank >= world || n < 0)
    return set_error("p2p_allreduce_mean: bad arguments (world=%d rank=%d)", world, rank);
  if (n == 0) return 0;
  PeerPtrs g{};
  for (int r = 0; r < world; ++r) {
    if (!peer_grad[r]) return set_error("p2p_allreduce_mean: peer %d pointer is NULL", r);
    g.p[r] = peer_grad[r];
  }
  const long long per = (n + world - 1) / world;
  long long blocks = (per + 255) / 256;
  if (blocks > 592) blocks = 592;
  p2p_allreduce_mean_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, rank, world, n);
  DLRM_CHECK_LAUNCH("p2p_allreduce_mean_kernel");
  return 0;
}

extern "C" int dlrm_b200_enable_peer_access(int device, int peer_device) {
  using namespace dlrm;
  if (device == peer_device) return 0;
  int can = 0;
  DLRM_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) return set_error("enable_peer_access: device %d cannot access device %d", device, peer_device);
  int cur = 0;
  DLRM_CUDA(cudaGetDevice(&cur));
  DLRM_CUDA(cudaSetDevice(device));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
  cudaSetDevice(cur);
  if (e != cudaSuccess) return set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer_device, cudaGetErrorString(e));
  return 0;
}

// Map a buffer exported by another process (cudaIpcGetMemHandle; PyTorch: storage._share_cuda_()) into
// THIS process for kernels running on `device`.  The handle must be opened with `device` current:
// a mapping opened under the exporter's device index is not reachable from kernels of another device,
// even with peer access enabled (measured on this stack).
extern "C" int dlrm_b200_ipc_open(const void* handle64, int device, void** base_out) {
  using namespace dlrm;
  if (!handle64 || !base_out) return set_error("ipc_open: NULL argument");
  int cur = 0;
  DLRM_CUDA(cudaGetDevice(&cur));
  DLRM_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error("cudaIpcOpenMemHandle on device %d: %s", device, cudaGetErrorString(e));
  }
  *base_out = p;
  return 0;
}

// Export the cudaMalloc allocation holding `ptr` (a device pointer of THIS process): the 64-byte handle
// for dlrm_b200_ipc_open in another process, and ptr's byte offset inside that allocation.
extern "C" int dlrm_b200_ipc_export(const void* ptr, void* handle64_out, int64_t* offset_out) {
  using namespace dlrm;
  if (!ptr || !handle64_out || !offset_out) return set_error("ipc_export: NULL argument");
  typedef int (*RangeFn)(unsigned long long*, size_t*, unsigned long long);
  static RangeFn range = nullptr;
  if (!range) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return set_error("ipc_export: cuMemGetAddressRange not available");
    range = reinterpret_cast<RangeFn>(f);
  }
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = range(&base, &size, (unsigned long long)(uintptr_t)ptr);
  if (rc != 0) return set_error("ipc_export: cuMemGetAddressRange failed (CUresult %d)", rc);
  cudaIpcMemHandle_t h;
  DLRM_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle64_out, &h, sizeof(h));
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
  return 0;
}

extern "C" int dlrm_b200_ipc_close(void* base) {
  using namespace dlrm;
  if (!base) return 0;
  DLRM_CUDA(cudaIpcCloseMemHandle(base));
  return 0;
}
