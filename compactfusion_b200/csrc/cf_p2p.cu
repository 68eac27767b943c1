// One-sided NVLink transport for the compressed payloads.
//
// Reference: the exchange step of compact_all_gather (main.py:409, dist.all_gather) and of the
// ring (ring.py:268-269, batch_isend_irecv).  Payloads are 0.2-0.9 MB: the NCCL collective is
// latency-bound and cannot be captured in a CUDA graph together with our kernels on this
// stack, so every layer would pay several CPU launches.  Here each rank pushes its payload
// straight into every peer's receive slot over NVLink (peer-mapped memory, 16-byte stores)
// and publishes a per-(layer, origin) counter flag; the receiver's decompress kernel waits
// on the flags of the origins it consumes.  No collective, no host involvement: the whole
// step (compress -> put -> decompress, all layers) is one CUDA graph.
//
// Memory model: the CTA synchronises after its stores, thread 0 issues a system-scope fence
// (cumulative over the CTA's stores, which the barrier ordered before it) and takes a ticket;
// the CTA that draws the last ticket fences again and publishes the flags with relaxed
// system-scope stores (fence + store = release pattern).  Receivers poll with relaxed
// system-scope loads and re-read the satisfied flag with an acquire load.
#include <stdlib.h>
#include <string.h>

#include "cf_common.cuh"
#include "cf_pipe.cuh"

namespace cf {

struct PutParams {
  const uint4* src;
  size_t n16;  // payload size in 16-byte units
  uint4* dst[CF_MAX_PEERS];
  uint32_t* flag[CF_MAX_PEERS];
  int n_peers;
  uint32_t* count;  // local: puts issued so far on this slot (the value published to the peers)
  uint32_t* done;   // local: CTA ticket counter, reset by the last CTA
};

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid (chunks, n_peers), block 256
__global__ void __launch_bounds__(256) k_p2p_put(const PutParams p) {
  pdl_wait();  // the payload was written by the preceding kernels of this stream
  pdl_launch_dependents();
  uint4* __restrict__ dst = p.dst[blockIdx.y];
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; i + 3 * stride < p.n16; i += 4 * stride) {
    const uint4 a = p.src[i], b = p.src[i + stride], c = p.src[i + 2 * stride], d = p.src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < p.n16; i += stride) dst[i] = p.src[i];
  uint32_t v = 0;
  if (threadIdx.x == 0) v = *p.count + 1u;  // written only by the previous put's last CTA: long complete
  __syncthreads();
  if (threadIdx.x == 0) {
    // cumulative over the whole CTA's stores (ordered before this fence by the barrier)
    __threadfence_system();
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned ticket = atomicAdd(p.done, 1u);
    if (ticket == total - 1) {
      // one fence (cumulative over the other CTAs' stores, observed through the ticket) + relaxed flag
      // stores: a st.release per flag would cost a MEMBAR.SYS round trip per destination
      __threadfence_system();
      for (int q = 0; q < p.n_peers; ++q) st_relaxed_sys(p.flag[q], v);
      *p.count = v;
      *p.done = 0u;
    }
  }
}

// Stand-alone flag wait for consumers that do not wait inside their own kernel (the low-rank reconstruct):
// one warp, lane i polls flag i until it reaches *expected (relaxed polls, then one acquire load), ~2 s bound.
// Kernels launched behind it on the stream (programmatic dependent launch: their griddepcontrol.wait, or plain
// stream order) then read payloads that have fully landed.
struct WaitParams {
  const uint32_t* flag[CF_MAX_PEERS];
  const uint32_t* expected;
  uint32_t* error;
  int n;
};
__global__ void __launch_bounds__(32) k_p2p_wait(const WaitParams p) {
  pdl_wait();   // *expected is written by this rank's own put, earlier on the stream
  const int lane = threadIdx.x;
  if (lane < p.n) {
    uint32_t want, cur;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(want) : "l"(p.expected) : "memory");
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(p.flag[lane]) : "memory");
    const long long t0 = clock64();
    while (static_cast<int32_t>(cur - want) < 0) {
      if (clock64() - t0 > 4000000000LL) {
        atomicExch(p.error, 1u);
        break;
      }
      __nanosleep(32);
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(p.flag[lane]) : "memory");
    }
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(p.flag[lane]) : "memory");
  }
  __syncwarp();
  __threadfence();
  pdl_launch_dependents();
}

}  // namespace cf

extern "C" {

int cf_p2p_wait(int n, const void* const* flags, const void* expected, void* error_word, cf_stream_t stream) {
  CF_CHECK_ARG(flags && expected && error_word, "null pointer");
  CF_CHECK_ARG(n >= 1 && n <= CF_MAX_PEERS, "n %d out of range [1,%d]", n, CF_MAX_PEERS);
  cf::WaitParams p{};
  p.n = n;
  p.expected = static_cast<const uint32_t*>(expected);
  p.error = static_cast<uint32_t*>(error_word);
  for (int i = 0; i < n; ++i) {
    CF_CHECK_ARG(flags[i] != nullptr, "flag %d is null", i);
    p.flag[i] = static_cast<const uint32_t*>(flags[i]);
  }
  cf::k_p2p_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  CF_CHECK_LAUNCH();
  return CF_OK;
}

int cf_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  CF_CHECK_ARG(dev_ptr != nullptr && handle64 != nullptr && bytes > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void* p = nullptr;
  CF_CHECK_CUDA(cudaMalloc(&p, bytes));
  CF_CHECK_CUDA(cudaMemset(p, 0, bytes));
  CF_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cf::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return CF_ERR_CUDA;
  }
  memcpy(handle64, &h, sizeof(h));
  *dev_ptr = p;
  return CF_OK;
}

int cf_ipc_open(const void* handle64, void** peer_ptr) {
  CF_CHECK_ARG(handle64 != nullptr && peer_ptr != nullptr, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  CF_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return CF_OK;
}

int cf_ipc_close(void* peer_ptr) {
  if (peer_ptr) CF_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return CF_OK;
}

int cf_ipc_free(void* dev_ptr) {
  if (dev_ptr) CF_CHECK_CUDA(cudaFree(dev_ptr));
  return CF_OK;
}

int cf_p2p_put(const void* src, size_t bytes, int n_peers, void* const* peer_dst, void* const* peer_flag,
               void* local_count, void* local_ticket, cf_stream_t stream) {
  CF_CHECK_ARG(src && peer_dst && peer_flag && local_count && local_ticket, "null pointer");
  CF_CHECK_ARG(n_peers >= 1 && n_peers <= CF_MAX_PEERS, "n_peers %d out of range [1,%d]", n_peers, CF_MAX_PEERS);
  CF_CHECK_ARG(bytes > 0 && bytes % 16 == 0 && cf::aligned16(src), "payload must be 16-byte aligned and sized");
  cf::PutParams p{};
  p.src = static_cast<const uint4*>(src);
  p.n16 = bytes / 16;
  p.n_peers = n_peers;
  for (int q = 0; q < n_peers; ++q) {
    CF_CHECK_ARG(peer_dst[q] && peer_flag[q] && cf::aligned16(peer_dst[q]), "peer %d: bad destination", q);
    p.dst[q] = static_cast<uint4*>(peer_dst[q]);
    p.flag[q] = static_cast<uint32_t*>(peer_flag[q]);
  }
  p.count = static_cast<uint32_t*>(local_count);
  p.done = static_cast<uint32_t*>(local_ticket);
  // enough CTAs to keep ~64 KB in flight per peer, few enough to leave the SMs to the codecs
  int chunks = static_cast<int>((p.n16 + 256 * 4 - 1) / (256 * 4));
  const int cap = cf::sm_count() / n_peers > 0 ? cf::sm_count() / n_peers : 1;
  if (chunks > cap) chunks = cap;
  if (chunks > 16) chunks = 16;
  if (chunks < 1) chunks = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(chunks, n_peers);
  cfg.blockDim = dim3(256);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const char* e = getenv("CF_PDL");
  cfg.numAttrs = (e && e[0] == '0') ? 0 : 1;
  CF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cf::k_p2p_put, p));
  return CF_OK;
}

}  // extern "C"
