// Peer-memory exchange for the all-gathered global InfoNCE: every rank owns one exchange block
// (cudaMalloc + CUDA IPC handle); the peers map it (cudaIpcOpenMemHandle -> NVLink / NVSwitch loads and stores) and the
// kernels read remote rows where they live.  Replaces the two host-issued NCCL all-gathers of a step with two
// single-launch device barriers that sit in the same stream as the loss kernels (no host round trip, no extra copy of
// the gathered embeddings: the normalise + bf16 hi/lo split kernel is the gather's only consumer and reads the peers'
// HBM directly).
//
// Exchange block (4-byte words):   [0,16)  flag[r] = last epoch rank r signalled to this rank (monotonic)
//                                  [32]    ticket of the multi-block sync kernel      [33] number of timed-out barriers
//                                  [64..)  slot 0 | slot 1, slot = pooled [2][B][D] fp32 | pack [2B+2] fp32
// Two barriers per step (epoch 2*step+1 after the pooled rows are published, 2*step+2 after the packs are) order every
// re-use of a slot after its last remote reader; the slot alternates with the step parity on top of that.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace cfa {

constexpr size_t kPeerHeaderWords = 64;
constexpr int kPeerTicketWord = 32, kPeerStatusWord = 33;
// A rank that never arrives: poison THIS step's losses (NaN) and count the event instead of hanging.  The wait is as long
// as a collective's watchdog (default 600 s, CFA_PEER_TIMEOUT_MS overrides): a rank-0-only checkpoint or evaluation, or a
// data-loader stall, must not look like a lost rank.  The status word is a COUNTER the host polls (cfa_peer_status /
// PeerExchange.check()): it is not sticky, a later step whose barrier completes is computed normally.
static unsigned long long peer_timeout_ns() {
  static unsigned long long ns = 0;
  if (ns == 0) {
    const char* e = getenv("CFA_PEER_TIMEOUT_MS");
    const double ms = e ? atof(e) : 600000.0;
    ns = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
  }
  return ns;
}

static inline size_t up32w(size_t n) { return (n + 31) & ~(size_t)31; }
size_t peer_slot_words(int B, int D) { return up32w((size_t)2 * B * D) + up32w((size_t)2 * B + 2); }

struct PeerBlocks {
  float* base[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// 1. push  : own block[push_off .. +push_n) = push_src            (all CTAs, grid-stride)
// 2. barrier: the last CTA to finish signals `epoch` into every peer's flag[rank] and waits for flag[r] >= epoch of all r
// 3. pull  : pull_dst[r * pull_n + i] = peer r block[pull_off + i]   (last CTA only; small)
__global__ void __launch_bounds__(1024)
peer_sync_kernel(const PeerBlocks blocks, int world, int rank, unsigned epoch, const float* __restrict__ push_src,
                 size_t push_off, size_t push_n, size_t pull_off, int pull_n, float* __restrict__ pull_dst,
                 float* __restrict__ tail2_sums /* optional [2]: sums over the ranks of the last two pulled words */,
                 unsigned long long timeout_ns) {
  float* own = blocks.base[rank];
  {
    float* dst = own + push_off;
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (((((uintptr_t)push_src) | ((uintptr_t)dst)) & 15) == 0) {
      const size_t n4 = push_n >> 2;
      for (size_t i = i0; i < n4; i += stride) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(push_src)[i];
      for (size_t i = (n4 << 2) + i0; i < push_n; i += stride) dst[i] = push_src[i];
    } else {
      for (size_t i = i0; i < push_n; i += stride) dst[i] = push_src[i];
    }
  }
  __shared__ int s_last, s_timeout;
  unsigned* own_u = reinterpret_cast<unsigned*>(own);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = atomicAdd(own_u + kPeerTicketWord, 1u) == gridDim.x - 1;
    s_timeout = 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) own_u[kPeerTicketWord] = 0;
  __threadfence();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();                                                      // my pushed rows before my flag
    st_release_sys(reinterpret_cast<unsigned*>(blocks.base[r]) + rank, epoch);   // "rank has published epoch", in r's block
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(own_u + r) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) { s_timeout = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_timeout) own_u[kPeerStatusWord] += 1;         // counter read by the host (cfa_peer_status)
  const bool bad = s_timeout != 0;
  for (int r = 0; r < world; ++r)
    for (int i = threadIdx.x; i < pull_n; i += blockDim.x) {
      const float x = ld_relaxed_sys(blocks.base[r] + pull_off + i);
      pull_dst[(size_t)r * pull_n + i] = bad ? __int_as_float(0x7fc00000) : x;    // a missing rank poisons the losses (NaN)
    }
  if (tail2_sums && pull_n >= 2) {                       // the packs end in (sum CE_a, sum CE_b): global sums, fixed rank order
    __syncthreads();
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += pull_dst[(size_t)r * pull_n + pull_n - 2 + threadIdx.x];
      tail2_sums[threadIdx.x] = s;
    }
  }
}

int peer_sync(void* const* h_blocks, int world, int rank, uint32_t epoch, const float* push_src, size_t push_off, size_t push_n,
              size_t pull_off, int pull_n, float* pull_dst, cudaStream_t st, float* tail2_sums) {
  PeerBlocks pb{};
  for (int r = 0; r < world; ++r) {
    if (!h_blocks[r]) return CFA_ERR_BAD_ARG;
    pb.base[r] = (float*)h_blocks[r];
  }
  int nblk = (int)((push_n / 4 + 1023) / 1024);
  if (nblk < 1) nblk = 1;
  if (nblk > 64) nblk = 64;
  peer_sync_kernel<<<nblk, 1024, 0, st>>>(pb, world, rank, epoch, push_src, push_off, push_n, pull_off, pull_n, pull_dst, tail2_sums,
                                          peer_timeout_ns());
  return launch_status();
}

}  // namespace cfa

using namespace cfa;

// ---- exchange-block life cycle (setup time, not on the step path) ----------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");

extern "C" int cfa_peer_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]) {
  if (!dev_ptr || !handle_out || bytes < kPeerHeaderWords * sizeof(float)) return CFA_ERR_BAD_ARG;
  void* p = nullptr;
  CFA_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return (int)e; }
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
  return CFA_OK;
}

extern "C" int cfa_peer_open(const unsigned char handle[64], void** dev_ptr) {
  if (!handle || !dev_ptr) return CFA_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  CFA_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return CFA_OK;
}

extern "C" int cfa_peer_close(void* dev_ptr) {
  if (!dev_ptr) return CFA_OK;
  CFA_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return CFA_OK;
}

extern "C" int cfa_peer_free(void* dev_ptr) {
  if (!dev_ptr) return CFA_OK;
  CFA_CUDA_TRY(cudaFree(dev_ptr));
  return CFA_OK;
}

// number of barriers of this rank that timed out so far: asynchronous copy of the status word into (pinned) host memory
extern "C" int cfa_peer_status(const void* own_block, unsigned int* h_count, void* stream) {
  if (!own_block || !h_count) return CFA_ERR_BAD_ARG;
  CFA_CUDA_TRY(cudaMemcpyAsync(h_count, (const unsigned*)own_block + kPeerStatusWord, sizeof(unsigned), cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  return CFA_OK;
}

// one device barrier (+ optional push / pull) on its own: tests and the two-rank bring-up use it directly
extern "C" int cfa_peer_sync(void* const* h_peer_blocks, int world, int rank, uint32_t epoch, const float* push_src,
                             size_t push_off_words, size_t push_words, size_t pull_off_words, int pull_words,
                             float* pull_dst, void* stream) {
  if (!h_peer_blocks || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return CFA_ERR_BAD_ARG;
  return peer_sync(h_peer_blocks, world, rank, epoch, push_src, push_off_words, push_words, pull_off_words, pull_words,
                   pull_dst, (cudaStream_t)stream, nullptr);
}
