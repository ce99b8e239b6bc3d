// byol_b200 — SyncBatchNorm statistic exchange over NVLink peer memory (no NCCL call, CUDA-graph capturable).
//
// Replaces the per-layer collectives of torch.nn.SyncBatchNorm reached from /root/reference/main.py:433
// (torch/nn/modules/_functions.py:49-74 all_gather of [mean, invstd, count] in forward, :158-159 all_reduce of
// [sum_dy, sum_dy_xmu] in backward): ~110 latency-bound NCCL launches per step in round 1.
//
// Every rank owns one SYMMETRIC buffer (torch.distributed._symmetric_memory: the same allocation is mapped into every
// peer's address space over NVLink/NVSwitch) laid out as
//     uint32 flags[NSLOTS][MAXW]        rank r's arrival flag for exchange slot s (written BY rank r INTO my buffer)
//     byte   data [NSLOTS][cap_bytes]   my published partial sums for exchange slot s
// One exchange = ONE single-CTA kernel per rank:
//     1. publish: copy the local partial sums into my own data[slot]           (plain stores + system fence)
//     2. signal : thread r stores the sequence number into flags[slot][me] of PEER r's buffer (st.release.sys)
//     3. wait   : thread r spins on MY flags[slot][r] until it holds the sequence number (ld.acquire.sys)
//     4. reduce : every rank reads all peers' data[slot] over NVLink and adds them IN RANK ORDER
//                 -> the sums are bit-identical on every rank (replicas stay bit-identical, main.py:440), deterministic
// The sequence number lives in device memory and is advanced by the kernel itself, so a captured CUDA graph replays
// correctly (kernel arguments are frozen at capture).  Slot reuse is safe with NSLOTS >= 2: a rank can only pass the
// wait of exchange k after every rank has ENTERED exchange k, i.e. finished reading exchange k - 1.
#include "common.cuh"

namespace byol {

static constexpr int XCHG_SLOTS = 4;
static constexpr int XCHG_MAXW = 8;
static constexpr int XCHG_FLAG_BYTES = 1024;   // >= XCHG_SLOTS * XCHG_MAXW * 4, keeps the data region aligned

struct PeerPtrs { uint64_t p[XCHG_MAXW]; };

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(512)
xchg_sum_kernel(T* __restrict__ vals, T* __restrict__ local_copy, int n, PeerPtrs peers, int world, int rank,
                size_t cap_bytes, uint32_t* __restrict__ counter) {
  const uint32_t seq = *counter + 1u;          // same value on every rank: all ranks run the same exchange sequence
  const int slot = (int)(seq % XCHG_SLOTS);
  uint8_t* mine = reinterpret_cast<uint8_t*>(peers.p[rank]);
  T* my_data = reinterpret_cast<T*>(mine + XCHG_FLAG_BYTES + (size_t)slot * cap_bytes);
  // 1. publish
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T v = vals[i];
    my_data[i] = v;
    if (local_copy != nullptr) local_copy[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal every peer (including myself), 3. wait for every peer
  if ((int)threadIdx.x < world) {
    const int r = (int)threadIdx.x;
    uint32_t* peer_flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(peers.p[r]));
    st_release_sys(peer_flags + slot * XCHG_MAXW + rank, seq);
    const uint32_t* my_flag = reinterpret_cast<const uint32_t*>(mine) + slot * XCHG_MAXW + r;
    // bounded spin (~ seconds): a missing peer traps (launch failure) instead of hanging the GPU
    unsigned long long spins = 0;
    while (ld_acquire_sys(my_flag) != seq) {
      __nanosleep(64);
      if (++spins > (1ull << 24)) __trap();
    }
  }
  __syncthreads();
  // 4. reduce in rank order (peer loads bypass L1: the data was written by another GPU)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T acc = (T)0;
    for (int r = 0; r < world; ++r) {
      const T* pd = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(peers.p[r]) + XCHG_FLAG_BYTES +
                                               (size_t)slot * cap_bytes);
      acc += __ldcv(pd + i);
    }
    vals[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *counter = seq;
}

}  // namespace byol

using namespace byol;

extern "C" int byol_xchg_layout(int* slots, int* max_world, int* flag_bytes) {
  if (slots) *slots = XCHG_SLOTS;
  if (max_world) *max_world = XCHG_MAXW;
  if (flag_bytes) *flag_bytes = XCHG_FLAG_BYTES;
  return 0;
}

// vals: n values (fp32, or fp64 when is_f64) — in: this rank's partial sums, out: the sum over all ranks (identical
// bits on every rank); local_copy (optional): receives the input values.  peer_ptrs: host array of `world` device
// addresses of the ranks' symmetric buffers (flags + XCHG_SLOTS * cap_bytes of data each); counter: one zero-initialised
// device uint32 per rank.
extern "C" int byol_xchg_sum(void* vals, void* local_copy, int n, int is_f64, const uint64_t* peer_ptrs, int world,
                             int rank, int64_t cap_bytes, void* counter, cudaStream_t stream) {
  BYOL_CHECK_ARG(vals && peer_ptrs && counter && n > 0, "byol_xchg_sum: bad args");
  BYOL_CHECK_ARG(world >= 1 && world <= XCHG_MAXW && rank >= 0 && rank < world, "byol_xchg_sum: world=%d rank=%d", world, rank);
  BYOL_CHECK_ARG((int64_t)n * (is_f64 ? 8 : 4) <= cap_bytes, "byol_xchg_sum: %d values exceed the slot capacity", n);
  PeerPtrs pp;
  for (int r = 0; r < XCHG_MAXW; ++r) pp.p[r] = r < world ? peer_ptrs[r] : 0ull;
  if (is_f64)
    xchg_sum_kernel<double><<<1, 512, 0, stream>>>((double*)vals, (double*)local_copy, n, pp, world, rank,
                                                   (size_t)cap_bytes, (uint32_t*)counter);
  else
    xchg_sum_kernel<float><<<1, 512, 0, stream>>>((float*)vals, (float*)local_copy, n, pp, world, rank,
                                                  (size_t)cap_bytes, (uint32_t*)counter);
  return check_launch("xchg_sum_kernel");
}
