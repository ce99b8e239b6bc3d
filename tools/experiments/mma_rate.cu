// Experiment (next round): issue-rate / shared-memory ceiling of SS-mode tcgen05.mma for M = 128, N in {64, 128, 256}.
// One CTA per SM (or two with co = 2) loops over fixed smem operands (128B-swizzled K-major tiles of zeros), 4 MMAs
// (K = 64) per "k-block", one commit per k-block, and reports cycles per MMA measured with clock64 around the loop.
// Question it answers: does a 128x256 tile lift the ~55 % tensor-pipe ceiling the 128x128 kernels show
// (8 KB of smem operand reads per 64 tensor cycles at N = 128 vs 12 KB per 128 cycles at N = 256)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -o mma_rate.so mma_rate.cu -lcuda
//   python tools/experiments/mma_rate.py
#include "../../byol_b200/csrc/common.cuh"
using namespace byol;
namespace byol { void set_last_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

template <int N>
__global__ void __launch_bounds__(128, 2)
mma_rate_kernel(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128, STAGES = 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bar = (uint64_t*)(sB + STAGES * B_BYTES);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (STAGES * (A_BYTES + B_BYTES)) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, N <= 128 ? 256 : 512); tmem_relinquish(); }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(1u, 128, N, 0u, 0u);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it & 1;
      const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + s * A_BYTES), 16, 1024);
      const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + s * B_BYTES), 16, 1024);
      const uint32_t d = tmem + (uint32_t)((it & 1) * N) % (N <= 128 ? 256u : 512u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_elect(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
    }
    umma_commit_elect(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, N <= 128 ? 256 : 512); }
}

template <int N>
static int run(long long* out, int iters, int ctas) {
  const int smem = 2 * (128 * 128 + N * 128) + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<N><<<ctas, 128, smem>>>(out, iters);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
}

// out: device int64[ctas]; returns 0 on success.  ctas = 148 (one per SM) or 296 (two per SM; N <= 128 only: TMEM)
extern "C" int mma_rate(long long* out, int n, int iters, int ctas) {
  if (n == 64) return run<64>(out, iters, ctas);
  if (n == 128) return run<128>(out, iters, ctas);
  if (n == 256) return run<256>(out, iters, ctas);
  return -1;
}
