// Experiment: UMMA A operand read from a LINEAR (no-swizzle) smem buffer with OVERLAPPING rows:
//   A'[m][8*c + e] = buf[(m + c) * 8 + e]   (row m, 16-byte K-chunk c starts 16 bytes after chunk c-1 of the same row
//   = chunk c of row m is chunk c-1 of row m+1): the im2col of a stride-1 window over 16-byte pixels, formed by the
//   descriptor alone.  Core matrix = 8 rows x 16 B contiguous (rows 16 B apart); variant 0: LBO = 16 (K step),
//   SBO = 128 (8-row step); variant 1: the two swapped.
// D[128 x 64] = A'[128 x 64] * B[64 x 64]^T, B via TMA (128-byte swizzle).
#include "../../byol_b200/csrc/common.cuh"
#include <string.h>
using namespace byol;
namespace byol { void set_last_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ uint64_t make_smem_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;   // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128, 1)
noswz_kernel(const bf16* __restrict__ buf /*[256*8]*/, const __grid_constant__ CUtensorMap tmB, float* out, int shift,
             int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 256 pixels x 16 B
  uint8_t* sB = smem + 8192;          // 64 rows x 128 B
  uint64_t* bar = (uint64_t*)(smem + 8192 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 256; i += 128)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(buf)[i];
  fence_proxy_async_smem();
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 8192);
    tma_load_2d(smem_u32(sB), &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 16u;
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
    constexpr uint32_t idesc = make_idesc(1u, 128, 64, 0u, 0u);
    for (int k = 0; k < 4; ++k) {
      // K = 16 per MMA = chunks 2k, 2k+1: start advances by 2 pixels (32 B)
      const uint64_t adesc = variant == 0 ? make_smem_desc_none(a_addr + 32u * k, 16, 128)
                                          : make_smem_desc_none(a_addr + 32u * k, 128, 16);
      umma_bf16(tmem, adesc, bdesc + (uint64_t)(2 * k), idesc, k != 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int q = 0; q < 32; ++q) out[(warp * 32 + lane) * 64 + c0 + q] = __uint_as_float(r[q]);
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int mk(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows) {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) return -1;
  cuuint64_t dims[2] = {64, rows}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, box_rows}; cuuint32_t es[2] = {1, 1};
  return ((PFN)ptr)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}
extern "C" int noswz_test(const void* buf /*[256][8] bf16*/, const void* B /*[64][64] bf16*/, float* out, int shift,
                          int variant) {
  CUtensorMap tb;
  if (mk(&tb, B, 64, 64)) return -1;
  cudaFuncSetAttribute(noswz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  noswz_kernel<<<1, 128, 32768>>>((const bf16*)buf, tb, out, shift, variant);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
}
