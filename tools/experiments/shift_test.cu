// Experiment: UMMA A operand read from a 128B-swizzled smem tile whose START is shifted by j*128 bytes
// (j rows), with descriptor base_offset = variant-dependent.  D[128 x 64] = A[rows j..j+127][64] * B[64][64]^T.
#include "../../byol_b200/csrc/common.cuh"
#include <string.h>
using namespace byol;
namespace byol { void set_last_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int j,
             int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 32768;         // 64 rows x 128 B
  uint64_t* bar = (uint64_t*)(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768 + 8192);
    tma_load_2d(smem_u32(sA), &tmA, bar, 0, 0);
    tma_load_2d(smem_u32(sA + 16384), &tmA, bar, 0, 128);
    tma_load_2d(smem_u32(sB), &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)j * 128u;
    uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
    if (variant == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;   // base_offset = (addr >> 7) & 7
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB), 16, 1024);
    constexpr uint32_t idesc = make_idesc(1u, 128, 64, 0u, 0u);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0);
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
extern "C" int shift_test(const void* A /*[256][64] bf16*/, const void* B /*[64][64] bf16*/, float* out, int j, int variant) {
  CUtensorMap ta, tb;
  if (mk(&ta, A, 256, 128) || mk(&tb, B, 64, 64)) return -1;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  shift_kernel<<<1, 128, 49152>>>(ta, tb, out, j, variant);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -3;
}
