// byol_b200 — sm_100a device helpers: mbarrier, TMA, tcgen05/TMEM, cp.async.
// Hand-written PTX wrappers; no CUTLASS/CuTe dependency.  Encodings follow the
// PTX ISA (tcgen05 shared-memory / instruction descriptors).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace byol {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------
// error plumbing shared by all host wrappers (thread-local last error string)
// ----------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);

// One process may drive several GPUs: per-kernel attributes (cudaFuncSetAttribute) and the SM count are per device,
// so the host wrappers cache them per device slot (capi.cu).
static constexpr int kMaxDevices = 64;
int device_slot();       // current CUDA device index, clamped to [0, kMaxDevices)
int device_sm_count();   // SM count of the current device (cached per device; 148 if the query fails)

#define BYOL_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::byol::set_last_error(__VA_ARGS__);        \
      return -1;                                  \
    }                                             \
  } while (0)

// ----------------------------------------------------------------------------
// small device utilities
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a pipeline bug traps (reported as a launch failure) instead of
// hanging the GPU.  try_wait suspends in hardware, so the bound is generous.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // (2^28 polls: waits that span a grid barrier plus a cross-rank exchange — the MMA warp of mlp_fused_fwd_kernel under
  // SyncBatchNorm — can legitimately last as long as the rank skew, e.g. right after the CUDA-graph capture)
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}

// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (UMMA / TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte with zero-fill
// ----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, bool valid) {
  uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiled, arrives on an mbarrier with complete_tx
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA store smem -> global (bulk async group), 2-D tiled
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// same, but allow the most recent group to be still in flight (double-buffered staging)
__device__ __forceinline__ void tma_store_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate (kind::f16)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Warp-convergent variants: the WHOLE warp executes the call with warp-uniform operands and one elected lane issues.
// Keeping the issuing warp's control flow convergent lets the compiler hold descriptors in uniform registers; inside
// an `if (lane == 0)` region it wraps every tcgen05.mma in a uniformisation loop (ELECT / R2UR.BROADCAST / BRA.U.ANY,
// ~100 cycles per MMA), which is the limit for narrow (N <= 64) MMAs.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// descriptors
// ----------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle, version 1 (sm_100).
//   start address  bits [0,14)   (>>4)
//   LBO            bits [16,30)  (>>4)
//   SBO            bits [32,46)  (>>4)
//   version = 1    bits [46,48)
//   layout  = 2    bits [61,64)  (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
//   c_format  bits [4,6)   1 = F32
//   a_format  bits [7,10)  0 = F16, 1 = BF16, 2 = TF32
//   b_format  bits [10,13)
//   a_major   bit 15       0 = K-major, 1 = MN-major
//   b_major   bit 16
//   N >> 3    bits [17,23)
//   M >> 4    bits [24,29)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// physical byte offset of (row, 16-byte chunk) inside a [rows][128 B] tile with
// the 128-byte swizzle (tile base must be 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ----------------------------------------------------------------------------
// 1-D bulk copy global -> shared (any multiple of 16 bytes), completes on an mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// Shared-memory matrix descriptor without swizzle: core matrix = 8 rows x 16 bytes, rows 16 bytes apart;
// K-major: LBO = byte step between the two 16-byte K chunks of one K = 16 MMA, SBO = step between 8-row groups.
// (Measured, tools/experiments/noswz_test.cu: LBO / SBO may be smaller than the 128-byte core matrix, i.e. rows and
// chunks may overlap: the descriptor alone forms the im2col of a 1-D window over 16-byte pixels.)
__device__ __forceinline__ uint64_t make_smem_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ----------------------------------------------------------------------------
// epilogue statistics (shared by the igemm and stem kernels)
// ----------------------------------------------------------------------------
// packed fp32x2 helpers (FADD2 / FFMA2 on sm_100): the statistics loops do two columns per instruction
__device__ __forceinline__ uint64_t f2_pack(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(uint64_t v) {
  uint32_t lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  return make_float2(__uint_as_float(lo), __uint_as_float(hi));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t w;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr));
  return w;
}
// one staged word = two bf16 columns -> {sum, sum of squares} accumulators
__device__ __forceinline__ void stat_acc(uint32_t w, uint64_t& s1, uint64_t& s2) {
  const uint64_t x = f2_pack(w << 16, w & 0xffff0000u);
  s1 = f2_add(s1, x);
  s2 = f2_fma(x, x, s2);
}

// Column statistics of a staged narrow chunk ([32 rows][64 B], 16-byte chunk j of row r at j ^ ((r >> 1) & 3)).
// lane -> column pair cp = lane & 15 of rows with parity lane >> 4 (two adjacent rows per warp-wide load: all 32
// banks, no conflicts); the caller combines the two parities with one shuffle when it flushes.
__device__ __forceinline__ void stats_narrow(uint32_t stage_base, int lane, int rows_valid, uint64_t& s1,
                                             uint64_t& s2) {
  const uint32_t cp = (uint32_t)lane & 15u, rh = (uint32_t)lane >> 4;
  const uint32_t jc = cp >> 2, wq = cp & 3u;
  uint32_t offq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) offq[q] = stage_base + rh * 64u + ((jc ^ (uint32_t)q) << 4) + wq * 4u;
  uint64_t a1 = 0ull, a2 = 0ull, b1 = 0ull, b2 = 0ull;
  if (rows_valid == 32) {
#pragma unroll
    for (int rr = 0; rr < 16; rr += 2) {   // row = 2*rr + rh, so (row >> 1) & 3 == rr & 3
      stat_acc(lds32(offq[rr & 3] + (uint32_t)rr * 128u), a1, a2);
      stat_acc(lds32(offq[(rr + 1) & 3] + (uint32_t)(rr + 1) * 128u), b1, b2);
    }
  } else {
    for (int rr = 0; rr < 16; ++rr)
      if (2 * rr + (int)rh < rows_valid) stat_acc(lds32(offq[rr & 3] + (uint32_t)rr * 128u), a1, a2);
  }
  s1 = f2_add(s1, f2_add(a1, b1));
  s2 = f2_add(s2, f2_add(a2, b2));
}


}  // namespace byol
