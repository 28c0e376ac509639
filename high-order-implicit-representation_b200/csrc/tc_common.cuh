// tc_common.cuh -- tcgen05 / TMEM / mbarrier plumbing shared by the tensor-core kernels (sm_100a inline PTX).
// Shared-memory operands use the SWIZZLE_128B_BASE32B layout (rows of 128 B, atoms of 4 rows, 32-byte chunks
// XOR-ed with row % 4): the one layout tcgen05 accepts for tf32 in BOTH majors (csrc/probe/umma_probe.cu).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <cuda_runtime.h>

// -DHOIR_DEBUG_BOUNDS (make debug -> libhoir_b200_dbg.so): asserts on the swizzled shared-memory offsets and the TMEM
// column arithmetic of the tensor-core kernels.  compute-sanitizer is not available on the GPU pool, so a debug build
// of the same sources, run by tests/test_gpu_debug_bounds.py, is what guards that addressing.
#ifdef HOIR_DEBUG_BOUNDS
#define HOIR_DBG_ASSERT(cond)                                                                  \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      printf("HOIR_DEBUG_BOUNDS violated at %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
      __trap();                                                                                \
    }                                                                                          \
  } while (0)
#else
#define HOIR_DBG_ASSERT(cond) ((void)0)
#endif

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // protocol bug: fail loudly instead of hanging
  }
}
// Same, for waits that are long and off the critical path (helper warps, the MMA warp between phases): back off
// with nanosleep so that the spinning warp does not take issue slots from the warps doing the arithmetic
// (measured: the tight spin was 18 % of all issued instructions of the training kernel).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(ns);
    if (++spins > (1u << 27)) __trap();
  }
}
// non-blocking poll of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// (fused_mlp_tc) all 256 sample threads
__device__ __forceinline__ void bar_samples() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// the two warps (w, w + 4) that share 32 samples; immediate barrier ids so that ptxas counts 6 named
// barriers instead of reserving all 16 (which would cap a kernel at one CTA per SM)
__device__ __forceinline__ void bar_pair(int warp) {
  switch (warp & 3) {
    case 0: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
  }
}

// shared-memory operand in the 128B_BASE32B layout: element (row r, column c), RS = bytes per 4 rows
__device__ __forceinline__ uint32_t sw4_off(uint32_t r, uint32_t c, uint32_t RS) {
  HOIR_DBG_ASSERT((c >> 5) * 512u + 512u <= RS);      // the column atom lies inside one 4-row group
  return (r & 3u) * 128u + ((((c & 31u) << 2)) ^ ((r & 3u) << 5)) + (c >> 5) * 512u + (r >> 2) * RS;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)(512u >> 4) << 16;                 // LBO: next 32-column atom
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;      // SBO: next 4-row atom
  d |= (uint64_t)1 << 46;                           // descriptor version
  d |= (uint64_t)1 << 61;                           // layout type 1: SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 8u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 4u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 4u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 8u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// 20 consecutive columns (one half of a 40-wide row); column offsets stay multiples of 4
__device__ __forceinline__ void tmem_st20(uint32_t addr, const float* v) {
#pragma unroll
  for (int q = 0; q < 5; ++q) tmem_st4(addr + 4 * q, v + 4 * q);
}
__device__ __forceinline__ void tmem_ld20(uint32_t addr, float* v) {
#pragma unroll
  for (int q = 0; q < 5; ++q) tmem_ld4(addr + 4 * q, v + 4 * q);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// TMA-style 1-D bulk copy global -> shared, completion counted on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t addr, const float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 32u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                 "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
                 "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
                 "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
                 "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31])) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
  HOIR_DBG_ASSERT((addr & 0xffffu) + 16u <= 512u && (addr >> 16) < 128u);   // columns / lane base inside TMEM
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

}  // namespace tc
