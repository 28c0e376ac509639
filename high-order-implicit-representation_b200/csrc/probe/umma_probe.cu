// umma_probe.cu -- bring-up probe for tcgen05.mma kind::tf32 with thread-written, un-swizzled
// shared-memory operands (both majors), TMEM accumulators, M = 64 / 128, and the 3-pass hi/lo
// split.  Standalone (no torch): nvcc -gencode arch=compute_100a,code=sm_100a umma_probe.cu
// Usage: umma_probe <M> <N> <K> <a_mn> <b_mn> <swapA> <swapB> <split> <repeats>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

struct Cfg { int M, N, K, a_mn, b_mn, swapA, swapB, split, repeats, a_sw, b_sw; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int sw) {
  // sw: 0 none, 1 32B, 2 64B, 3 128B, 4 128B_base32B  -> descriptor layout_type 0, 6, 4, 2, 1
  const uint64_t lt = sw == 0 ? 0 : (sw == 1 ? 6 : (sw == 2 ? 4 : (sw == 3 ? 2 : 1)));
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= lt << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // D format F32
  d |= 2u << 7;                      // A format TF32
  d |= 2u << 10;                     // B format TF32
  d |= (uint32_t)(a_mn & 1) << 15;   // A major
  d |= (uint32_t)(b_mn & 1) << 16;   // B major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// operand element: e = index in the contiguous dim (size E elements of 4 B), r = index in the other dim.
// atom = 8 rows (4 for sw=4) of Wb bytes; atoms adjacent along e, then along r.
__device__ __host__ __forceinline__ uint32_t atom_w(int sw) { return sw == 0 ? 16u : (sw == 1 ? 32u : (sw == 2 ? 64u : 128u)); }
__device__ __host__ __forceinline__ uint32_t atom_rows(int sw) { return sw == 4 ? 4u : 8u; }
__device__ __host__ __forceinline__ uint32_t atom_e_stride(int sw) { return atom_w(sw) * atom_rows(sw); }
__device__ __host__ __forceinline__ uint32_t atom_r_stride(int sw, int E) { return (((uint32_t)E * 4u + atom_w(sw) - 1u) / atom_w(sw)) * atom_e_stride(sw); }
__device__ __host__ __forceinline__ uint32_t op_off(int e, int r, int E, int sw) {
  const uint32_t Wb = atom_w(sw), R = atom_rows(sw);
  const uint32_t eb = (uint32_t)e * 4u;
  uint32_t in_row = eb % Wb;
  const uint32_t rr = (uint32_t)r % R;
  if (sw >= 1 && sw <= 3) in_row ^= ((rr & (Wb / 16u - 1u)) << 4);
  if (sw == 4) in_row ^= ((rr & 3u) << 5);
  return rr * Wb + in_row + (eb / Wb) * atom_e_stride(sw) + ((uint32_t)r / R) * atom_r_stride(sw, E);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(Cfg c, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
             long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform -> uniform datapath
  const int M = c.M, N = c.N, K = c.K;
  const uint32_t a_bytes = ((uint32_t)(M + 32) * (K + 32) * 4 + 1023u) & ~1023u, b_bytes = ((uint32_t)(N + 32) * (K + 32) * 4 + 1023u) & ~1023u;
  uint8_t* sA = smem;
  uint8_t* sAl = sA + a_bytes;
  uint8_t* sB = sAl + a_bytes;
  uint8_t* sBl = sB + b_bytes;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // ---- write operands (logical A[M][K], B[K][N] row-major in global)
  for (int idx = tid; idx < M * K; idx += 128) {
    const int m = idx / K, k = idx % K;
    const float v = A[idx];
    const uint32_t off = c.a_mn ? op_off(m, k, M, c.a_sw) : op_off(k, m, K, c.a_sw);
    *reinterpret_cast<float*>(sA + off) = v;
    *reinterpret_cast<float*>(sAl + off) = tf32_lo(v);
  }
  for (int idx = tid; idx < K * N; idx += 128) {
    const int k = idx / N, n = idx % N;
    const float v = B[idx];
    const uint32_t off = c.b_mn ? op_off(n, k, N, c.b_sw) : op_off(k, n, K, c.b_sw);
    *reinterpret_cast<float*>(sB + off) = v;
    *reinterpret_cast<float*>(sBl + off) = tf32_lo(v);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const bool ts = c.a_sw == 9;
  if (ts) {   // A[M=128][K] row of this thread -> TMEM lane tid, columns 256.. (raw) and 256+K.. (lo)
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t v[8], l[8];
      for (int j = 0; j < 8; ++j) { const float a = A[(size_t)tid * K + k0 + j]; v[j] = __float_as_uint(a); l[j] = __float_as_uint(tf32_lo(a)); }
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)k0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(addr + (uint32_t)K), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- issue
  long long t0 = 0, t1 = 0;
  if (c.swapA == 8) {
    if (warp == 1) {   // warp-uniform issue loop: descriptors live in uniform registers
      const uint32_t idesc = make_idesc(M, N, 0, 0);
      const uint32_t lbo = 128, sbo = (uint32_t)(K / 4) * 128;
      const uint64_t a0 = make_desc(smem_u32(sA), lbo, sbo, 0), b0 = make_desc(smem_u32(sB), lbo, sbo, 0);
      const uint32_t nacc = c.swapB > 0 ? c.swapB : 1;
      t0 = clock64();
      const uint32_t leader = elect_one();
      for (int rep = 0; rep < c.repeats; ++rep) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t dcol = ((uint32_t)ks & (nacc - 1)) * 32u;
          if (leader) {
            if (c.a_sw == 9) mma_tf32_ts(tmem + dcol, tmem + 256u + ks * 8u, b0 + (uint64_t)(ks * 16), idesc, 1);
            else mma_tf32(tmem + dcol, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, 1);
          }
        }
      }
      __syncwarp();
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
      __syncwarp();
    }
  } else if (tid == 0) {
    // K-major : SBO = stride between 8-row groups; LBO = stride between 16B K-chunks (no swizzle only)
    // MN-major: no swizzle: SBO = stride between 16B MN-chunks, LBO = stride between 8-k groups;
    //           swizzled : LBO = stride between MN atoms,       SBO = stride between 8-k (4-k) groups
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    if (!c.a_mn) { a_lbo = atom_e_stride(c.a_sw); a_sbo = atom_r_stride(c.a_sw, K); }
    else if (c.a_sw == 0) { a_sbo = atom_e_stride(0); a_lbo = atom_r_stride(0, M); }
    else { a_lbo = atom_e_stride(c.a_sw); a_sbo = atom_r_stride(c.a_sw, M); }
    if (!c.b_mn) { b_lbo = atom_e_stride(c.b_sw); b_sbo = atom_r_stride(c.b_sw, K); }
    else if (c.b_sw == 0) { b_sbo = atom_e_stride(0); b_lbo = atom_r_stride(0, N); }
    else { b_lbo = atom_e_stride(c.b_sw); b_sbo = atom_r_stride(c.b_sw, N); }
    if (c.swapA) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
    if (c.swapB) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t idesc = make_idesc(M, N, c.a_mn, c.b_mn);
    const int passes = c.split ? 3 : 1;
    uint64_t ad[3][16], bd[3][16];   // descriptors precomputed: the timed loop only issues
    const int ksteps = K / 8;
    for (int p = 0; p < passes; ++p) {
      const uint32_t abase = smem_u32(p == 1 ? sAl : sA), bbase = smem_u32(p == 2 ? sBl : sB);
      for (int ks = 0; ks < ksteps && ks < 16; ++ks) {
        const uint32_t kb = ks * 32u;
        const uint32_t a_adv = c.a_mn ? ks * atom_r_stride(c.a_sw, M) * (c.a_sw == 4 ? 2u : 1u)
                                      : (kb / atom_w(c.a_sw)) * atom_e_stride(c.a_sw) + kb % atom_w(c.a_sw);
        const uint32_t b_adv = c.b_mn ? ks * atom_r_stride(c.b_sw, N) * (c.b_sw == 4 ? 2u : 1u)
                                      : (kb / atom_w(c.b_sw)) * atom_e_stride(c.b_sw) + kb % atom_w(c.b_sw);
        ad[p][ks] = make_desc(abase + a_adv, a_lbo, a_sbo, c.a_sw);
        bd[p][ks] = make_desc(bbase + b_adv, b_lbo, b_sbo, c.b_sw);
      }
    }
    if (c.swapA == 7) {   // timing mode: fully unrolled issue, descriptors advanced with one add
      const uint64_t a0 = ad[0][0], b0 = bd[0][0];
      const uint64_t da = ksteps > 1 ? ad[0][1] - ad[0][0] : 0, db = ksteps > 1 ? bd[0][1] - bd[0][0] : 0;
      mma_tf32(tmem, a0, b0, idesc, 0);
      t0 = clock64();
      for (int rep = 0; rep < c.repeats; ++rep) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t dcol = (uint32_t)(ks % (c.swapB > 0 ? c.swapB : 1)) * 32u;   // swapB = number of accumulators
          if (ts) mma_tf32_ts(tmem + dcol, tmem + 256u + ks * 8u, b0 + ks * db, idesc, 1);
          else mma_tf32(tmem + dcol, a0 + ks * da, b0 + ks * db, idesc, 1);
        }
      }
    } else {
    t0 = clock64();
    for (int rep = 0; rep < c.repeats; ++rep) {
      for (int p = 0; p < passes; ++p) {
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) {
          if (ts) mma_tf32_ts(tmem, tmem + 256u + (p == 1 ? (uint32_t)K : 0u) + ks * 8u, bd[p][ks], idesc, (rep | p | ks) != 0);
          else mma_tf32(tmem, ad[p][ks], bd[p][ks], idesc, (rep | p | ks) != 0);
        }
      }
    }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  // ---- wait
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  if (tid == (c.swapA == 8 ? 32 : 0)) { t1 = clock64(); cycles[0] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- read D
  int row = -1;
  if (M == 128) row = tid;
  else if (lane < 16) row = warp * 16 + lane;   // M = 64: 16 lanes of each 32-lane sub-partition
  for (int n0 = 0; n0 < N; n0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row >= 0)
      for (int j = 0; j < 8; ++j) D[(size_t)row * N + n0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 10) { printf("usage: M N K a_mn b_mn swapA swapB split repeats [a_sw b_sw]\n"); return 1; }
  Cfg c = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]), argc > 10 ? atoi(argv[10]) : 0, argc > 11 ? atoi(argv[11]) : 0};
  const int M = c.M, N = c.N, K = c.K;
  std::vector<float> A((size_t)M * K), B((size_t)K * N), D((size_t)M * N, -777.f);
  srand(1234);
  for (auto& v : A) v = c.split ? (float)rand() / RAND_MAX * 2.f - 1.f : (float)(rand() % 9 - 4);
  for (auto& v : B) v = c.split ? (float)rand() / RAND_MAX * 2.f - 1.f : (float)(rand() % 7 - 3);
  float *dA, *dB, *dD; long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)((M + 32) * (K + 32) + (N + 32) * (K + 32)) * 8 + 8192;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(c, dA, dB, dD, dC);
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)k * N + n];
      ref *= c.repeats;
      const double err = fabs((double)D[(size_t)m * N + n] - ref);
      if (err > max_err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (err > 1e-3 * (1 + fabs(ref)) && bad < 6) { printf("  mismatch D[%d][%d] = %g, want %g\n", m, n, D[(size_t)m * N + n], ref); ++bad; }
    }
  const int mmas = c.swapA >= 7 ? c.repeats * 8 : c.repeats * (c.split ? 3 : 1) * (K / 8);
  printf("M=%d N=%d K=%d a_mn=%d b_mn=%d swapA=%d swapB=%d split=%d rep=%d a_sw=%d b_sw=%d : max_err %.3e (rel %.3e) %s | %lld cycles, %.1f cyc/MMA\n",
         M, N, K, c.a_mn, c.b_mn, c.swapA, c.swapB, c.split, c.repeats, c.a_sw, c.b_sw, max_err, max_err / (max_ref + 1e-30),
         max_err <= 1e-3 * (1 + max_ref) * (c.split ? 0.01 : 1) ? "OK" : "WRONG", cyc, (double)cyc / mmas);
  return 0;
}
