// tmem_align_probe.cu -- which column alignments do tcgen05.st / tcgen05.ld .32x32b.xN accept?
// Writes a known pattern with .xN at column offset `off` and reads every column back with .x1.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(int* result) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t lane_base = tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  int bad = 0;
  for (int off = 0; off < 16; ++off) {
    // zero 32 columns
    for (int c = 0; c < 32; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const uint32_t v0 = 1000u * off + threadIdx.x, v1 = v0 + 100000u, v2 = v0 + 200000u, v3 = v0 + 300000u;
    // x2 at off, x4 at off + 8
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(lane_base + off), "r"(v0), "r"(v1) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_base + off + 8), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[c]) : "r"(lane_base + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const bool ok2 = r[off] == v0 && r[off + 1] == v1;
    const bool ok4 = r[off + 8] == v0 && r[off + 9] == v1 && r[off + 10] == v2 && r[off + 11] == v3;
    // and vector loads at the same offsets
    uint32_t a0, a1, b0, b1, b2, b3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(lane_base + off));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(lane_base + off + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const bool okl2 = a0 == v0 && a1 == v1, okl4 = b0 == v0 && b1 == v1 && b2 == v2 && b3 == v3;
    if (threadIdx.x == 5) result[off] = (ok2 ? 1 : 0) | (ok4 ? 2 : 0) | (okl2 ? 4 : 0) | (okl4 ? 8 : 0);
    bad += !(ok2 && ok4 && okl2 && okl4);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64));
  if (threadIdx.x == 5) result[16] = bad;
}
int main() {
  int* d; cudaMalloc(&d, 17 * sizeof(int)); cudaMemset(d, 0xff, 17 * sizeof(int));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  int h[17]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  for (int off = 0; off < 16; ++off) printf("col offset %2d: st.x2 %d  st.x4(+8) %d  ld.x2 %d  ld.x4(+8) %d\n", off, h[off] & 1, (h[off] >> 1) & 1, (h[off] >> 2) & 1, (h[off] >> 3) & 1);
  printf("bad (thread 5) %d\n", h[16]);
  return 0;
}
