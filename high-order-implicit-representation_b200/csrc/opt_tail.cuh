// opt_tail.cuh -- the tail of a training step, fused: optimizer update (torch.optim.Adam /
// SparseLion semantics, /root/reference/high_order_implicit_representation/networks.py:93-102) on the
// flat parameter vector, refresh of the kernel-side packed weights [in][node][out_pad], zeroing of the
// flat gradient for the next step and hand-over of the loss -- one pass over the parameters.
//
// It runs either (a) inside the fused whole-network kernels, after a grid-wide barrier that follows
// the gradient flush (single GPU, coherent batches: ONE launch per training step), or (b) as the
// stand-alone opt_tail_kernel after the NCCL all-reduce / the second pass of the layer-0 gradient.
// Flat walk over the boundary-layout parameters: coalesced reads, scattered packed writes.
#pragma once
#include "hoir_common.cuh"

struct OptTail {
  int kind;                               // 0: none, HOIR_OPT_ADAM, HOIR_OPT_SPARSE_LION
  int keep_grad;                          // 1: do not zero the gradient / loss accumulator
  int num_layers;
  int in_f[HOIR_MAX_LAYERS], nodes[HOIR_MAX_LAYERS], out_pad[HOIR_MAX_LAYERS];
  int row0[HOIR_MAX_LAYERS + 1];          // prefix sums of out*in links
  long long woff[HOIR_MAX_LAYERS + 1];    // boundary-layout float offsets; woff[L] = number of parameters
  long long poff[HOIR_MAX_LAYERS];        // packed float offsets
  long long alt_off;                      // >= 0: second copy of the layer-0 weights, [in][half][node][alt_hh]
  int alt_hh;
  // layers whose kernel-side copy is a pre-swizzled hi | lo chunk image [chunk][hi | lo][outp x 32] (tcgen05 B operand of
  // the fused Fourier kernel) instead of the plain [in][node][out_pad] table: bit l of img_mask
  int img_mask;
  int img_nd[HOIR_MAX_LAYERS], img_ipc[HOIR_MAX_LAYERS], img_outp[HOIR_MAX_LAYERS];
  float* w;
  float* g;                               // [n + 1]: gradients, then the loss accumulator
  float* m;
  float* v;
  float* packed;
  float* loss_out;
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt;
  unsigned int* sync;                     // [2] arrive / depart counters of the in-kernel grid barrier
  // peer exchange (world > 1): the gradient is the rank-ordered sum over every rank's symmetric buffer
  int world, rank;
  const float* peer_g[HOIR_MAX_PEERS];    // this step's gradient buffer of every rank (g == peer_g[rank])
  unsigned int* peer_flag[HOIR_MAX_PEERS];   // flag array of every rank; slot [rank] is ours to write
  unsigned int flag_value;                // the step number
  float* zero_buf;                        // own other-parity buffer, zeroed for the step after next
};

namespace hoir {

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier for a grid whose CTAs are all resident (cooperative launch).  Every thread of
// every CTA must call it; writes (including atomics) made before it are visible to L2 reads after it.
__device__ __forceinline__ void grid_barrier_arrive_wait(unsigned int* sync) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&sync[0], 1u);
    unsigned int spins = 0;
    while (ld_acquire_u32(&sync[0]) < gridDim.x) {
      __nanosleep(64);
      if (++spins > (1u << 26)) __trap();   // a CTA never arrived: fail loudly instead of hanging
    }
    __threadfence();
  }
  __syncthreads();
}
// the last CTA to leave re-arms the counters for the next launch
__device__ __forceinline__ void grid_barrier_depart(unsigned int* sync) {
  if (threadIdx.x == 0) {
    if (atomicAdd(&sync[1], 1u) == gridDim.x - 1) {
      sync[0] = 0u;
      sync[1] = 0u;
      __threadfence();
    }
  }
}

// Barrier number `phase` (1, 2, ...) of a kernel that synchronises its (fully resident) grid several times on ONE counter
// that only counts up; the last CTA to leave the kernel re-arms it with grid_sync_rearm.
__device__ __forceinline__ void grid_sync_phase(unsigned int* counter, unsigned int phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = phase * gridDim.x;
    unsigned int spins = 0;
    while (ld_acquire_u32(counter) < target) {
      __nanosleep(32);
      if (++spins > (1u << 26)) __trap();   // a CTA never arrived: fail loudly instead of hanging
    }
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_sync_rearm(unsigned int* counter, unsigned int* depart) {
  if (threadIdx.x == 0 && atomicAdd(depart, 1u) == gridDim.x - 1) {
    *counter = 0u;
    *depart = 0u;
    __threadfence();
  }
}

__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {   // peer memory over NVLink: never from L1
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// Cross-GPU rendezvous of one step.  Precondition: every gradient add of THIS GPU is complete and visible
// to the calling CTA 0 (grid barrier, or kernel boundary).  CTA 0 publishes the flag to every rank; every CTA
// waits until all ranks have published theirs.
__device__ __forceinline__ void peer_exchange_barrier(const OptTail& t) {
  if (blockIdx.x == 0 && threadIdx.x < t.world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.peer_flag[threadIdx.x] + t.rank), "r"(t.flag_value) : "memory");
  }
  if (threadIdx.x < t.world) {
    const unsigned int* mine = t.peer_flag[t.rank] + threadIdx.x;
    unsigned int spins = 0;
    while (ld_acquire_sys_u32(mine) < t.flag_value) {
      __nanosleep(64);
      if (++spins > (1u << 30)) __trap();   // a peer never arrived (~1 min): fail loudly instead of hanging
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void opt_tail_rows(const OptTail& t, int gwarp, int nwarps, int lane) {
  const float one_m_b1 = 1.f - t.b1, one_m_b2 = 1.f - t.b2, step_size = t.lr / t.bc1;
  const long long n = t.woff[t.num_layers];
  const long long stride = (long long)nwarps * 32;
  // flat walk over the boundary-layout parameters (coalesced); iterations are independent, so the
  // unrolled loads overlap -- a small grid (batch 256: one CTA) is latency-, not bandwidth-bound here
#pragma unroll 4
  for (long long k = (long long)gwarp * 32 + lane; k < n; k += stride) {
    int l = 0;
    while (l + 1 < t.num_layers && k >= t.woff[l + 1]) ++l;
    const int e = (int)(k - t.woff[l]);
    const int nodes = t.nodes[l], in_f = t.in_f[l], op = t.out_pad[l];
    const int r = e / nodes, q = e - r * nodes;
    const int o = r / in_f, i = r - o * in_f;
    float gk;
    if (t.world > 1) {                    // rank-ordered sum over the peers' buffers: identical on every rank
      gk = 0.f;
      for (int r = 0; r < t.world; ++r) gk += ld_peer_f32(t.peer_g[r] + k);
      t.zero_buf[k] = 0.f;
    } else {
      gk = __ldcg(t.g + k);               // written by other SMs' atomics: read at L2
    }
    float wk = t.w[k];
    if (t.kind == HOIR_OPT_ADAM) {
      if (t.wd != 0.f) gk = fmaf(t.wd, wk, gk);
      const float m0 = t.m[k];
      const float mk = m0 + one_m_b1 * (gk - m0);          // lerp, as torch does
      const float vk = t.b2 * t.v[k] + one_m_b2 * gk * gk;
      t.m[k] = mk;
      t.v[k] = vk;
      const float denom = sqrtf(vk) / t.bc2_sqrt + t.eps;
      wk = wk - step_size * (mk / denom);
      t.w[k] = wk;
    } else if (gk != 0.f) {               // SparseLion: only entries with a non-zero gradient
      const float ea = t.m[k];
      const float wd = wk * (1.f - t.lr * t.wd);
      const float u = ea * t.b1 + gk * one_m_b1;
      const float sg = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
      wk = wd - t.lr * sg;
      t.w[k] = wk;
      t.m[k] = ea * t.b2 + gk * one_m_b2;
    }
    if (!t.keep_grad && t.world <= 1) t.g[k] = 0.f;
    if ((t.img_mask >> l) & 1) {
      const int c = i / t.img_ipc[l], col = (i - c * t.img_ipc[l]) * t.img_nd[l] + q, outp = t.img_outp[l];
      float* chunk = t.packed + t.poff[l] + (long long)c * 2 * outp * 32;
      const unsigned int off = (((unsigned)o >> 2) * 512u + ((unsigned)o & 3u) * 128u +
                                (((unsigned)col << 2) ^ (((unsigned)o & 3u) << 5))) >> 2;
      chunk[off] = wk;
      chunk[outp * 32 + off] = wk - __uint_as_float(__float_as_uint(wk) & 0xffffe000u);   // tf32 residual
    } else {
      t.packed[t.poff[l] + ((long long)i * nodes + q) * op + o] = wk;
    }
    if (l == 0 && t.alt_off >= 0) {
      const int half = o >= t.alt_hh;
      t.packed[t.alt_off + ((long long)(i * 2 + half) * nodes + q) * t.alt_hh + (o - half * t.alt_hh)] = wk;
    }
  }
  if (gwarp == 0 && lane == 0) {
    if (t.world > 1) {
      float ls = 0.f;
      for (int r = 0; r < t.world; ++r) ls += ld_peer_f32(t.peer_g[r] + n);
      if (t.loss_out != nullptr) t.loss_out[0] = ls;
      t.zero_buf[n] = 0.f;
    } else {
      if (t.loss_out != nullptr) t.loss_out[0] = __ldcg(t.g + n);
      if (!t.keep_grad) t.g[n] = 0.f;
    }
  }
}

// called by every thread of a fused kernel at its very end
__device__ __forceinline__ void opt_tail_in_kernel(const OptTail& t) {
  if (t.kind == 0) return;
  grid_barrier_arrive_wait(t.sync);
  if (t.world > 1) peer_exchange_barrier(t);
  const int wpb = blockDim.x >> 5;
  opt_tail_rows(t, blockIdx.x * wpb + (threadIdx.x >> 5), gridDim.x * wpb, threadIdx.x & 31);
  grid_barrier_depart(t.sync);
}

}  // namespace hoir

// fused FORWARD of the neighborhood-interpolator family (fused_interp.cu): hoir_mlp_forward / hoir_interp_frame
bool hoir_interp_supported(const hoir_mlp_desc* m);
long long hoir_interp_packed_offset(const hoir_mlp_desc* m, int layer);      // layer == num_layers: total floats
int hoir_interp_pack(const hoir_mlp_desc* m, const float* const* w, float* packed, cudaStream_t st);
int hoir_interp_forward(const hoir_mlp_desc* m, long long B, const float* x, const float* packed, float* y, int post_scale,
                        float* acts, cudaStream_t st);      // acts: [layer - 1][pre-norm | normalized][B][hid] or nullptr

// fused Fourier-network family (fused_fourier.cu), dispatched from the C ABI entry points of fused_mlp.cu
bool hoir_fourier_supported(const hoir_mlp_desc* m);
long long hoir_fourier_packed_offset(const hoir_mlp_desc* m, int layer);     // layer == 3: total floats
int hoir_fourier_pack(const hoir_mlp_desc* m, const float* const* w, float* packed, cudaStream_t st);
void hoir_fourier_fill_tail(const hoir_mlp_desc* m, OptTail* t);
int hoir_fourier_forward(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                         const float* packed, float* y, int post_scale, cudaStream_t st);
int hoir_fourier_train(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh, int target_kind,
                       const void* target, const float* gy_ext, const float* packed, float* const* gw, float* loss_sum,
                       float loss_scale, const OptTail* tail, cudaStream_t st, const char* where);
