// gw0_tc.cu -- dL/dw of a many-segment piecewise layer (the input layer of the neighborhood interpolator: 216 inputs x
// 51 nodes x 50 outputs, /root/reference/examples/implicit_neighborhood.py:37-41) on the tensor cores.
//
//   dL/dw[o][i][node] = sum_b dL/dy[b][o] * Phi_i[b][node],   Phi_i[b][.] = the N basis values of sample b at its segment's
//                                                              nodes, zero elsewhere (a one-hot-like row)
//
// slab_gw_kernel (slab_kernels.cu) scatters: a shared-memory read-modify-write of N rows of `out` floats per
// (sample, input) -- 12.4 shared-memory wavefronts per pair with the data pipe 76 % busy; shared-memory float atomics
// are a CAS loop, so that is its floor.  Here the sum over samples is a tcgen05 contraction with NO software
// read-modify-write.  Per CTA (8 inputs = 2 blocks of 4, a slice of the batch), per unit of 64 samples:
//   A  (smem, K-major)  = (dL/dy)^T as fp16: rows 0..63 = hi halves of the outputs, rows 64..127 = lo halves (v = hi + lo
//                         carries 22 bits, like tf32 x 3; dL/dy is first scaled by a power of two so that its largest
//                         entry sits near 2^14 -- gradients of 1e-7 would be fp16 subnormals -- a 10 us max-|.| pre-pass);
//                         column = sample.  Staged once per unit, shared by all inputs of the CTA.
//   B  (smem, K-major)  = Phi^T of a block as fp16: row = (input of the block, node slot) -- 4 x 64 = 256 rows --, column =
//                         sample; a hi and a lo tile.  The tiles stay zero except for the N cells per (sample, input) that
//                         the writer lanes (lane = sample, one warp per input) set before and clear after the unit's MMAs.
//   D  (TMEM, fp32)     += A * B_hi + A * B_lo  (= G_hi*P_hi + G_lo*P_hi + G_hi*P_lo + G_lo*P_lo), 128 x 256 per block,
//                         resident for the whole slice; rows o and 64 + o are summed, unscaled and flushed once
//                         (red.global.add, coalesced through shared memory).
// M = 128, N = 256 with both operands in shared memory is the shape the tensor pipe runs at its rate (128 cycles; K = 16
// fp16 values per instruction, 8 in tf32); an MMA with N = 64 costs ~85 cycles whatever its size and a TMEM A operand ~50
// more (the first versions of this kernel -- Phi^T as the TMEM operand with N = outputs: slower than the scatter;
// (dL/dy)^T in TMEM with N = 192: 150 cycles per MMA; tf32 operands: 0.585 ms against 0.50 ms now).  8 MMAs per unit of 64
// samples and block = 4 tensor cycles per (sample, input).  The zero products cost no accuracy (x + 0 is exact in the
// fp32 accumulator).
#include <stdlib.h>
#include <string.h>
#include "hoir_common.cuh"
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int Z_UNIT = 64, Z_WORKERS = 256, Z_THREADS = Z_WORKERS + 32, Z_MMAW = Z_WORKERS / 32;
constexpr int Z_NBLK = 2, Z_NB = 256;                    // blocks per CTA, B rows (= MMA N) per block: 4 inputs x 64 node slots or 2 x 128
constexpr int Z_D_COL = 0;                               // D: 2 blocks x 256 columns = all of TMEM
constexpr uint32_t Z_B_BYTES = Z_NB * 128;               // one Phi^T tile (hi or lo): 256 rows x 64 samples (fp16)
constexpr uint32_t Z_A_BYTES = 128 * 128;                // one (dL/dy)^T tile: 128 rows (hi | lo halves) x 64 samples (fp16)
enum { ZB_AFULL0 = 0, ZB_AFULL1, ZB_AEMPTY0, ZB_AEMPTY1, ZB_BFULL0, ZB_BFULL1, ZB_BEMPTY0, ZB_BEMPTY1, ZB_DONE, ZB_COUNT };
constexpr uint32_t ZS_B = 0, ZS_G = ZS_B + Z_NBLK * 2 * Z_B_BYTES, ZS_BAR = ZS_G + 2 * Z_A_BYTES, ZS_TOTAL = ZS_BAR + 256;
constexpr int Z_FC = 128, Z_FP = Z_FC + 1;               // flush: columns per pass, pitch of the transposed accumulator rows (floats)
static_assert(128 * Z_FP * 4 <= (int)ZS_G, "the flush tile fits the Phi^T tiles");

struct Gw0Params {
  long long B, per_slice;
  const float* x;
  const float* gy;
  float* gw;
  int ldo, in_f, out_f, nodes, stride, segments;
  float segf, inv_segf, rescale;
  int pow2;
  const unsigned* gmax_bits;      // max |dL/dy| of the launch (float bits), written by gy_maxabs_kernel
  LagrangeTable lag;
};

// fp16 operands: dL/dy is scaled by a power of two S so that max |S * dL/dy| lies in [2^13, 2^14) -- gradients of 1e-7 would
// otherwise be fp16 subnormals; the accumulators are divided by S in the flush.  A value v is fed as hi = fp16(v) and
// lo = fp16(v - hi): 22 significant bits, what the tf32 x 3 scheme of the other kernels carries.
__global__ void gy_maxabs_kernel(const float* __restrict__ gy, long long B, int out_f, int ldo, unsigned* __restrict__ gmax_bits) {
  float m = 0.f;
  const long long total = B * out_f;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long b = k / out_f;
    m = fmaxf(m, fabsf(__ldg(gy + b * ldo + (k - b * out_f))));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(gmax_bits, __float_as_uint(m));    // (non-negative floats order like their bits)
}
__device__ __forceinline__ float gy_scale(const unsigned* gmax_bits) {
  const float m = __uint_as_float(*gmax_bits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                    // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 14 - e);       // S * m in [2^13, 2^14)
}
__device__ __forceinline__ void split16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// byte offset of element (row r, column c < 64) of a one-atom-wide K-major fp16 operand (rows of 128 B, 4-row atoms)
__device__ __forceinline__ uint32_t h_off(uint32_t r, uint32_t c) {
  return (r >> 2) * 512u + (r & 3u) * 128u + ((c << 1) ^ ((r & 3u) << 5));
}
// tcgen05.mma kind::f16 (fp16 operands, fp32 accumulator), both operands in shared memory; K = 16 per instruction
__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {     // D fp32, A / B fp16, both K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, int SPS>                                // SPS: log2 of the node slots per input (6, 7 or 8)
__global__ void __launch_bounds__(Z_THREADS, 1) gw0_tc_kernel(const Gw0Params p) {
  constexpr int Z_IB = Z_NB >> SPS;                        // inputs per block
  extern __shared__ __align__(1024) uint8_t zsm[];
  uint8_t* sB = zsm + ZS_B;                                // [block][hi | lo] Phi^T tiles
  uint8_t* sA = zsm + ZS_G;                                // [2] (dL/dy)^T tiles: rows 0..63 hi halves, 64..127 lo halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(zsm + ZS_BAR);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + ZB_COUNT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int i0 = blockIdx.x * Z_NBLK * Z_IB;
  const long long b_lo = (long long)blockIdx.y * p.per_slice;
  const long long b_hi = b_lo + p.per_slice < p.B ? b_lo + p.per_slice : p.B;
  const int units = b_hi > b_lo ? (int)((b_hi - b_lo + Z_UNIT - 1) / Z_UNIT) : 0;

  if (warp == Z_MMAW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[ZB_AFULL0], 128); mbar_init(&bars[ZB_AFULL1], 128);
    mbar_init(&bars[ZB_AEMPTY0], 1); mbar_init(&bars[ZB_AEMPTY1], 1);
    mbar_init(&bars[ZB_BFULL0], Z_IB * 32); mbar_init(&bars[ZB_BFULL1], Z_IB * 32);
    mbar_init(&bars[ZB_BEMPTY0], 1); mbar_init(&bars[ZB_BEMPTY1], 1);
    mbar_init(&bars[ZB_DONE], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int k = tid; k < (int)(ZS_G / 16); k += Z_THREADS) reinterpret_cast<float4*>(zsm)[k] = make_float4(0.f, 0.f, 0.f, 0.f);   // the Phi^T tiles start empty
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;

  if (warp == Z_MMAW) {
    // =========================== MMA warp ===================================================================
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc_f16(128, Z_NB);       // A: (dL/dy)^T K-major, B: Phi^T K-major, N = 256, fp16
    for (int u = 0; u < units; ++u) {
      const uint32_t ab = u & 1;
      mbar_wait(&bars[ZB_AFULL0 + ab], (u >> 1) & 1);
      for (int blk = 0; blk < Z_NBLK; ++blk) {
        mbar_wait(&bars[ZB_BFULL0 + blk], u & 1);
        fence_after();
        if (leader) {
          const uint32_t bbase = smem_u32(sB + blk * 2 * Z_B_BYTES);
          const uint64_t dB[2] = {make_desc(bbase, 512), make_desc(bbase + Z_B_BYTES, 512)};
          const uint64_t dA = make_desc(smem_u32(sA + ab * Z_A_BYTES), 512);
#pragma unroll
          for (int j = 0; j < Z_UNIT / 16; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              mma_ss_f16(tmem + Z_D_COL + blk * Z_NB, dA + (uint64_t)((j * 32) >> 4), dB[h] + (uint64_t)((j * 32) >> 4), idesc,
                     ((uint32_t)u | j | h) != 0);
          umma_commit(&bars[ZB_BEMPTY0 + blk]);
          if (blk == Z_NBLK - 1) umma_commit(&bars[ZB_AEMPTY0 + ab]);
          if (blk == Z_NBLK - 1 && u == units - 1) umma_commit(&bars[ZB_DONE]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // =========================== dL/dy warps: rows of the unit -> (dL/dy)^T tile (hi rows | lo rows) ===========
    const bool gvec = ((p.out_f | p.ldo) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gy) & 15) == 0;
    constexpr int GV = Z_UNIT * 16 / 128;                  // float4 per thread: 64 samples x 64 outputs / 4
    const float S = gy_scale(p.gmax_bits);
    float4 gpre[GV];
    auto prefetch = [&](int u) {                           // the unit's rows, requested one unit ahead
      const long long b0 = b_lo + (long long)u * Z_UNIT;
#pragma unroll
      for (int r = 0; r < GV; ++r) {
        const int it = tid + r * 128, sl = it % Z_UNIT, o4 = it / Z_UNIT;
        const long long b = b0 + sl;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < units && b < b_hi && 4 * o4 < p.out_f) {
          if (gvec) {
            v = __ldg(reinterpret_cast<const float4*>(p.gy + b * p.ldo) + o4);
          } else {
            const float* g = p.gy + b * p.ldo + 4 * o4;
            v.x = __ldg(g);
            if (4 * o4 + 1 < p.out_f) v.y = __ldg(g + 1);
            if (4 * o4 + 2 < p.out_f) v.z = __ldg(g + 2);
            if (4 * o4 + 3 < p.out_f) v.w = __ldg(g + 3);
          }
        }
        gpre[r] = v;
      }
    };
    prefetch(0);
    for (int u = 0; u < units; ++u) {
      const uint32_t ab = u & 1;
      uint8_t* at = sA + ab * Z_A_BYTES;
      mbar_wait(&bars[ZB_AEMPTY0 + ab], ((u >> 1) & 1) ^ 1);      // the MMAs of unit u - 2 are done with this tile
      // element (output o, sample sl): lanes = samples, conflict-free
#pragma unroll
      for (int r = 0; r < GV; ++r) {
        const int it = tid + r * 128, sl = it % Z_UNIT, o4 = it / Z_UNIT;
        const float v4[4] = {gpre[r].x, gpre[r].y, gpre[r].z, gpre[r].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __half hi, lo;
          split16(v4[c] * S, hi, lo);
          *reinterpret_cast<__half*>(at + h_off((uint32_t)(4 * o4 + c), (uint32_t)sl)) = hi;
          *reinterpret_cast<__half*>(at + h_off((uint32_t)(64 + 4 * o4 + c), (uint32_t)sl)) = lo;
        }
      }
      fence_async_smem();
      mbar_arrive(&bars[ZB_AFULL0 + ab]);
      prefetch(u + 1);
    }
  } else {
    // =========================== Phi^T warps: lane = sample, tasks = the 6 inputs of the CTA ====================
    // warp 4 + w takes the tasks w and w + 4 (task t: block t / 3, input t % 3 of the block)
    const int wq = warp - 4;
    int prev[2][2] = {{-1, -1}, {-1, -1}};                  // [task][sample half]: first row of the cells set for the last unit
    float xpre[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    auto prefetch = [&](int u) {
#pragma unroll
      for (int r = 0; r < (Z_NBLK * Z_IB + 3) / 4; ++r) {
        const int t = wq + 4 * r, i = i0 + t;
        if (t >= Z_NBLK * Z_IB) continue;
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const long long b = b_lo + (long long)u * Z_UNIT + hs * 32 + lane;
          xpre[r][hs] = (u < units && i < p.in_f && b < b_hi) ? __ldg(p.x + b * p.in_f + i) : 0.f;
        }
      }
    };
    prefetch(0);
    for (int u = 0; u < units; ++u) {
      float xv[2][2] = {{xpre[0][0], xpre[0][1]}, {xpre[1][0], xpre[1][1]}};
      prefetch(u + 1);
#pragma unroll
      for (int r = 0; r < (Z_NBLK * Z_IB + 3) / 4; ++r) {
        const int t = wq + 4 * r;
        if (t >= Z_NBLK * Z_IB) continue;
        const int blk = t / Z_IB, m = t - blk * Z_IB, i = i0 + t;
        uint8_t* th = sB + blk * 2 * Z_B_BYTES;
        uint8_t* tl = th + Z_B_BYTES;
        int row[2];
        float l[2][N];
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {                       // the lane's two samples of the unit: lane, lane + 32
          const long long b = b_lo + (long long)u * Z_UNIT + hs * 32 + lane;
          const int id = hoir::segment_index(xv[r][hs], 1.f, 2.f, p.segf, p.segments);
          const float xi = p.pow2 ? hoir::local_coordinate<true>(xv[r][hs], id, 1.f, 2.f, p.segf, p.inv_segf)
                                  : hoir::local_coordinate<false>(xv[r][hs], id, 1.f, 2.f, p.segf, p.inv_segf);
          hoir::lagrange<N>(p.lag.X, p.lag.D, xi, l[hs]);
          const float sc = (i < p.in_f && b < b_hi) ? p.rescale : 0.f;      // rows past the slice / inputs past the layer: nothing
#pragma unroll
          for (int j = 0; j < N; ++j) l[hs][j] *= sc;
          row[hs] = (m << SPS) + id * p.stride;
          HOIR_DBG_ASSERT(row[hs] >= 0 && row[hs] + N <= Z_NB);
        }
        mbar_wait(&bars[ZB_BEMPTY0 + blk], (u & 1) ^ 1);                    // the MMAs of unit u - 1 are done with the tiles
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {
          const uint32_t col = (uint32_t)(hs * 32 + lane);
          if (prev[r][hs] >= 0) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const uint32_t off = h_off((uint32_t)(prev[r][hs] + j), col);
              *reinterpret_cast<__half*>(th + off) = __float2half_rn(0.f);
              *reinterpret_cast<__half*>(tl + off) = __float2half_rn(0.f);
            }
          }
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const uint32_t off = h_off((uint32_t)(row[hs] + j), col);
            __half hi, lo;
            split16(l[hs][j], hi, lo);
            *reinterpret_cast<__half*>(th + off) = hi;
            *reinterpret_cast<__half*>(tl + off) = lo;
          }
          prev[r][hs] = row[hs];
        }
        fence_async_smem();
        mbar_arrive(&bars[ZB_BFULL0 + blk]);
      }
    }
  }
  // =========================== flush =========================================================================
  // accumulator lane = (tf32 half, output), column = (input, node slot): transposed through shared memory (the Phi^T tiles
  // are free now) so that rows o and 64 + o are summed and consecutive lanes add to consecutive nodes of one (o, i)
  if (units > 0 && warp != Z_MMAW) {
    mbar_wait(&bars[ZB_DONE], 0);
    fence_after();
  }
  __syncthreads();
  float* sF = reinterpret_cast<float*>(zsm);
  const float inv_S = 1.f / gy_scale(p.gmax_bits);
  for (int hb = 0; hb < Z_NBLK * Z_NB / Z_FC && units > 0; ++hb) {     // passes of 128 accumulator columns
    if (warp < 4) {
      const uint32_t tm_lane = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int g = 0; g < Z_FC; g += 16) {
        float d[16];
        tmem_ld16(tm_lane + Z_D_COL + hb * Z_FC + g, d);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) sF[tid * Z_FP + g + c] = d[c];
      }
    }
    __syncthreads();
    // items (output o, column n of the pass), n fastest
    for (int it = tid; it < 64 * Z_FC; it += Z_THREADS) {
      const int o = it / Z_FC, n = it - o * Z_FC, col = hb * Z_FC + n;
      const int t = col >> SPS, q = col & ((1 << SPS) - 1), i = i0 + t;
      if (o < p.out_f && i < p.in_f && q < p.nodes) {
        const float v = (sF[o * Z_FP + n] + sF[(64 + o) * Z_FP + n]) * inv_S;
        if (v != 0.f) hoir::red_add(p.gw + ((size_t)o * p.in_f + i) * p.nodes + q, v);
      }
    }
    __syncthreads();
  }
  fence_before();
  __syncthreads();
  if (warp == Z_MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

// many-segment piecewise layer whose node slots fit 64, 128 or 256 rows of a Phi^T block and whose outputs fit 64 accumulator lanes
bool hoir_gw0_tc_eligible(const hoir_layer_desc* d, long long B) {
  static const bool off = getenv("HOIR_NO_GW0_TC") != nullptr;
  static const long long min_b = getenv("HOIR_GW0_MIN_B") ? atoll(getenv("HOIR_GW0_MIN_B")) : 4096;
  if (off || d == nullptr || B < min_b) return false;
  if (d->kind != HOIR_CONTINUOUS && d->kind != HOIR_DISCONTINUOUS) return false;
  if (d->length != 2.f || d->periodicity > 0.f || d->segments <= 8 || d->n < 2 || d->n > 3) return false;
  if (d->out_features > 64 || d->in_features < 12) return false;
  return hoir_layer_nodes(d) <= 256;
}

int hoir_gw0_tc(const hoir_layer_desc* d, long long B, const float* x, const float* gy, int ldo, float* gw, cudaStream_t stream) {
  Gw0Params p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.x = x;
  p.gy = gy;
  p.gw = gw;
  p.ldo = ldo;
  p.in_f = d->in_features;
  p.out_f = d->out_features;
  p.nodes = hoir_layer_nodes(d);
  p.stride = hoir::piecewise_stride(d->kind, d->n);
  p.segments = d->segments;
  p.segf = (float)d->segments;
  p.inv_segf = 1.f / (float)d->segments;
  p.rescale = d->rescale == 0.f ? 1.f : d->rescale;
  p.pow2 = hoir::is_pow2(d->segments);
  hoir_make_lagrange_table(d->n, 1.f, &p.lag);
  const int sps = p.nodes <= 64 ? 6 : (p.nodes <= 128 ? 7 : 8);
  const int per_cta = Z_NBLK * (Z_NB >> sps);
  const int groups = (p.in_f + per_cta - 1) / per_cta;
  // a slice = what one CTA accumulates in TMEM before its single flush.  The tensor core TRUNCATES its fp32 accumulator
  // (measured in fused_fourier.cu), a bias that grows with the number of non-zero terms per element: at most 2048 samples
  // per slice (~80 non-zero terms per node at 50 segments) keeps dL/dw within 6e-6 of the fp64 oracle at any batch; the
  // slices then add up in fp32 atomics (round to nearest).
  long long slices = hoir_num_sms() / groups;
  static const long long cap = getenv("HOIR_GW0_SLICE") ? atoll(getenv("HOIR_GW0_SLICE")) : 2048;   // (experiments)
  const long long min_slices = (B + cap - 1) / cap, max_slices = (B + 1023) / 1024;
  if (slices < min_slices) slices = min_slices;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  p.per_slice = ((B + slices - 1) / slices + Z_UNIT - 1) / Z_UNIT * Z_UNIT;
  slices = (B + p.per_slice - 1) / p.per_slice;
  void* ws = nullptr;
  if (int e = hoir_get_workspace(5, 256, &ws)) return e;
  p.gmax_bits = reinterpret_cast<const unsigned*>(ws);
  HOIR_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(unsigned), stream));
  {
    const long long total = B * p.out_f;
    long long grid = (total + 256 * 8 - 1) / (256 * 8);
    if (grid > 4LL * hoir_num_sms()) grid = 4LL * hoir_num_sms();
    gy_maxabs_kernel<<<(unsigned)grid, 256, 0, stream>>>(gy, B, p.out_f, ldo, reinterpret_cast<unsigned*>(ws));
  }
  auto kern = sps == 6 ? (d->n == 2 ? gw0_tc_kernel<2, 6> : gw0_tc_kernel<3, 6>)
            : sps == 7 ? (d->n == 2 ? gw0_tc_kernel<2, 7> : gw0_tc_kernel<3, 7>)
                       : (d->n == 2 ? gw0_tc_kernel<2, 8> : gw0_tc_kernel<3, 8>);
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZS_TOTAL));
  kern<<<dim3((unsigned)groups, (unsigned)slices), Z_THREADS, ZS_TOTAL, stream>>>(p);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
