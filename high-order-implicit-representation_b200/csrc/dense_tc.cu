// dense_tc.cu -- single high-order layers whose densified basis expansion is a real dense contraction,
// on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   y[b, o] = sum_k Phi[b, k] * W[k, o],      k = (input i, basis slot q),  Phi[b, (i, q)] = phi_q(x[b, i])
//
// Fourier layers (phi = 1/2, sin / cos pairs: every slot non-zero, K = in * n -- 1240 for the 40-wide
// n = 31 hidden layer of BASELINE config C3) and piecewise-Lagrange layers with few segments (hidden /
// output layers of every reference network: 3 .. 6 nodes per input, n of them non-zero) qualify; layers
// with many segments (the 100- / 50-segment input layers) are gathers and stay on the CUDA cores
// (layer_kernels.cu, slab_kernels.cu).  Algorithm: oracle/hol_oracle.py; call sites
// /root/reference/high_order_implicit_representation/networks.py:41-55, 148-161.
//
// K is cut into chunks of 32 columns (one 128-byte swizzle atom): IPC = 32 / ND inputs of ND slots each,
// zero padded.  Every product runs as 3 tf32 passes  A*B + A_lo*B + A*B_lo  (x_lo = x - trunc_tf32(x)),
// fp32-level accuracy (1.3e-6 relative, csrc/probe/umma_probe.cu).
//
//   forward   D[256 x OUTP]  = Phi * W        A: TMEM, written by the sample threads chunk by chunk
//                                             B: smem ring of W chunks (K-major), streamed from L2 by
//                                                cp.async.bulk out of a pre-swizzled hi/lo image
//   dL/dx     T[256 x 32]    = G * W_chunk^T  A: TMEM (G staged once per tile), B: the same W ring (MN-major)
//                                             epilogue: dL/dx[b, i] = sum_q phi'_q(x) * T[b, (i, q)]
//   dL/dw     gW[128 x OUTP] += Phi^T * G     CTA = (block of 128 K-columns, batch slice); A: smem units of
//                                             32 samples (MN-major), B: smem G^T (K-major); the accumulator
//                                             lives in TMEM for the whole kernel and is flushed once.
#include <stdlib.h>
#include <string.h>
#include "hoir_common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

struct DenseParams {
  long long B;
  int in_f, out_f, n, nodes, nch, outp, outp8;
  int ldo, gx_acc;          // row stride of y / gy (an output block of a wider layer: > out_f); gx += instead of =
  float rescale, theta_c, length;
  int odd_is_sin;
  const float* x;
  const float* gy;
  float* y;
  float* gx;
  float* gw;
  const float* wimg;        // [nch][hi | lo][outp x 32] pre-swizzled chunk images
  LagrangeTable lag;
};

// ---- basis policies: ND slots per input, IPC inputs per 32-column chunk ---------------------------------
template <int NSLOT>
struct FourierP {
  static constexpr int ND = NSLOT, IPC = 32 / NSLOT;
  // phi_0 = 1/2, then slot j holds sin or cos of harmonic (j + 1) / 2 of theta = c_1 * x / length
  // (SURVEY.md A.5); harmonics by the angle-addition recurrence from one sincosf.
  template <bool DER>
  __device__ static __forceinline__ void eval(const DenseParams& p, float x, float* __restrict__ val,
                                              float* __restrict__ der) {
    const float theta = __fdiv_rn(__fmul_rn(p.theta_c, x), p.length);
    const float dth = p.theta_c / p.length;
    float s1, c1;
    sincosf(theta, &s1, &c1);
    float sh = s1, ch = c1;
    val[0] = 0 < p.n ? 0.5f : 0.f;
    if (DER) der[0] = 0.f;
#pragma unroll
    for (int j = 1; j < ND; ++j) {
      const bool is_sin = ((j & 1) == 1) == (p.odd_is_sin != 0);
      const bool on = j < p.n;
      val[j] = on ? (is_sin ? sh : ch) : 0.f;
      if (DER) {
        const float hf = (float)((j + 1) / 2) * dth;
        der[j] = on ? (is_sin ? ch * hf : -sh * hf) : 0.f;
      }
      if ((j & 1) == 0) {   // both slots of harmonic h written: advance to h + 1
        const float sn = fmaf(sh, c1, ch * s1), cn = fmaf(ch, c1, -sh * s1);
        sh = sn;
        ch = cn;
      }
    }
  }
};

template <int N, int SEG, bool CONT>
struct PieceP {
  static constexpr int STRIDE = CONT ? N - 1 : N;
  static constexpr int ND = STRIDE * SEG + (CONT ? 1 : 0), IPC = 32 / ND;
  static constexpr bool POW2 = (SEG & (SEG - 1)) == 0;
  template <bool DER>
  __device__ static __forceinline__ void eval(const DenseParams& p, float x, float* __restrict__ val,
                                              float* __restrict__ der) {
    const int id = hoir::segment_index(x, 1.f, 2.f, (float)SEG, SEG);
    const float xi = hoir::local_coordinate<POW2>(x, id, 1.f, 2.f, (float)SEG, 1.f / (float)SEG);
    float l[N], d[N];
    if (DER) hoir::lagrange_deriv<N>(p.lag.X, p.lag.D, xi, l, d);
    else hoir::lagrange<N>(p.lag.X, p.lag.D, xi, l);
    const float slope = hoir::local_slope<POW2>(id, 2.f, (float)SEG);
    const int base = id * STRIDE;
#pragma unroll
    for (int q = 0; q < ND; ++q) {
      float v = 0.f, dv = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        v = (q == base + j) ? l[j] : v;
        if (DER) dv = (q == base + j) ? d[j] : dv;
      }
      val[q] = v;
      if (DER) der[q] = dv * slope;
    }
  }
};

// byte offset of element (row r, column c < 32) of a one-atom-wide operand (rows of 128 B, RS = 512)
__device__ __forceinline__ uint32_t chunk_off(uint32_t r, uint32_t c) {
  return (r >> 2) * 512u + (r & 3u) * 128u + ((c << 2) ^ ((r & 3u) << 5));
}

// ---- weight image: w[out][in][nodes] -> [chunk][hi | lo][OUTP rows x 32 columns], swizzled ---------------
__global__ void dense_pack_kernel(const float* __restrict__ w, float* __restrict__ img, int in_f, int out_f,
                                  int nodes, int nd, int ipc, int nch, int outp) {
  const long long total = (long long)nch * outp * 32;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(k & 31);
    const long long r = k >> 5;
    const int o = (int)(r % outp), c = (int)(r / outp);
    const int m = col / nd, q = col - m * nd, i = c * ipc + m;
    float v = 0.f;
    if (m < ipc && i < in_f && q < nodes && o < out_f) v = w[((size_t)o * in_f + i) * nodes + q];
    float* chunk = img + (size_t)c * 2 * outp * 32;
    const uint32_t off = chunk_off((uint32_t)o, (uint32_t)col) >> 2;
    chunk[off] = v;
    chunk[(size_t)outp * 32 + off] = tf32_lo(v);
  }
}

// ---- forward / dL/dx -------------------------------------------------------------------------------------
constexpr int DT_TILE = 256, DT_SAMPLE_THREADS = 256, DT_THREADS = 320, DT_STAGES = 4;
constexpr int DT_A_COL = 0, DT_D_COL = 256;
enum { DB_AFULL0 = 0, DB_AFULL1, DB_AEMPTY0, DB_AEMPTY1, DB_DFULL, DB_DEMPTY, DB_GFULL, DB_TFULL0, DB_TFULL1,
       DB_TEMPTY0, DB_TEMPTY1, DB_WFULL0, DB_WEMPTY0 = DB_WFULL0 + DT_STAGES, DB_COUNT = DB_WEMPTY0 + DT_STAGES };

template <class P, bool GXMODE>
__global__ void __launch_bounds__(DT_THREADS, 1) dense_tc_kernel(const DenseParams p) {
  constexpr int ND = P::ND, IPC = P::IPC;
  extern __shared__ __align__(1024) uint8_t dsmem[];
  const uint32_t stage_bytes = (uint32_t)p.outp * 256u;          // hi + lo images of one chunk
  uint8_t* sW = dsmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsmem + DT_STAGES * 16384);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + DB_COUNT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[DB_AFULL0], DT_SAMPLE_THREADS); mbar_init(&bars[DB_AFULL1], DT_SAMPLE_THREADS);
    mbar_init(&bars[DB_AEMPTY0], 1); mbar_init(&bars[DB_AEMPTY1], 1);
    mbar_init(&bars[DB_DFULL], 1); mbar_init(&bars[DB_DEMPTY], DT_SAMPLE_THREADS);
    mbar_init(&bars[DB_GFULL], DT_SAMPLE_THREADS);
    mbar_init(&bars[DB_TFULL0], 1); mbar_init(&bars[DB_TFULL1], 1);
    mbar_init(&bars[DB_TEMPTY0], DT_SAMPLE_THREADS); mbar_init(&bars[DB_TEMPTY1], DT_SAMPLE_THREADS);
    for (int s = 0; s < DT_STAGES; ++s) { mbar_init(&bars[DB_WFULL0 + s], 1); mbar_init(&bars[DB_WEMPTY0 + s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;
  const long long tiles = (p.B + DT_TILE - 1) / DT_TILE;
  const int nch = p.nch;

  if (warp == 9) {
    // =========================== W loader: one elected lane streams chunk images into the ring =========
    if (elect_one()) {
      uint32_t qw = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int c = 0; c < nch; ++c, ++qw) {
          const uint32_t st = qw % DT_STAGES, ph = (qw / DT_STAGES) & 1;
          mbar_wait(&bars[DB_WEMPTY0 + st], ph ^ 1);
          mbar_expect_tx(&bars[DB_WFULL0 + st], stage_bytes);
          bulk_g2s(sW + st * 16384, p.wimg + (size_t)c * 2 * p.outp * 32, stage_bytes, &bars[DB_WFULL0 + st]);
        }
      }
    }
  } else if (warp == 8) {
    // =========================== MMA warp ==============================================================
    const uint32_t leader = elect_one();
    const uint32_t id_fwd = make_idesc(128, p.outp, 0, 0);      // A: TMEM, B: W chunk K-major (rows = outputs)
    constexpr uint32_t ID_GX = make_idesc(128, 32, 0, 1);       // A: TMEM, B: W chunk MN-major (N = chunk columns)
    uint32_t qw = 0, qa = 0, n_tile = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      if (!GXMODE) {
        if (n_tile > 0) mbar_wait(&bars[DB_DEMPTY], (n_tile - 1) & 1);
      } else {
        mbar_wait(&bars[DB_GFULL], n_tile & 1);
      }
      for (int c = 0; c < nch; ++c, ++qw, ++qa) {
        const uint32_t st = qw % DT_STAGES, wph = (qw / DT_STAGES) & 1;
        const uint32_t buf = qa & 1, ph = (qa >> 1) & 1;
        if (!GXMODE) mbar_wait(&bars[DB_AFULL0 + buf], ph);
        else mbar_wait(&bars[DB_TEMPTY0 + buf], ph ^ 1);
        mbar_wait(&bars[DB_WFULL0 + st], wph);
        fence_after();
        if (leader) {
          const uint32_t wbase = smem_u32(sW + st * 16384);
          const uint64_t dW[2] = {make_desc(wbase, 512), make_desc(wbase + (uint32_t)p.outp * 128u, 512)};
          if (!GXMODE) {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
              for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  mma_ts(tmem + DT_D_COL + blk * 64, tmem + DT_A_COL + blk * 128 + buf * 64 + (pass == 1 ? 32 : 0) + j * 8,
                         dW[pass == 2] + (uint64_t)((j * 32) >> 4), id_fwd, (c | pass | j) != 0);
            umma_commit(&bars[DB_AEMPTY0 + buf]);
            umma_commit(&bars[DB_WEMPTY0 + st]);
            if (c == nch - 1) umma_commit(&bars[DB_DFULL]);
          } else {
            const int ksteps = p.outp8 >> 3;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
              for (int pass = 0; pass < 3; ++pass)
                for (int j = 0; j < ksteps; ++j)
                  mma_ts(tmem + DT_D_COL + (blk * 2 + buf) * 32, tmem + DT_A_COL + blk * 128 + (pass == 1 ? p.outp8 : 0) + j * 8,
                         dW[pass == 2] + (uint64_t)((j * 2 * 512) >> 4), ID_GX, (pass | j) != 0);
            umma_commit(&bars[DB_TFULL0 + buf]);
            umma_commit(&bars[DB_WEMPTY0 + st]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== sample warps: one thread per sample ===================================
    const int s = tid, blk = warp >> 2;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t qa = 0, n_tile = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      const long long b = tile * DT_TILE + s;
      const bool live = b < p.B;
      const float* xr = p.x + (live ? b : 0) * p.in_f;
      if (!GXMODE) {
        for (int c = 0; c < nch; ++c, ++qa) {
          const uint32_t buf = qa & 1, ph = (qa >> 1) & 1;
          float v[32], lo[32];
#pragma unroll
          for (int k = IPC * ND; k < 32; ++k) v[k] = 0.f;
#pragma unroll
          for (int m = 0; m < IPC; ++m) {
            const int i = c * IPC + m;
            const float xv = (live && i < p.in_f) ? __ldg(xr + i) : 0.f;
            P::template eval<false>(p, xv, v + m * ND, nullptr);
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) lo[k] = tf32_lo(v[k]);
          mbar_wait(&bars[DB_AEMPTY0 + buf], ph ^ 1);
          fence_after();
          const uint32_t col = tm_lane + DT_A_COL + blk * 128 + buf * 64;
          tmem_st32(col, v);
          tmem_st32(col + 32, lo);
          tmem_wait_st();
          fence_before();
          mbar_arrive(&bars[DB_AFULL0 + buf]);
        }
        mbar_wait(&bars[DB_DFULL], n_tile & 1);
        fence_after();
        for (int g = 0; g < p.outp; g += 16) {
          float d[16];
          tmem_ld16(tm_lane + DT_D_COL + blk * 64 + g, d);
          tmem_wait_ld();
          if (live) {
#pragma unroll
            for (int o = 0; o < 16; ++o)
              if (g + o < p.out_f) p.y[b * p.ldo + g + o] = d[o] * p.rescale;
          }
        }
        fence_before();
        mbar_arrive(&bars[DB_DEMPTY]);
      } else {
        // stage G (scaled dL/dy row, hi | lo); the previous tile's products are complete (its last T chunk was consumed)
        const float* gr = p.gy + (live ? b : 0) * p.ldo;
        for (int g = 0; g < p.outp8; g += 8) {
          float v[8], lo[8];
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            v[o] = (live && g + o < p.out_f) ? __ldg(gr + g + o) * p.rescale : 0.f;
            lo[o] = tf32_lo(v[o]);
          }
          tmem_st8(tm_lane + DT_A_COL + blk * 128 + g, v);
          tmem_st8(tm_lane + DT_A_COL + blk * 128 + p.outp8 + g, lo);
        }
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[DB_GFULL]);
        for (int c = 0; c < nch; ++c, ++qa) {
          const uint32_t buf = qa & 1, ph = (qa >> 1) & 1;
          // basis derivatives of this chunk's inputs while the tensor pipe works
          float der[IPC * ND];
#pragma unroll
          for (int m = 0; m < IPC; ++m) {
            const int i = c * IPC + m;
            const float xv = (live && i < p.in_f) ? __ldg(xr + i) : 0.f;
            float val[ND];
            P::template eval<true>(p, xv, val, der + m * ND);
          }
          mbar_wait(&bars[DB_TFULL0 + buf], ph);
          fence_after();
          float t[32];
          tmem_ld16(tm_lane + DT_D_COL + (blk * 2 + buf) * 32, t);
          tmem_ld16(tm_lane + DT_D_COL + (blk * 2 + buf) * 32 + 16, t + 16);
          tmem_wait_ld();
          fence_before();
          mbar_arrive(&bars[DB_TEMPTY0 + buf]);
#pragma unroll
          for (int m = 0; m < IPC; ++m) {
            const int i = c * IPC + m;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < ND; ++q) acc = fmaf(der[m * ND + q], t[m * ND + q], acc);
            if (live && i < p.in_f) {
              float* gxp = p.gx + b * p.in_f + i;
              *gxp = p.gx_acc ? *gxp + acc : acc;
            }
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- dL/dw -----------------------------------------------------------------------------------------------
constexpr int GW_UNIT = 32, GW_WORKERS = 256, GW_THREADS = GW_WORKERS + 32, GW_MMAW = GW_WORKERS / 32;   // 8 worker warps + MMA warp
constexpr uint32_t GW_U_RS = 128 / 32 * 512, GW_U_BYTES = GW_UNIT / 4 * GW_U_RS;   // 2048, 16384
enum { GB_UFULL0 = 0, GB_UFULL1, GB_UEMPTY0, GB_UEMPTY1, GB_DONE, GB_COUNT };

template <class P>
__global__ void __launch_bounds__(GW_THREADS, 2) dense_gw_kernel(const DenseParams p, long long per_slice) {
  constexpr int ND = P::ND, IPC = P::IPC;
  extern __shared__ __align__(1024) uint8_t dsmem[];
  uint8_t* sU = dsmem;                                   // [2][hi | lo] units: rows = samples, 128 columns
  uint8_t* sG = dsmem + 4 * GW_U_BYTES;                  // [2][hi | lo] G^T: rows = outputs, 32 samples
  const uint32_t g_bytes = (uint32_t)p.outp * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 4 * 8192);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + GB_COUNT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int mb = blockIdx.x;                             // block of 4 chunks = 128 K-columns
  const long long b_lo = (long long)blockIdx.y * per_slice;
  const long long b_hi = b_lo + per_slice < p.B ? b_lo + per_slice : p.B;
  const int units = b_hi > b_lo ? (int)((b_hi - b_lo + GW_UNIT - 1) / GW_UNIT) : 0;

  if (warp == GW_MMAW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[GB_UFULL0], GW_WORKERS); mbar_init(&bars[GB_UFULL1], GW_WORKERS);
    mbar_init(&bars[GB_UEMPTY0], 1); mbar_init(&bars[GB_UEMPTY1], 1); mbar_init(&bars[GB_DONE], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;

  if (warp == GW_MMAW) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(128, p.outp, 1, 0);     // A: Phi unit MN-major, B: G^T K-major
    for (int u = 0; u < units; ++u) {
      const uint32_t buf = u & 1, ph = (u >> 1) & 1;
      mbar_wait(&bars[GB_UFULL0 + buf], ph);
      fence_after();
      if (leader) {
        const uint32_t ubase = smem_u32(sU + buf * 2 * GW_U_BYTES), gbase = smem_u32(sG + buf * 2 * 8192);
        const uint64_t dU[2] = {make_desc(ubase, GW_U_RS), make_desc(ubase + GW_U_BYTES, GW_U_RS)};
        const uint64_t dG[2] = {make_desc(gbase, 512), make_desc(gbase + g_bytes, 512)};
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int j = 0; j < GW_UNIT / 8; ++j)
            mma_ss(tmem, dU[pass == 1] + (uint64_t)((j * 2 * GW_U_RS) >> 4), dG[pass == 2] + (uint64_t)((j * 32) >> 4),
                   idesc, ((uint32_t)u | pass | j) != 0);
        umma_commit(&bars[GB_UEMPTY0 + buf]);
        if (u == units - 1) umma_commit(&bars[GB_DONE]);
      }
      __syncwarp();
    }
  } else {
    for (int u = 0; u < units; ++u) {
      const uint32_t buf = u & 1, ph = (u >> 1) & 1;
      const long long b0 = b_lo + (long long)u * GW_UNIT;
      // dL/dy rows of the unit: issued first (16-byte loads when the rows allow it) so that their latency hides
      // behind the basis evaluation below; element (output o, sample sl) -> G^T smem, sl fastest (conflict-free)
      const bool gvec = ((p.out_f | p.ldo) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gy) & 15) == 0;
      constexpr int GV = (GW_UNIT * 16 + GW_WORKERS - 1) / GW_WORKERS;   // float4 per thread: 32 samples x (outp <= 64) / 4
      float4 gpre[GV];
      if (gvec) {
#pragma unroll
        for (int r = 0; r < GV; ++r) {
          const int it = tid + r * GW_WORKERS, sl = it % GW_UNIT, o4 = it / GW_UNIT;
          const long long b = b0 + sl;
          gpre[r] = (b < b_hi && 4 * o4 < p.out_f) ? __ldg(reinterpret_cast<const float4*>(p.gy + b * p.ldo) + o4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(&bars[GB_UEMPTY0 + buf], ph ^ 1);
      uint8_t* ur = sU + buf * 2 * GW_U_BYTES;
      uint8_t* ul = ur + GW_U_BYTES;
      // Phi unit: items (chunk cc of this block, sample sl, input m of the chunk), m fastest
      for (int it = tid; it < 4 * GW_UNIT * IPC; it += GW_WORKERS) {
        const int m = it % IPC, r2 = it / IPC, sl = r2 % GW_UNIT, cc = r2 / GW_UNIT;
        const int c = mb * 4 + cc, i = c * IPC + m;
        const long long b = b0 + sl;
        float val[ND];
        if (b < b_hi && c < p.nch && i < p.in_f) {
          P::template eval<false>(p, __ldg(p.x + b * p.in_f + i), val, nullptr);
        } else {
#pragma unroll
          for (int q = 0; q < ND; ++q) val[q] = 0.f;
        }
        const uint32_t rowoff = (uint32_t)(sl >> 2) * GW_U_RS + (uint32_t)(sl & 3) * 128u + (uint32_t)cc * 512u;
        const uint32_t sw = (uint32_t)(sl & 3) << 5;
        if (ND == 32) {
#pragma unroll
          for (int q4 = 0; q4 < ND / 4; ++q4) {
            const uint32_t off = rowoff + (((uint32_t)q4 << 4) ^ sw);
            *reinterpret_cast<float4*>(ur + off) = make_float4(val[4 * q4], val[4 * q4 + 1], val[4 * q4 + 2], val[4 * q4 + 3]);
            *reinterpret_cast<float4*>(ul + off) = make_float4(tf32_lo(val[4 * q4]), tf32_lo(val[4 * q4 + 1]),
                                                               tf32_lo(val[4 * q4 + 2]), tf32_lo(val[4 * q4 + 3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < ND; ++q) {
            const uint32_t off = rowoff + ((((uint32_t)(m * ND + q)) << 2) ^ sw);
            *reinterpret_cast<float*>(ur + off) = val[q];
            *reinterpret_cast<float*>(ul + off) = tf32_lo(val[q]);
          }
        }
      }
      if (IPC * ND < 32 && u < 2) {   // pad columns of each chunk: zero once per buffer
        for (int it = tid; it < 4 * GW_UNIT * (32 - IPC * ND); it += GW_WORKERS) {
          const int k = it % (32 - IPC * ND), r2 = it / (32 - IPC * ND), sl = r2 % GW_UNIT, cc = r2 / GW_UNIT;
          const uint32_t off = (uint32_t)(sl >> 2) * GW_U_RS + (uint32_t)(sl & 3) * 128u + (uint32_t)cc * 512u +
                               ((((uint32_t)(IPC * ND + k)) << 2) ^ ((uint32_t)(sl & 3) << 5));
          *reinterpret_cast<float*>(ur + off) = 0.f;
          *reinterpret_cast<float*>(ul + off) = 0.f;
        }
      }
      uint8_t* gr = sG + buf * 2 * 8192;
      uint8_t* gl = gr + g_bytes;
      if (gvec) {
#pragma unroll
        for (int r = 0; r < GV; ++r) {
          const int it = tid + r * GW_WORKERS, sl = it % GW_UNIT, o4 = it / GW_UNIT;
          if (4 * o4 < p.outp) {
            const float v4[4] = {gpre[r].x * p.rescale, gpre[r].y * p.rescale, gpre[r].z * p.rescale, gpre[r].w * p.rescale};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t off = chunk_off((uint32_t)(4 * o4 + c), (uint32_t)sl);
              *reinterpret_cast<float*>(gr + off) = v4[c];
              *reinterpret_cast<float*>(gl + off) = tf32_lo(v4[c]);
            }
          }
        }
      } else {
        for (int it = tid; it < p.outp * GW_UNIT; it += GW_WORKERS) {
          const int sl = it % GW_UNIT, o = it / GW_UNIT;
          const long long b = b0 + sl;
          const float v = (b < b_hi && o < p.out_f) ? __ldg(p.gy + b * p.ldo + o) * p.rescale : 0.f;
          const uint32_t off = chunk_off((uint32_t)o, (uint32_t)sl);
          *reinterpret_cast<float*>(gr + off) = v;
          *reinterpret_cast<float*>(gl + off) = tf32_lo(v);
        }
      }
      fence_async_smem();
      mbar_arrive(&bars[GB_UFULL0 + buf]);
    }
    if (units > 0) {
      mbar_wait(&bars[GB_DONE], 0);
      fence_after();
      // accumulator lane rt = K-column of this block: chunk rt / 32, slot rt % 32; warps w and w + 4 share the
      // lane quadrant w & 3 and take alternating groups of 16 output columns
      const int rt = tid & 127;
      const int c = mb * 4 + (rt >> 5), within = rt & 31, m = within / ND, q = within - m * ND, i = c * IPC + m;
      const bool valid = c < p.nch && m < IPC && i < p.in_f && q < p.nodes;
      const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int g = (warp >> 2) * 16; g < p.outp; g += 16 * (GW_WORKERS / 128)) {
        float d[16];
        tmem_ld16(tm_lane + g, d);
        tmem_wait_ld();
        if (valid) {
#pragma unroll
          for (int o = 0; o < 16; ++o)
            if (g + o < p.out_f && d[o] != 0.f) hoir::red_add(p.gw + ((size_t)(g + o) * p.in_f + i) * p.nodes + q, d[o]);
        }
      }
      fence_before();
    }
  }
  __syncthreads();
  if (warp == GW_MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
  (void)lane;
}

// ---- host side -------------------------------------------------------------------------------------------
struct DenseEntry {
  int kind, n, seg;          // piecewise: exact (n, segments); Fourier: n <= slots
  int nd, ipc;
  void (*fwd)(const DenseParams);
  void (*gx)(const DenseParams);
  void (*gw)(const DenseParams, long long);
};
template <class P>
constexpr DenseEntry dense_entry(int kind, int n, int seg) {
  return DenseEntry{kind, n, seg, P::ND, P::IPC, dense_tc_kernel<P, false>, dense_tc_kernel<P, true>, dense_gw_kernel<P>};
}
const DenseEntry kDense[] = {
    dense_entry<FourierP<4>>(HOIR_FOURIER, 4, 0),
    dense_entry<FourierP<8>>(HOIR_FOURIER, 8, 0),
    dense_entry<FourierP<16>>(HOIR_FOURIER, 16, 0),
    dense_entry<FourierP<32>>(HOIR_FOURIER, 32, 0),
    dense_entry<PieceP<2, 2, true>>(HOIR_CONTINUOUS, 2, 2),
    dense_entry<PieceP<3, 2, true>>(HOIR_CONTINUOUS, 3, 2),
    dense_entry<PieceP<2, 2, false>>(HOIR_DISCONTINUOUS, 2, 2),
    dense_entry<PieceP<3, 2, false>>(HOIR_DISCONTINUOUS, 3, 2),
};

int dense_sms() { return hoir_num_sms(); }

// ---- measured tensor-pipe roofline denominator: tcgen05.mma kind::tf32, M = 128, N = 256, K = 8 per instruction, both
// operands resident in shared memory (SWIZZLE_128B_BASE32B, K-major), accumulator in TMEM; one CTA per SM issues
// `iters` rounds of 4 back-to-back MMAs (one 32-column K atom) and commits once.
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters) {
  extern __shared__ __align__(1024) uint8_t psm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int k = tid; k < (128 + 256) * 128 / 4; k += 128) reinterpret_cast<float*>(psm)[k] = 1.0f + (float)(k & 7) * 0.125f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint64_t dA = make_desc(smem_u32(psm), 512), dB = make_desc(smem_u32(psm + 128 * 128), 512);
      constexpr uint32_t ID = make_idesc(128, 256, 0, 0);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_ss(tmem, dA + (uint64_t)((j * 32) >> 4), dB + (uint64_t)((j * 32) >> 4), ID, (it | j) != 0);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait_backoff(&bar, 0, 256);
  fence_after();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

}  // namespace

// elapsed milliseconds and FLOPs of `iters` rounds of 4 tcgen05 tf32 MMAs (128 x 256 x 8) on every SM: the measured
// tensor roofline denominator for the tf32 x 3 kernels (bench.py)
extern "C" int hoir_tf32_peak(int iters, float* ms, double* flops, void* stream) {
  if (iters < 1 || !ms || !flops) return hoir_set_error(HOIR_EINVAL, "bad arguments", "hoir_tf32_peak");
  const int sms = dense_sms();
  const size_t smem = (128 + 256) * 128 + 1024;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  HOIR_CHECK_CUDA(cudaEventCreate(&e0));
  HOIR_CHECK_CUDA(cudaEventCreate(&e1));
  cudaStream_t s = (cudaStream_t)stream;
  tf32_peak_kernel<<<sms, 128, smem, s>>>(iters / 8 + 1);      // warm-up
  HOIR_CHECK_CUDA(cudaEventRecord(e0, s));
  tf32_peak_kernel<<<sms, 128, smem, s>>>(iters);
  HOIR_CHECK_CUDA(cudaEventRecord(e1, s));
  HOIR_CHECK_CUDA(cudaEventSynchronize(e1));
  HOIR_CHECK_CUDA(cudaEventElapsedTime(ms, e0, e1));
  *flops = 2.0 * 128.0 * 256.0 * 8.0 * 4.0 * (double)iters * (double)sms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// nullptr: this layer is not a dense contraction (or too small to be worth the tensor path)
const void* hoir_dense_tc_find(const hoir_layer_desc* d, long long B) {
  static const bool off = getenv("HOIR_NO_DENSE_TC") != nullptr;
  // (from batch 128: at the reference's batch size, 256, one tensor-core tile is 2x faster than the generic tile kernel)
  if (off || d == nullptr || B < 128) return nullptr;
  if (d->length != 2.f || d->periodicity > 0.f || d->out_features > 64 || d->in_features < 1) return nullptr;
  for (const DenseEntry& e : kDense) {
    if (e.kind != d->kind) continue;
    if (d->kind == HOIR_FOURIER) {
      if (d->n <= e.n) return &e;      // entries are ordered by slot count
    } else if (d->n == e.n && d->segments == e.seg) {
      return &e;
    }
  }
  return nullptr;
}

static int dense_fill(const void* entry, const hoir_layer_desc* d, long long B, const float* w, DenseParams* p,
                      cudaStream_t stream) {
  const DenseEntry& e = *static_cast<const DenseEntry*>(entry);
  memset(p, 0, sizeof(*p));
  p->B = B;
  p->in_f = d->in_features;
  p->out_f = d->out_features;
  p->ldo = d->out_features;
  p->n = d->n;
  p->nodes = hoir_layer_nodes(d);
  p->nch = (d->in_features + e.ipc - 1) / e.ipc;
  p->outp = (d->out_features + 15) & ~15;
  p->outp8 = (d->out_features + 7) & ~7;
  p->rescale = d->rescale == 0.f ? 1.f : d->rescale;
  p->length = d->length;
  FourierTable ft;
  hoir_make_fourier_table(2, d->fourier_arg_scale == 0.f ? 2.f : d->fourier_arg_scale, &ft);
  p->theta_c = ft.c[1];
  p->odd_is_sin = d->fourier_odd_is_sin;
  hoir_make_lagrange_table(d->kind == HOIR_FOURIER ? 1 : d->n, 1.f, &p->lag);
  // weight image in the library workspace (slot 1), refreshed per call
  const size_t img_floats = (size_t)p->nch * 2 * p->outp * 32;
  void* ws = nullptr;
  if (int err = hoir_get_workspace(1, img_floats * sizeof(float), &ws)) return err;
  const long long total = (long long)p->nch * p->outp * 32;
  int grid = (int)((total + 255) / 256);
  if (grid > dense_sms() * 8) grid = dense_sms() * 8;
  dense_pack_kernel<<<grid, 256, 0, stream>>>(w, static_cast<float*>(ws), p->in_f, p->out_f, p->nodes, e.nd, e.ipc,
                                              p->nch, p->outp);
  HOIR_CHECK_CUDA(cudaGetLastError());
  p->wimg = static_cast<const float*>(ws);
  return HOIR_OK;
}

// ldo: row stride of y (>= out_features; an output block of a wider layer)
int hoir_dense_tc_forward(const void* entry, const hoir_layer_desc* d, long long B, const float* x, const float* w,
                          float* y, int ldo, cudaStream_t stream) {
  const DenseEntry& e = *static_cast<const DenseEntry*>(entry);
  DenseParams p;
  if (int err = dense_fill(entry, d, B, w, &p, stream)) return err;
  p.x = x;
  p.y = y;
  p.ldo = ldo;
  const size_t smem = DT_STAGES * 16384 + 512;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(e.fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (B + DT_TILE - 1) / DT_TILE;
  const int grid = (int)(tiles < dense_sms() ? tiles : dense_sms());
  e.fwd<<<grid, DT_THREADS, smem, stream>>>(p);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// ldo: row stride of gy; gx_acc: add this block's dL/dx to gx instead of overwriting it
int hoir_dense_tc_backward(const void* entry, const hoir_layer_desc* d, long long B, const float* x, const float* w,
                           const float* gy, int ldo, float* gx, int gx_acc, float* gw, cudaStream_t stream) {
  const DenseEntry& e = *static_cast<const DenseEntry*>(entry);
  DenseParams p;
  if (int err = dense_fill(entry, d, B, w, &p, stream)) return err;
  p.x = x;
  p.gy = gy;
  p.ldo = ldo;
  p.gx = gx;
  p.gx_acc = gx_acc;
  p.gw = gw;
  if (gx != nullptr) {
    const size_t smem = DT_STAGES * 16384 + 512;
    HOIR_CHECK_CUDA(cudaFuncSetAttribute(e.gx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (B + DT_TILE - 1) / DT_TILE;
    const int grid = (int)(tiles < dense_sms() ? tiles : dense_sms());
    e.gx<<<grid, DT_THREADS, smem, stream>>>(p);
    HOIR_CHECK_CUDA(cudaGetLastError());
  }
  if (gw != nullptr) {
    const size_t smem = 4 * GW_U_BYTES + 4 * 8192 + 512;
    HOIR_CHECK_CUDA(cudaFuncSetAttribute(e.gw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nmb = (p.nch + 3) / 4;
    // one wave of two CTAs per SM -- but every CTA flushes its whole [128 x OUTP] accumulator with global adds, so
    // a slice should amortise that (and the TMEM / barrier prologue) over a dozen units
    long long slices = (2LL * dense_sms()) / nmb;
    // measured: a single 128-column block cut into 296 slices of 7 units costs 128 us, 85 slices of 24 units 83 us;
    // with two or more blocks, halving the CTA count to lengthen the slices doubled the time instead
    const int min_units = nmb == 1 ? 24 : 12;
    const long long max_slices = (B + (long long)min_units * GW_UNIT - 1) / ((long long)min_units * GW_UNIT);
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    long long per = (B + slices - 1) / slices;
    per = (per + GW_UNIT - 1) / GW_UNIT * GW_UNIT;
    slices = (B + per - 1) / per;
    e.gw<<<dim3((unsigned)nmb, (unsigned)slices), GW_THREADS, smem, stream>>>(p, per);
    HOIR_CHECK_CUDA(cudaGetLastError());
  }
  return HOIR_OK;
}
