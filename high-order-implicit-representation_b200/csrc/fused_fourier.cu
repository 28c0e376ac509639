// fused_fourier.cu -- whole-network kernel for Fourier-series implicit MLPs
//   in -> hid -> hid -> out,  layer_type = "fourier"  (HighOrderMLP as built at
//   /root/reference/high_order_implicit_representation/networks.py:41-55 with mlp.layer_type=fourier, mlp.n_in=31:
//   README.md:44; Quirk Q1: the hidden layer takes n_in terms, the output layer n_out or n).
//
// ONE cooperative launch per training step (Net.eval_step + backward + optimizer.step, networks.py:64-70, 93-102):
//
//   phase 1  per 256-sample tile (thread per sample, 8 sample warps + MMA warp + W-loader warp), activations stay on
//            chip from the coordinates to the loss and back:
//              layer 0 and the hidden layer:  D[256 x hid] = Phi[256 x in*32] * W   on tcgen05 (kind::tf32, 3 passes
//              A*B + A_lo*B + A*B_lo); Phi is generated chunk by chunk (one input = 32 basis columns: 1/2, sin / cos
//              pairs by the angle-addition recurrence from one sincosf) by the sample threads straight into TMEM, W
//              streams through a 4-stage smem ring (cp.async.bulk from the pre-swizzled hi/lo image);
//              MaxAbs in registers; the 3-term output layer, MSE and its backward on the CUDA cores;
//              dL/dx of the hidden layer:  T[256 x 32] = G * W_chunk^T  per input on tcgen05, contracted with the basis
//              derivatives;  MaxAbs backward.
//            The layer inputs a0, a1, a2 and the layer-output gradients g0, g1, gy of the tile are written ONCE to a
//            workspace ([tile][feature][256]: coalesced) for the weight-gradient phase.
//   phase 2  (after a grid barrier) dL/dw of all three layers: every CTA takes a slice of the batch and keeps the
//            accumulators of ALL blocks of 128 basis columns in TMEM (hidden layer 10 x 48 columns + output layer 2 x 16
//            = 512; layer 0 in a second pass over the slice).  Per unit of 32 samples G^T (K-major) is staged once,
//            Phi^T (MN-major) is regenerated two blocks at a time into a double-buffered swizzled smem stage.  The
//            tensor core's fp32 accumulation truncates, so a chain is kept to 2048 samples: the accumulators are then
//            added into the CTA's PRIVATE gradient slab (plain read-modify-write, boundary layout) -- no atomics.
//   phase 2b (after a grid barrier) the slabs are summed in a fixed order into the gradient vector (deterministic) and
//            re-zeroed.
//   phase 3  (after a grid barrier) optimizer tail: Adam / SparseLion on the flat vector, refresh of the weight images,
//            gradient zeroing, loss hand-over (opt_tail.cuh; with peers: the NVLink gradient exchange).
//
// Forward only (hoir_mlp_forward, rendering): phase 1 up to the output layer.
// Basis conventions (argument scale, sin/cos order) come from the layer descriptor (spec.py).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hoir_common.cuh"
#include "opt_tail.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

// in-kernel cycle accounting of CTA 0 (printed when HOIR_FOURIER_DBG has bit 5 set): compiled in with -DHOIR_FOURIER_TIMERS=1
#ifndef HOIR_FOURIER_TIMERS
#define HOIR_FOURIER_TIMERS 0
#endif
#if HOIR_FOURIER_TIMERS
#define FCLK() clock64()
#else
#define FCLK() 0LL
#endif

constexpr int FT = 256;                       // samples per tile == sample threads
constexpr int F_THREADS = 320, F_STAGES = 4;  // 8 sample warps + MMA warp (8) + loader warp (9)
constexpr uint32_t F_STAGE_BYTES = 16384;
constexpr int F_APITCH = 41, F_MAXH = 48;     // activation row pitch (floats); hidden width padded to 16
constexpr int F_A_COL = 0, F_D_COL = 256, F_T_COL = 256;   // (the T accumulators of dL/dx reuse the D columns: 2 x 2 x 64)
constexpr int F_UNIT = 32;                    // samples per dL/dw unit
constexpr int F_FLUSH = 64;                   // units per accumulator flush: the tensor core's fp32 accumulation truncates,
                                              // so a chain is kept to 2048 samples (measured: 1.7e-4 relative after 22 k)
// phase-2 Phi^T stage: 32 samples (rows) x 256 basis columns (two blocks of 128), hi and lo images
constexpr uint32_t F_U_RS = 256 / 32 * 512, F_U_BYTES = F_UNIT / 4 * F_U_RS;   // 4096 per 4-row group, 32768 per image
constexpr uint32_t F_G_BYTES = 16384;         // G^T unit: [hi | lo] of the hidden gradient (<= 48 rows) + [hi | lo] of dL/dy (16 rows)
enum { FB_AFULL0 = 0, FB_AFULL1, FB_AEMPTY0, FB_AEMPTY1, FB_DFULL, FB_DEMPTY, FB_GFULL, FB_TFULL0, FB_TFULL1,
       FB_TEMPTY0, FB_TEMPTY1, FB_WFULL0, FB_WEMPTY0 = FB_WFULL0 + F_STAGES, FB_UFULL0 = FB_WEMPTY0 + F_STAGES,
       FB_UEMPTY0 = FB_UFULL0 + 3, FB_GTFULL0 = FB_UEMPTY0 + 3, FB_GTFULL1, FB_GTEMPTY0, FB_GTEMPTY1, FB_GWDONE, FB_ACCFREE, FB_COUNT };
// shared memory plan (bytes): phase 1 [W ring | a1 tile | a2 tile], phase 2 [2 Phi^T stages | 2 G^T units] over it
constexpr uint32_t FS_W = 0, FS_A1 = F_STAGES * F_STAGE_BYTES, FS_A2 = FS_A1 + FT * F_APITCH * 4;
constexpr int F_USTAGES = 3;                  // Phi^T stages in flight (the workers run up to two stages ahead of the tensor pipe)
constexpr uint32_t FS_U = 0, FS_G = F_USTAGES * 2 * F_U_BYTES;
constexpr uint32_t FS_W2 = FS_G + 2 * F_G_BYTES, FS_BAR = FS_W2 + 2048, FS_TOTAL = FS_BAR + 512;
static_assert(FS_TOTAL <= 227 * 1024, "shared-memory plan");
static_assert(FS_A2 + FT * F_APITCH * 4 <= FS_W2, "phase-1 tiles end before the output-layer weights");

struct FourierParams {
  long long B;
  const float* x;                  // [B][in0] or nullptr (mesh)
  hoir_mesh_desc mesh;
  int in0, hid, out;               // features
  int n0, n1, n2;                  // terms of layer 0, hidden, output
  int outp, outp8;                 // hid rounded up to 16 / 8
  int op2;                         // out rounded up to 4
  int normalize;
  float norm_eps, arg_over_len, dth;   // basis argument = pi * arg_over_len * x;  dth = d(argument)/dx
  int odd_is_sin;
  const float* img0;               // [in0][hi | lo][outp x 32] pre-swizzled chunk images
  const float* img1;               // [hid][hi | lo][outp x 32]
  const float* w2;                 // [hid][n2][op2]
  // forward only
  float* y;
  int post_scale;
  // training
  int target_kind;
  const void* target;
  const float* gy_ext;
  float loss_scale;
  float* loss_sum;
  float* gw[3];
  float* ws;                       // workspace arrays [tile][feature][256]
  long long ws_a0, ws_a1, ws_a2, ws_g0, ws_g1, ws_gy;   // float offsets
  unsigned int* sync;              // [4]: tail barrier (0, 1), phase counter / depart counter of this kernel (2, 3)
  float* slabs;                    // [p2_ctas][n_params]: private gradient slabs (zero on entry, zero on exit)
  long long n_params, woff[4];     // boundary-layout float offsets of the layers in a slab / the flat gradient
  int p2_ctas;                     // CTAs that take a batch slice in phase 2
  int dbg;                         // HOIR_FOURIER_DBG bit 5 (with -DHOIR_FOURIER_TIMERS=1): print CTA 0's cycle accounting
  OptTail tail;
};

// phi_0 = 1/2, then slot j holds sin or cos of harmonic (j + 1) / 2 of theta = arg_scale * pi * x / length:
// one sincospif (exact range reduction) and the angle-addition recurrence.  NT > 0: the number of terms is the
// compile-time NT (no per-slot masks); ODD_SIN: odd slots hold the sines.
template <int ND, int NT, bool ODD_SIN, bool DER>
__device__ __forceinline__ void fourier_eval(const FourierParams& p, int n, float x, float* __restrict__ val,
                                             float* __restrict__ der) {
  float s1, c1;
  sincospif(x * p.arg_over_len, &s1, &c1);
  float sh = s1, ch = c1;
  val[0] = (NT > 0 || 0 < n) ? 0.5f : 0.f;
  if (DER) der[0] = 0.f;
#pragma unroll
  for (int j = 1; j < ND; ++j) {
    const bool is_sin = ((j & 1) == 1) == ODD_SIN;
    const bool on = NT > 0 ? (j < NT) : (j < n);
    val[j] = on ? (is_sin ? sh : ch) : 0.f;
    if (DER) {
      const float hf = (float)((j + 1) / 2) * p.dth;
      der[j] = on ? (is_sin ? ch * hf : -sh * hf) : 0.f;
    }
    if ((j & 1) == 0 && j + 1 < ND) {
      const float sn = fmaf(sh, c1, ch * s1), cn = fmaf(ch, c1, -sh * s1);
      sh = sn;
      ch = cn;
    }
  }
}

// shared-memory descriptor (SWIZZLE_128B_BASE32B) with an explicit stride between 32-column atoms
__device__ __forceinline__ uint64_t make_desc_lbo(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ uint32_t chunk_off(uint32_t r, uint32_t c) {   // one-atom-wide operand, rows of 128 B
  return (r >> 2) * 512u + (r & 3u) * 128u + ((c << 2) ^ ((r & 3u) << 5));
}

// MaxAbsNormalization over W registers, first index wins ties
template <int W>
__device__ __forceinline__ void maxabs_regs(float* v, int width, float eps, float& inv_d, int& kidx) {
  float m = -1.f, vk = 0.f;
  int k = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float a = fabsf(v[i]);
    if (i < width && a > m) { m = a; k = i; vk = v[i]; }
  }
  inv_d = __fdiv_rn(1.f, __fadd_rn(m, eps));
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] *= inv_d;
  kidx = vk > 0.f ? k : (vk < 0.f ? (k | 0x10000) : (k | 0x20000));
}
__device__ __forceinline__ float kidx_sign(int kidx) { return (kidx & 0x10000) ? -1.f : ((kidx & 0x20000) ? 0.f : 1.f); }

template <bool TRAIN, bool ODD_SIN, int NT>
__global__ void __launch_bounds__(F_THREADS, 1) fourier_mlp_kernel(const FourierParams p) {
  extern __shared__ __align__(1024) uint8_t fsm[];
  uint8_t* sW = fsm + FS_W;
  float* sA1 = reinterpret_cast<float*>(fsm + FS_A1);
  float* sA2 = reinterpret_cast<float*>(fsm + FS_A2);
  float* sW2 = reinterpret_cast<float*>(fsm + FS_W2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(fsm + FS_BAR);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + FB_COUNT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[FB_AFULL0], FT); mbar_init(&bars[FB_AFULL1], FT);
    mbar_init(&bars[FB_AEMPTY0], 1); mbar_init(&bars[FB_AEMPTY1], 1);
    mbar_init(&bars[FB_DFULL], 1); mbar_init(&bars[FB_DEMPTY], FT);
    mbar_init(&bars[FB_GFULL], FT);
    mbar_init(&bars[FB_TFULL0], 1); mbar_init(&bars[FB_TFULL1], 1);
    mbar_init(&bars[FB_TEMPTY0], FT); mbar_init(&bars[FB_TEMPTY1], FT);
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(&bars[FB_WFULL0 + s], 1); mbar_init(&bars[FB_WEMPTY0 + s], 1); }
    for (int k = 0; k < F_USTAGES; ++k) { mbar_init(&bars[FB_UFULL0 + k], FT); mbar_init(&bars[FB_UEMPTY0 + k], 1); }
    mbar_init(&bars[FB_GWDONE], 1);
    mbar_init(&bars[FB_GTFULL0], FT); mbar_init(&bars[FB_GTFULL1], FT);
    mbar_init(&bars[FB_GTEMPTY0], 1); mbar_init(&bars[FB_GTEMPTY1], 1);
    mbar_init(&bars[FB_ACCFREE], FT);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int k = tid; k < p.hid * p.n2 * p.op2; k += F_THREADS) sW2[k] = __ldg(p.w2 + k);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;
  const long long tiles = (p.B + FT - 1) / FT;
  const uint32_t stage_bytes = (uint32_t)p.outp * 256u;       // hi + lo image of one chunk
  const int nch0 = p.in0, nch1 = p.hid;                        // ND = 32: one input per chunk

  if (warp == 9) {
    // =========================== W loader ==============================================================
    if (elect_one()) {
      uint32_t qw = 0;
      auto load = [&](const float* img, int c) {
        const uint32_t st = qw % F_STAGES, ph = (qw / F_STAGES) & 1;
        mbar_wait(&bars[FB_WEMPTY0 + st], ph ^ 1);
        mbar_expect_tx(&bars[FB_WFULL0 + st], stage_bytes);
        bulk_g2s(sW + st * F_STAGE_BYTES, img + (size_t)c * 2 * p.outp * 32, stage_bytes, &bars[FB_WFULL0 + st]);
        ++qw;
      };
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int c = 0; c < nch0; ++c) load(p.img0, c);
        for (int c = 0; c < nch1; ++c) load(p.img1, c);
        if (TRAIN)
          for (int c = 0; c < nch1; ++c) load(p.img1, c);
      }
    }
  } else if (warp == 8) {
    // =========================== MMA warp ==============================================================
    const uint32_t leader = elect_one();
    const uint32_t id_fwd = make_idesc(128, p.outp, 0, 0);      // A: TMEM, B: W chunk K-major (rows = outputs)
    constexpr uint32_t ID_GX = make_idesc(128, 64, 0, 1);       // A: TMEM (G), B: two W chunks MN-major (N = 64 basis columns)
    uint32_t qw = 0, qa = 0, qd = 0, qt = 0, n_tile = 0;
    auto fwd_layer = [&](int nch) {
      if (qd > 0) mbar_wait(&bars[FB_DEMPTY], (qd - 1) & 1);    // the previous accumulation has been read out
      for (int c = 0; c < nch; ++c, ++qw, ++qa) {
        const uint32_t st = qw % F_STAGES, wph = (qw / F_STAGES) & 1, buf = qa & 1, ph = (qa >> 1) & 1;
        mbar_wait(&bars[FB_AFULL0 + buf], ph);
        mbar_wait(&bars[FB_WFULL0 + st], wph);
        fence_after();
        if (leader) {
          const uint32_t wbase = smem_u32(sW + st * F_STAGE_BYTES);
          const uint64_t dW[2] = {make_desc(wbase, 512), make_desc(wbase + (uint32_t)p.outp * 128u, 512)};
#pragma unroll
          for (int blk = 0; blk < 2; ++blk)
#pragma unroll
            for (int pass = 0; pass < 3; ++pass)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                mma_ts(tmem + F_D_COL + blk * 64, tmem + F_A_COL + blk * 128 + buf * 64 + (pass == 1 ? 32 : 0) + j * 8,
                       dW[pass == 2] + (uint64_t)((j * 32) >> 4), id_fwd, (c | pass | j) != 0);
          umma_commit(&bars[FB_AEMPTY0 + buf]);
          umma_commit(&bars[FB_WEMPTY0 + st]);
          if (c == nch - 1) umma_commit(&bars[FB_DFULL]);
        }
        __syncwarp();
      }
      ++qd;
    };
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      fwd_layer(nch0);
      fwd_layer(nch1);
      if (TRAIN) {
        mbar_wait(&bars[FB_GFULL], n_tile & 1);
        const int ksteps = p.outp8 >> 3;
        // two inputs (64 basis columns) per product: the W chunk images of inputs 2 cp, 2 cp + 1 sit in consecutive ring
        // stages (in0 and hid are even), addressed as ONE MN-major B operand whose column atoms are a stage apart
        for (int cp = 0; cp < nch1 / 2; ++cp, qw += 2, ++qt) {
          const uint32_t st = qw % F_STAGES, wph = (qw / F_STAGES) & 1, buf = qt & 1, ph = (qt >> 1) & 1;
          mbar_wait(&bars[FB_TEMPTY0 + buf], ph ^ 1);
          mbar_wait(&bars[FB_WFULL0 + st], wph);
          mbar_wait(&bars[FB_WFULL0 + st + 1], wph);
          fence_after();
          if (leader) {
            const uint32_t wbase = smem_u32(sW + st * F_STAGE_BYTES);
            const uint64_t dW[2] = {make_desc_lbo(wbase, F_STAGE_BYTES, 512),
                                    make_desc_lbo(wbase + (uint32_t)p.outp * 128u, F_STAGE_BYTES, 512)};
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
              for (int pass = 0; pass < 3; ++pass)
                for (int j = 0; j < ksteps; ++j)
                  mma_ts(tmem + F_T_COL + (blk * 2 + buf) * 64, tmem + F_A_COL + blk * 128 + (pass == 1 ? p.outp8 : 0) + j * 8,
                         dW[pass == 2] + (uint64_t)((j * 2 * 512) >> 4), ID_GX, (pass | j) != 0);
            umma_commit(&bars[FB_TFULL0 + buf]);
            umma_commit(&bars[FB_WEMPTY0 + st]);
            umma_commit(&bars[FB_WEMPTY0 + st + 1]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== sample warps: one thread per sample ===================================
    const int s = tid, blk = warp >> 2;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* a1 = sA1 + s * F_APITCH;
    float* a2 = sA2 + s * F_APITCH;
    uint32_t qa = 0, qd = 0, qt = 0, n_tile = 0;
    float loss_local = 0.f;
    long long tm_eval = 0, tm_wait = 0, tm_st = 0, tm_d = 0, tm_out = 0, tm_ob = 0, tm_gxe = 0, tm_gxw = 0, tm_gxr = 0, tm_all = FCLK();
    // one basis chunk (input value xv, n terms) -> TMEM A staging
    auto stage = [&](float xv, int n) {
      const uint32_t buf = qa & 1, ph = (qa >> 1) & 1;
      float v[32], lo[32];
      const long long c0 = FCLK();
      fourier_eval<32, NT, ODD_SIN, false>(p, n, xv, v, nullptr);
#pragma unroll
      for (int k = 0; k < 32; ++k) lo[k] = tf32_lo(v[k]);
      const long long c1 = FCLK();
      mbar_wait(&bars[FB_AEMPTY0 + buf], ph ^ 1);
      fence_after();
      const long long c2 = FCLK();
      const uint32_t col = tm_lane + F_A_COL + blk * 128 + buf * 64;
      tmem_st32(col, v);
      tmem_st32(col + 32, lo);
      tmem_wait_st();
      fence_before();
      mbar_arrive(&bars[FB_AFULL0 + buf]);
      const long long c3 = FCLK();
      tm_eval += c1 - c0; tm_wait += c2 - c1; tm_st += c3 - c2;
      ++qa;
    };
    // accumulator of the finished layer -> h[0 .. F_MAXH)
    auto read_d = [&](float* h) {
      const long long c0 = FCLK();
      mbar_wait(&bars[FB_DFULL], qd & 1);
      tm_d += FCLK() - c0;
      fence_after();
#pragma unroll
      for (int g = 0; g < F_MAXH; g += 16) {
        if (g < p.outp) tmem_ld16(tm_lane + F_D_COL + blk * 64 + g, h + g);
      }
      tmem_wait_ld();
      fence_before();
      mbar_arrive(&bars[FB_DEMPTY]);
      ++qd;
    };
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      const long long b = tile * FT + s;
      const bool live = b < p.B;
      float* wsa0 = p.ws + p.ws_a0 + (size_t)tile * p.in0 * FT + s;
      float* wsa1 = p.ws + p.ws_a1 + (size_t)tile * p.hid * FT + s;
      float* wsa2 = p.ws + p.ws_a2 + (size_t)tile * p.hid * FT + s;
      float* wsg0 = p.ws + p.ws_g0 + (size_t)tile * p.hid * FT + s;
      float* wsg1 = p.ws + p.ws_g1 + (size_t)tile * p.hid * FT + s;
      float* wsgy = p.ws + p.ws_gy + (size_t)tile * 4 * FT + s;
      // ---- layer 0 (the inputs pass through the idle a2 row: ONE copy of the basis code, not one per input)
      {
        float x0[HOIR_MAX_ROTATIONS * 2];
#pragma unroll
        for (int i = 0; i < HOIR_MAX_ROTATIONS * 2; ++i) x0[i] = 0.f;
        if (live) {
          if (p.x != nullptr) {
#pragma unroll
            for (int i = 0; i < HOIR_MAX_ROTATIONS * 2; ++i)
              if (i < p.in0) x0[i] = __ldg(p.x + b * p.in0 + i);
          } else {
            hoir::mesh_position(p.mesh, p.mesh.first_pixel + b, x0, HOIR_MAX_ROTATIONS);
          }
        }
#pragma unroll
        for (int i = 0; i < HOIR_MAX_ROTATIONS * 2; ++i)
          if (i < p.in0) a2[i] = x0[i];
#pragma unroll 1
        for (int i = 0; i < p.in0; ++i) {
          const float xv = a2[i];
          if (TRAIN) wsa0[(size_t)i * FT] = xv;
          stage(xv, p.n0);
        }
      }
      float h[F_MAXH];
      float inv_d0 = 1.f, inv_d1 = 1.f;
      int kidx0 = 0, kidx1 = 0;
      read_d(h);
      if (p.normalize) maxabs_regs<F_MAXH>(h, p.hid, p.norm_eps, inv_d0, kidx0);
#pragma unroll
      for (int i = 0; i < F_MAXH; ++i)
        if (i < p.hid) {
          a1[i] = h[i];
          if (TRAIN) wsa1[(size_t)i * FT] = live ? h[i] : 0.f;
        }
      // ---- hidden layer
#pragma unroll 1
      for (int c = 0; c < nch1; ++c) stage(a1[c], p.n1);
      read_d(h);
      if (p.normalize) maxabs_regs<F_MAXH>(h, p.hid, p.norm_eps, inv_d1, kidx1);
      // ---- output layer (CUDA cores): y[o] = sum_i sum_j phi_j(a2_i) W2[i][j][o]
      const long long co0 = FCLK();
      float yo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < F_MAXH; ++i)
        if (i < p.hid) {
          a2[i] = h[i];
          if (TRAIN) wsa2[(size_t)i * FT] = live ? h[i] : 0.f;
        }
#pragma unroll 8
      for (int i = 0; i < p.hid; ++i) {
        float val[4];
        fourier_eval<4, 0, ODD_SIN, false>(p, p.n2, a2[i], val, nullptr);
        const float* wr = sW2 + (size_t)i * p.n2 * p.op2;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < p.n2) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + j * p.op2);
            yo[0] = fmaf(val[j], w4.x, yo[0]);
            yo[1] = fmaf(val[j], w4.y, yo[1]);
            yo[2] = fmaf(val[j], w4.z, yo[2]);
            yo[3] = fmaf(val[j], w4.w, yo[3]);
          }
      }
      tm_out += FCLK() - co0;
      if (!TRAIN) {
        if (live) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.out) p.y[b * p.out + o] = p.post_scale ? 0.5f * (yo[o] + 1.f) : yo[o];
        }
        continue;
      }
      // ---- loss and dL/dy
      const long long cb0 = FCLK();
      float gy[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (o < p.out) {
            if (p.gy_ext != nullptr) {
              gy[o] = __ldg(p.gy_ext + b * p.out + o);
            } else {
              float t;
              if (p.target_kind == HOIR_TARGET_U8_IMAGE) {
                const float u = (float)__ldg(static_cast<const unsigned char*>(p.target) + (p.mesh.first_pixel + b) * p.out + o);
                t = __fsub_rn(__fdiv_rn(__fmul_rn(u, 2.f), 255.f), 1.f);
              } else {
                t = __ldg(static_cast<const float*>(p.target) + b * p.out + o);
              }
              const float diff = yo[o] - t;
              loss_local = fmaf(diff, diff, loss_local);
              gy[o] = 2.f * diff * p.loss_scale;
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) wsgy[(size_t)o * FT] = gy[o];
      // ---- output layer backward: dL/da2_i = sum_j phi'_j(a2_i) sum_o W2[i][j][o] gy[o];  MaxAbs backward -> g1
      float dot = 0.f;
#pragma unroll 8
      for (int i = 0; i < p.hid; ++i) {
        const float xa = a2[i];
        float val[4], der[4];
        fourier_eval<4, 0, ODD_SIN, true>(p, p.n2, xa, val, der);
        const float* wr = sW2 + (size_t)i * p.n2 * p.op2;
        float gx = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < p.n2) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + j * p.op2);
            const float t = fmaf(w4.w, gy[3], fmaf(w4.z, gy[2], fmaf(w4.y, gy[1], w4.x * gy[0])));
            gx = fmaf(der[j], t, gx);
          }
        dot = fmaf(gx, xa, dot);
        a2[i] = gx;
      }
      {
        const int k1 = kidx1 & 0xffff;
        const float sgn1 = kidx_sign(kidx1);
        // G (dL/d hidden-layer output), hi | lo, into the A staging columns: every forward product of this tile is
        // complete (D of the hidden layer was read)
        for (int g = 0; g < p.outp8; g += 8) {
          float v[8], lo[8];
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            float gi = 0.f;
            if (g + o < p.hid) {
              gi = a2[g + o];
              if (p.normalize) gi = (gi - (g + o == k1 ? sgn1 * dot : 0.f)) * inv_d1;
              wsg1[(size_t)(g + o) * FT] = gi;
            }
            v[o] = gi;
            lo[o] = tf32_lo(gi);
          }
          tmem_st8(tm_lane + F_A_COL + blk * 128 + g, v);
          tmem_st8(tm_lane + F_A_COL + blk * 128 + p.outp8 + g, lo);
        }
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[FB_GFULL]);
      }
      // ---- dL/da1 through the tensor cores: T chunk of input c contracted with the basis derivatives
      tm_ob += FCLK() - cb0;
      dot = 0.f;
#pragma unroll 1
      for (int cp = 0; cp < nch1 / 2; ++cp, ++qt) {
        const uint32_t buf = qt & 1, ph = (qt >> 1) & 1;
        const float xa0 = a1[2 * cp], xa1 = a1[2 * cp + 1];
        float val[32], der0[32], der1[32];
        const long long cg0 = FCLK();
        fourier_eval<32, NT, ODD_SIN, true>(p, p.n1, xa0, val, der0);
        fourier_eval<32, NT, ODD_SIN, true>(p, p.n1, xa1, val, der1);
        const long long cg1 = FCLK();
        mbar_wait(&bars[FB_TFULL0 + buf], ph);
        fence_after();
        const long long cg2 = FCLK();
        tm_gxe += cg1 - cg0; tm_gxw += cg2 - cg1;
        float t[32];
        const uint32_t tcol = tm_lane + F_T_COL + (blk * 2 + buf) * 64;
        tmem_ld16(tcol, t);
        tmem_ld16(tcol + 16, t + 16);
        tmem_wait_ld();
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) acc0 = fmaf(der0[q], t[q], acc0);
        tmem_ld16(tcol + 32, t);
        tmem_ld16(tcol + 48, t + 16);
        tmem_wait_ld();
        fence_before();
        mbar_arrive(&bars[FB_TEMPTY0 + buf]);
#pragma unroll
        for (int q = 0; q < 32; ++q) acc1 = fmaf(der1[q], t[q], acc1);
        dot = fmaf(acc0, xa0, dot);
        dot = fmaf(acc1, xa1, dot);
        a2[2 * cp] = acc0;
        a2[2 * cp + 1] = acc1;
        tm_gxr += FCLK() - cg2;
      }
      {
        const int k0 = kidx0 & 0xffff;
        const float sgn0 = kidx_sign(kidx0);
        for (int i = 0; i < p.hid; ++i) {
          float gi = a2[i];
          if (p.normalize) gi = (gi - (i == k0 ? sgn0 * dot : 0.f)) * inv_d0;
          wsg0[(size_t)i * FT] = gi;
        }
      }
    }
    if (HOIR_FOURIER_TIMERS && (p.dbg & 32) && blockIdx.x == 0 && tid == 0)
      printf("tiles %u all %lld eval %lld wait %lld st %lld dwait %lld out %lld outbwd %lld gx: eval %lld wait %lld rest %lld\n", n_tile, FCLK() - tm_all, tm_eval, tm_wait, tm_st, tm_d, tm_out, tm_ob, tm_gxe, tm_gxw, tm_gxr);
    if (TRAIN && p.gy_ext == nullptr) {
      const float sum = hoir::warp_sum(loss_local) * p.loss_scale;
      if (lane == 0 && sum != 0.f) hoir::red_add(p.loss_sum, sum);
    }
  }
  if (!TRAIN) {
    fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    return;
  }
  // =============================== phase 2: dL/dw of the three layers ===================================
  const long long tp0 = FCLK();
  fence_before();
  hoir::grid_sync_phase(p.sync + 2, 1);
  fence_after();
  const long long tp1 = FCLK();
  if ((int)blockIdx.x < p.p2_ctas) {
    const long long units_all = tiles * (FT / F_UNIT);
    const long long per = (units_all + p.p2_ctas - 1) / p.p2_ctas;
    const long long u_lo = per * blockIdx.x, u_hi = u_lo + per < units_all ? u_lo + per : units_all;
    const int units = u_hi > u_lo ? (int)(u_hi - u_lo) : 0;
    uint8_t* sU = fsm + FS_U;
    uint8_t* sG = fsm + FS_G;
    float* slab = p.slabs + (size_t)blockIdx.x * p.n_params;
    const int nb0 = (p.in0 + 3) / 4, nb1 = (p.hid + 3) / 4, nch2 = (p.hid + 7) / 8, nb2 = (nch2 + 3) / 4;
    const uint32_t gh_bytes = (uint32_t)p.outp * 128u;           // one image (hi or lo) of the hidden G^T
    uint32_t qu = 0, qg = 0, qf = 0;                             // stage / G-unit / flush counters (same in every role)
    // pass 0: hidden layer (block pairs) + output layer; pass 1: layer 0.  A stage = up to two blocks of one layer.
    for (int pass = 0; pass < 2; ++pass) {
      const int ns_a = pass == 0 ? (nb1 + 1) / 2 : (nb0 + 1) / 2;      // stages of the first layer of the pass
      const int ns = ns_a + (pass == 0 ? (nb2 + 1) / 2 : 0);
      // stage st -> (layer, first block, accumulator columns)
      auto stage_layer = [&](int st) { return pass == 0 ? (st < ns_a ? 1 : 2) : 0; };
      auto stage_mb = [&](int st) { return 2 * (pass == 0 && st >= ns_a ? st - ns_a : st); };
      auto acc_col = [&](int layer, int mb) { return layer == 2 ? nb1 * p.outp + mb * 16 : mb * p.outp; };
      if (warp == 8) {
        const uint32_t leader = elect_one();
        for (int u = 0; u < units; ++u, ++qg) {
          const uint32_t gbuf = qg & 1, gph = (qg >> 1) & 1;
          if (u > 0 && u % F_FLUSH == 0) {
            mbar_wait(&bars[FB_ACCFREE], (qf - 1) & 1);          // accumulators read out
          }
          mbar_wait(&bars[FB_GTFULL0 + gbuf], gph);
          for (int st = 0; st < ns; ++st, ++qu) {
            const uint32_t ubuf = qu % F_USTAGES, uph = (qu / F_USTAGES) & 1;
            mbar_wait(&bars[FB_UFULL0 + ubuf], uph);
            fence_after();
            if (leader) {
              const int layer = stage_layer(st), mb0 = stage_mb(st);
              const int nb = layer == 0 ? nb0 : (layer == 1 ? nb1 : nb2);
              const uint32_t ubase = smem_u32(sU + ubuf * 2 * F_U_BYTES);
              const uint32_t gbase = smem_u32(sG + gbuf * F_G_BYTES) + (layer == 2 ? 2 * gh_bytes : 0);
              const uint32_t glo = layer == 2 ? 16u * 128u : gh_bytes;
              const uint64_t dG[2] = {make_desc(gbase, 512), make_desc(gbase + glo, 512)};
              const uint32_t idesc = make_idesc(128, layer == 2 ? 16 : p.outp, 1, 0);   // A: Phi^T MN-major, B: G^T K-major
              for (int h = 0; h < 2 && mb0 + h < nb; ++h) {
                const uint64_t dU[2] = {make_desc(ubase + h * 2048, F_U_RS), make_desc(ubase + F_U_BYTES + h * 2048, F_U_RS)};
                const uint32_t d_col = tmem + acc_col(layer, mb0 + h);
#pragma unroll
                for (int ps = 0; ps < 3; ++ps)
#pragma unroll
                  for (int j = 0; j < F_UNIT / 8; ++j)
                    mma_ss(d_col, dU[ps == 1] + (uint64_t)((j * 2 * F_U_RS) >> 4), dG[ps == 2] + (uint64_t)((j * 32) >> 4),
                           idesc, ((uint32_t)(u % F_FLUSH) | ps | j) != 0);
              }
              umma_commit(&bars[FB_UEMPTY0 + ubuf]);
              if (st == ns - 1) {
                umma_commit(&bars[FB_GTEMPTY0 + gbuf]);
                if ((u + 1) % F_FLUSH == 0 || u == units - 1) umma_commit(&bars[FB_GWDONE]);
              }
            }
            __syncwarp();
          }
          if ((u + 1) % F_FLUSH == 0 || u == units - 1) ++qf;
        }
      } else if (warp < 8) {
        long long t_ld = 0, t_gw = 0, t_gs = 0, t_ev = 0, t_uw = 0, t_us = 0, t_l2 = 0, t_fl = 0, t_o2 = 0;
        const long long t_pass0 = FCLK();
        constexpr int GV = (F_MAXH * F_UNIT) / FT, YV = (16 * F_UNIT) / FT, XV1 = (F_MAXH / 4 + 1) / 2, XV2 = (8 * 8 * F_UNIT) / FT;
        struct UnitRegs { float g[GV], y[YV], x1[XV1], x2[XV2]; };
        UnitRegs cur, nxt;
        auto issue_loads = [&](long long gu_, UnitRegs& q) {
          const long long tile_ = gu_ / (FT / F_UNIT);
          const int sub_ = (int)(gu_ - tile_ * (FT / F_UNIT)) * F_UNIT;
          const float* ga = p.ws + (pass == 0 ? p.ws_g1 : p.ws_g0) + (size_t)tile_ * p.hid * FT + sub_;
#pragma unroll
          for (int r = 0; r < GV; ++r) {
            const int it = tid + r * FT, sl = it % F_UNIT, o = it / F_UNIT;
            q.g[r] = o < p.hid ? __ldcg(ga + (size_t)o * FT + sl) : 0.f;
          }
          const int in_f = pass == 0 ? p.hid : p.in0;
          const float* xa = p.ws + (pass == 0 ? p.ws_a1 : p.ws_a0) + (size_t)tile_ * in_f * FT + sub_;
#pragma unroll
          for (int r = 0; r < XV1; ++r) {
            const int i = (2 * r) * 4 + tid / F_UNIT;            // stage r: first block 2 r, column atom tid / 32
            q.x1[r] = (r < ns_a && i < in_f) ? __ldcg(xa + (size_t)i * FT + (tid % F_UNIT)) : 0.f;
          }
          const float* gya = p.ws + p.ws_gy + (size_t)tile_ * 4 * FT + sub_;
#pragma unroll
          for (int r = 0; r < YV; ++r) {
            const int it = tid + r * FT, sl = it % F_UNIT, o = it / F_UNIT;
            q.y[r] = (pass == 0 && o < p.out) ? __ldcg(gya + (size_t)o * FT + sl) : 0.f;
          }
          const float* x2 = p.ws + p.ws_a2 + (size_t)tile_ * p.hid * FT + sub_;
#pragma unroll
          for (int r = 0; r < XV2; ++r) {
            const int it = tid + r * FT, sl = it % F_UNIT, i = it / F_UNIT;     // (first block pair of the output layer)
            q.x2[r] = (pass == 0 && i < p.hid) ? __ldcg(x2 + (size_t)i * FT + sl) : 0.f;
          }
        };
        if (units > 0) issue_loads(u_lo, cur);
        for (int u = 0; u < units; ++u, ++qg) {
          const long long k0 = FCLK();
          const uint32_t gbuf = qg & 1, gph = (qg >> 1) & 1;
          const long long gu = u_lo + u, tile = gu / (FT / F_UNIT);
          const int sub = (int)(gu - tile * (FT / F_UNIT)) * F_UNIT;
          // ---- the global loads of unit u + 1 are issued now, those of unit u arrived during unit u - 1
          issue_loads(u + 1 < units ? gu + 1 : gu, nxt);
          // ---- G^T of the unit: element (output o, sample sl), sl fastest
          const long long k1 = FCLK();
          mbar_wait(&bars[FB_GTEMPTY0 + gbuf], gph ^ 1);
          const long long k2 = FCLK();
          {
            uint8_t* gb = sG + gbuf * F_G_BYTES;
#pragma unroll
            for (int r = 0; r < GV; ++r) {
              const int it = tid + r * FT, sl = it % F_UNIT, o = it / F_UNIT;
              if (o < p.outp) {
                const uint32_t off = chunk_off((uint32_t)o, (uint32_t)sl);
                HOIR_DBG_ASSERT(off + 4u <= gh_bytes && 2 * gh_bytes + 2 * 16 * 128 <= F_G_BYTES);
                *reinterpret_cast<float*>(gb + off) = cur.g[r];
                *reinterpret_cast<float*>(gb + gh_bytes + off) = tf32_lo(cur.g[r]);
              }
            }
            if (pass == 0) {
#pragma unroll
              for (int r = 0; r < YV; ++r) {
                const int it = tid + r * FT, sl = it % F_UNIT, o = it / F_UNIT;
                const uint32_t off = 2 * gh_bytes + chunk_off((uint32_t)o, (uint32_t)sl);
                *reinterpret_cast<float*>(gb + off) = cur.y[r];
                *reinterpret_cast<float*>(gb + 16 * 128 + off) = tf32_lo(cur.y[r]);
              }
            }
          }
          fence_async_smem();
          mbar_arrive(&bars[FB_GTFULL0 + gbuf]);
          const long long k3 = FCLK();
          t_ld += k1 - k0; t_gw += k2 - k1; t_gs += k3 - k2;
          // ---- Phi^T stages
          for (int st = 0; st < ns; ++st, ++qu) {
            const uint32_t ubuf = qu % F_USTAGES, uph = (qu / F_USTAGES) & 1;
            const int layer = stage_layer(st), mb0 = stage_mb(st);
            uint8_t* ur = sU + ubuf * 2 * F_U_BYTES;
            uint8_t* ul = ur + F_U_BYTES;
            if (layer != 2) {
              // 32 slots per input: items (column atom ca = 0 .. 7 of the stage, sample sl): one input each.  The basis
              // is evaluated BEFORE the wait for the stage buffer (only the stores sit behind the tensor pipe).
              const int in_f = layer == 0 ? p.in0 : p.hid, nterm = layer == 0 ? p.n0 : p.n1;
              {
                const int sl = tid % F_UNIT, ca = tid / F_UNIT;
                const int i = mb0 * 4 + ca;
                const bool on = i < in_f;
                const long long e0 = FCLK();
                float xv = 0.f;
#pragma unroll
                for (int r = 0; r < XV1; ++r) xv = r == st ? cur.x1[r] : xv;
                float val[32];
                if (on) {
                  fourier_eval<32, NT, ODD_SIN, false>(p, nterm, xv, val, nullptr);
                } else {
#pragma unroll
                  for (int q = 0; q < 32; ++q) val[q] = 0.f;
                }
                // lanes 4 .. 7 of every 8 swap neighbouring 16-byte groups: the 8 lanes of a quarter warp then store
                // to 8 different bank groups (rows 4 apart alias in the swizzled layout: 2-way conflicts otherwise)
                const bool swp = (tid >> 2) & 1;
#pragma unroll
                for (int q8 = 0; q8 < 4; ++q8)
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float a = val[8 * q8 + e], b2 = val[8 * q8 + 4 + e];
                    val[8 * q8 + e] = swp ? b2 : a;
                    val[8 * q8 + 4 + e] = swp ? a : b2;
                  }
                const uint32_t rowoff = (uint32_t)(sl >> 2) * F_U_RS + (uint32_t)(sl & 3) * 128u + (uint32_t)ca * 512u;
                const uint32_t sw = ((uint32_t)(sl & 3) << 5) ^ (swp ? 16u : 0u);
                const long long e1 = FCLK();
                mbar_wait(&bars[FB_UEMPTY0 + ubuf], uph ^ 1);
                const long long e2 = FCLK();
                t_ev += e1 - e0; t_uw += e2 - e1;
                const long long e3 = FCLK();
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                  const uint32_t off = rowoff + (((uint32_t)q4 << 4) ^ sw);
                  HOIR_DBG_ASSERT(off + 16u <= F_U_BYTES);
                  *reinterpret_cast<float4*>(ur + off) = make_float4(val[4 * q4], val[4 * q4 + 1], val[4 * q4 + 2], val[4 * q4 + 3]);
                  *reinterpret_cast<float4*>(ul + off) = make_float4(tf32_lo(val[4 * q4]), tf32_lo(val[4 * q4 + 1]),
                                                                     tf32_lo(val[4 * q4 + 2]), tf32_lo(val[4 * q4 + 3]));
                }
                t_l2 += FCLK() - e3;
              }
            } else {
              // 4 slots per input, 8 inputs per column atom: items (atom ca, input m of the atom, sample sl)
              const float* xa = p.ws + p.ws_a2 + (size_t)tile * p.hid * FT + sub;
              const long long l0 = FCLK();
              mbar_wait(&bars[FB_UEMPTY0 + ubuf], uph ^ 1);
              const long long l1 = FCLK();
              t_uw += l1 - l0;
#pragma unroll
              for (int r = 0; r < XV2; ++r) {
                const int it = tid + r * FT;
                const int sl = it % F_UNIT, r2 = it / F_UNIT, m = r2 % 8, ca = r2 / 8;
                const int c = mb0 * 4 + ca, i = c * 8 + m;
                float val[4];
                if (c < nch2 && i < p.hid) {
                  const float xv = mb0 == 0 ? cur.x2[r] : __ldcg(xa + (size_t)i * FT + sl);
                  fourier_eval<4, 0, ODD_SIN, false>(p, p.n2, xv, val, nullptr);
                } else {
                  val[0] = val[1] = val[2] = val[3] = 0.f;
                }
                const uint32_t off = (uint32_t)(sl >> 2) * F_U_RS + (uint32_t)(sl & 3) * 128u + (uint32_t)ca * 512u +
                                     (((uint32_t)m << 4) ^ ((uint32_t)(sl & 3) << 5));
                HOIR_DBG_ASSERT(off + 16u <= F_U_BYTES);
                *reinterpret_cast<float4*>(ur + off) = make_float4(val[0], val[1], val[2], val[3]);
                *reinterpret_cast<float4*>(ul + off) = make_float4(tf32_lo(val[0]), tf32_lo(val[1]), tf32_lo(val[2]), tf32_lo(val[3]));
              }
              t_o2 += FCLK() - l1;
            }
            const long long s0 = FCLK();
            fence_async_smem();
            mbar_arrive(&bars[FB_UFULL0 + ubuf]);
            t_us += FCLK() - s0;
          }
          const long long f0 = FCLK();
          // ---- accumulators -> the CTA's private gradient slab (every F_FLUSH units and at the end of the pass)
          if ((u + 1) % F_FLUSH == 0 || u == units - 1) {
            mbar_wait(&bars[FB_GWDONE], qf & 1);
            ++qf;
            fence_after();
            const int rt = tid & 127;                             // accumulator lane = basis column of the block
            const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            for (int st = 0; st < ns; ++st) {
              const int layer = stage_layer(st), mb0 = stage_mb(st);
              const int nb = layer == 0 ? nb0 : (layer == 1 ? nb1 : nb2);
              const int nd = layer == 2 ? 4 : 32, ipc = 32 / nd;
              const int in_f = layer == 0 ? p.in0 : p.hid, out_f = layer == 2 ? p.out : p.hid;
              const int nterm = layer == 0 ? p.n0 : (layer == 1 ? p.n1 : p.n2), nch = (in_f + ipc - 1) / ipc;
              const int outp = layer == 2 ? 16 : p.outp;
              float* gl = slab + p.woff[layer];
              for (int h = 0; h < 2 && mb0 + h < nb; ++h) {
                const int c = (mb0 + h) * 4 + (rt >> 5), within = rt & 31, m = within / nd, q = within - m * nd, i = c * ipc + m;
                const bool valid = c < nch && i < in_f && q < nterm;
                for (int g = (warp >> 2) * 16; g < outp; g += 32) {
                  float d[16];
                  tmem_ld16(tm_lane + acc_col(layer, mb0 + h) + g, d);
                  tmem_wait_ld();
                  if (valid) {              // fire-and-forget adds into the PRIVATE slab: no load round trips; one add
#pragma unroll                      // per address and flush, so the order of the sum is the order of the flushes
                    for (int o = 0; o < 16; ++o)
                      if (g + o < out_f && d[o] != 0.f) hoir::red_add(gl + ((size_t)(g + o) * in_f + i) * nterm + q, d[o]);
                  }
                }
              }
            }
            fence_before();
            mbar_arrive(&bars[FB_ACCFREE]);
          }
          t_fl += FCLK() - f0;
          cur = nxt;
        }
        if (HOIR_FOURIER_TIMERS && (p.dbg & 32) && blockIdx.x == 0 && tid == 0)
          printf("pass %d units %d total %lld: loads %lld gwait %lld gstage %lld eval %lld uwait %lld stores %lld outlayer-stage %lld fence+arrive %lld flush %lld\n", pass, units, FCLK() - t_pass0, t_ld, t_gw, t_gs, t_ev, t_uw, t_l2, t_o2, t_us, t_fl);
      }
      // between the passes: every accumulator of pass 0 has been read (the workers' last flush), and the MMA warp
      // must not start pass 1 before that
      __syncthreads();
      if (warp == 8 && units > 0) {
        mbar_wait(&bars[FB_ACCFREE], (qf - 1) & 1);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  // =============================== phase 2b: slabs -> gradient vector, fixed order ========================
  const long long tp2 = FCLK();
  hoir::grid_sync_phase(p.sync + 2, 2);
  const long long tp3 = FCLK();
  {
    // a thread sums 4 consecutive parameters over the slabs (16-byte loads, 8 in flight, fixed order)
    const long long gthreads = (long long)gridDim.x * F_THREADS, n4 = p.n_params >> 2;      // n_params % 4 == 0 (host)
    for (long long k4 = blockIdx.x * (long long)F_THREADS + tid; k4 < n4; k4 += gthreads) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c0 = 0; c0 < p.p2_ctas; c0 += 8) {
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          v[c] = c0 + c < p.p2_ctas ? __ldcg(reinterpret_cast<const float4*>(p.slabs + (size_t)(c0 + c) * p.n_params) + k4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          sum.x += v[c].x; sum.y += v[c].y; sum.z += v[c].z; sum.w += v[c].w;
          if (c0 + c < p.p2_ctas)
            reinterpret_cast<float4*>(p.slabs + (size_t)(c0 + c) * p.n_params)[k4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const float sv[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long k = 4 * k4 + e;
        if (k < p.woff[3]) {
          const int l = k >= p.woff[2] ? 2 : (k >= p.woff[1] ? 1 : 0);
          float* gp = p.gw[l] + (k - p.woff[l]);
          *gp = __ldcg(gp) + sv[e];
        }
      }
    }
  }
  hoir::grid_sync_rearm(p.sync + 2, p.sync + 3);
  const long long tp4 = FCLK();
  hoir::opt_tail_in_kernel(p.tail);
  if (HOIR_FOURIER_TIMERS && (p.dbg & 32) && blockIdx.x == 0 && tid == 0)
    printf("barrier1 %lld phase2 %lld barrier2 %lld reduce %lld tail %lld\n", tp1 - tp0, tp2 - tp1, tp3 - tp2, tp4 - tp3, FCLK() - tp4);
}

// ---- weight images / packed output layer ------------------------------------------------------------------
__global__ void fourier_pack_image_kernel(const float* __restrict__ w, float* __restrict__ img, int in_f, int out_f,
                                          int n, int outp) {
  const long long total = (long long)in_f * outp * 32;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(k & 31);
    const long long r = k >> 5;
    const int o = (int)(r % outp), c = (int)(r / outp);
    const float v = (col < n && o < out_f) ? w[((size_t)o * in_f + c) * n + col] : 0.f;
    float* chunk = img + (size_t)c * 2 * outp * 32;
    const uint32_t off = chunk_off((uint32_t)o, (uint32_t)col) >> 2;
    chunk[off] = v;
    chunk[(size_t)outp * 32 + off] = tf32_lo(v);
  }
}
__global__ void fourier_pack_plain_kernel(const float* __restrict__ w, float* __restrict__ packed, int in_f, int out_f,
                                          int n, int op) {
  const int total = in_f * n * op;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    const int o = k % op, r = k / op, q = r % n, i = r / n;
    packed[k] = o < out_f ? w[((size_t)o * in_f + i) * n + q] : 0.f;
  }
}

struct Geo {
  int in0, hid, out, n0, n1, n2, outp, outp8, op2;
  long long off[4];        // packed float offsets of layers 0, 1, 2 and the total
};
bool geometry(const hoir_mlp_desc* m, Geo* g) {
  static const bool off = getenv("HOIR_NO_FOURIER_FUSED") != nullptr;
  if (off || m == nullptr || m->num_layers != 3) return false;
  for (int l = 0; l < 3; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    if (d.kind != HOIR_FOURIER || d.length != 2.f || d.periodicity > 0.f) return false;
    if (d.rescale != 0.f && d.rescale != 1.f) return false;
    if (l > 0 && d.in_features != m->layers[l - 1].out_features) return false;
    if (d.fourier_arg_scale != m->layers[0].fourier_arg_scale || d.fourier_odd_is_sin != m->layers[0].fourier_odd_is_sin) return false;
  }
  const hoir_layer_desc &l0 = m->layers[0], &l1 = m->layers[1], &l2 = m->layers[2];
  if (l0.in_features < 1 || l0.in_features > 2 * HOIR_MAX_ROTATIONS) return false;
  if (l0.out_features != l1.out_features || l0.out_features < 8 || l0.out_features > 40) return false;   // a row per thread: 40 registers
  if ((l0.in_features | l0.out_features) & 1) return false;       // dL/dx takes the hidden inputs two at a time (W ring stage pairs)
  if (l0.n < 1 || l0.n > 32 || l1.n < 1 || l1.n > 32 || l2.n < 1 || l2.n > 4) return false;
  if (l2.out_features < 1 || l2.out_features > 4) return false;
  g->in0 = l0.in_features; g->hid = l0.out_features; g->out = l2.out_features;
  g->n0 = l0.n; g->n1 = l1.n; g->n2 = l2.n;
  g->outp = (g->hid + 15) & ~15; g->outp8 = (g->hid + 7) & ~7; g->op2 = (g->out + 3) & ~3;
  g->off[0] = 0;
  g->off[1] = g->off[0] + (long long)g->in0 * 2 * g->outp * 32;
  g->off[2] = g->off[1] + (long long)g->hid * 2 * g->outp * 32;
  g->off[3] = g->off[2] + (((long long)g->hid * g->n2 * g->op2 + 3) & ~3LL);
  if ((size_t)g->hid * g->n2 * g->op2 * 4 > 2048) return false;      // the output-layer table in shared memory
  return true;
}

int fill(const hoir_mlp_desc* m, const Geo& g, long long B, const float* x, const hoir_mesh_desc* mesh, const float* packed,
         FourierParams* p, const char* where) {
  memset(p, 0, sizeof(*p));
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", where);
  if ((x == nullptr) == (mesh == nullptr) && B > 0)
    return hoir_set_error(HOIR_EINVAL, "exactly one of x / mesh must be given", where);
  if (!packed && B > 0) return hoir_set_error(HOIR_EINVAL, "null packed weights", where);
  if (mesh) {
    if (mesh->rotations * 2 != g.in0 || mesh->rows < 1 || mesh->cols < 1 || mesh->normalize_mode < 0 ||
        mesh->normalize_mode > 2 || mesh->first_pixel < 0 || mesh->first_pixel + B > (long long)mesh->rows * mesh->cols)
      return hoir_set_error(HOIR_EINVAL, "mesh descriptor does not match the network / batch", where);
    p->mesh = *mesh;
  }
  p->B = B;
  p->x = x;
  p->in0 = g.in0; p->hid = g.hid; p->out = g.out;
  p->n0 = g.n0; p->n1 = g.n1; p->n2 = g.n2;
  p->outp = g.outp; p->outp8 = g.outp8; p->op2 = g.op2;
  p->normalize = m->normalize;
  p->norm_eps = m->norm_eps;
  const float arg_scale = m->layers[0].fourier_arg_scale == 0.f ? 2.f : m->layers[0].fourier_arg_scale;
  FourierTable ft;
  hoir_make_fourier_table(2, arg_scale, &ft);
  p->arg_over_len = arg_scale / m->layers[0].length;      // basis argument = pi * arg_over_len * x
  p->dth = ft.c[1] / m->layers[0].length;                 // fl32(arg_scale * pi) / length, as the per-layer kernels
  p->odd_is_sin = m->layers[0].fourier_odd_is_sin;
  p->img0 = packed + g.off[0];
  p->img1 = packed + g.off[1];
  p->w2 = packed + g.off[2];
  return HOIR_OK;
}

const void* pick_kernel(bool train, bool odd_sin, bool n31) {
  if (train) {
    if (odd_sin) return n31 ? (const void*)fourier_mlp_kernel<true, true, 31> : (const void*)fourier_mlp_kernel<true, true, 0>;
    return n31 ? (const void*)fourier_mlp_kernel<true, false, 31> : (const void*)fourier_mlp_kernel<true, false, 0>;
  }
  if (odd_sin) return n31 ? (const void*)fourier_mlp_kernel<false, true, 31> : (const void*)fourier_mlp_kernel<false, true, 0>;
  return n31 ? (const void*)fourier_mlp_kernel<false, false, 31> : (const void*)fourier_mlp_kernel<false, false, 0>;
}

}  // namespace

// ---- entry points used by fused_mlp.cu (the C ABI dispatches Fourier networks here) -------------------------
bool hoir_fourier_supported(const hoir_mlp_desc* m) {
  Geo g;
  return geometry(m, &g);
}
long long hoir_fourier_packed_offset(const hoir_mlp_desc* m, int layer) {
  Geo g;
  if (!geometry(m, &g) || layer < 0 || layer > 3) return -1;
  return g.off[layer];
}
int hoir_fourier_pack(const hoir_mlp_desc* m, const float* const* w, float* packed, cudaStream_t st) {
  Geo g;
  if (!geometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused Fourier network", "hoir_mlp_pack");
  const int sms = hoir_num_sms();
  fourier_pack_image_kernel<<<sms, 256, 0, st>>>(w[0], packed + g.off[0], g.in0, g.hid, g.n0, g.outp);
  fourier_pack_image_kernel<<<sms * 2, 256, 0, st>>>(w[1], packed + g.off[1], g.hid, g.hid, g.n1, g.outp);
  fourier_pack_plain_kernel<<<4, 256, 0, st>>>(w[2], packed + g.off[2], g.hid, g.out, g.n2, g.op2);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
// image geometry of the tail's packed refresh (opt_tail.cuh): layers 0, 1 are chunk images, layer 2 is plain
void hoir_fourier_fill_tail(const hoir_mlp_desc* m, OptTail* t) {
  Geo g;
  if (!geometry(m, &g)) return;
  t->img_mask = 3;
  for (int l = 0; l < 2; ++l) {
    t->img_nd[l] = 32;
    t->img_ipc[l] = 1;
    t->img_outp[l] = g.outp;
  }
  for (int l = 0; l < 3; ++l) t->poff[l] = g.off[l];
  t->alt_off = -1;
}

int hoir_fourier_forward(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                         const float* packed, float* y, int post_scale, cudaStream_t st) {
  Geo g;
  if (!geometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused Fourier network", "hoir_mlp_forward");
  FourierParams p;
  if (int err = fill(m, g, B, x, mesh, packed, &p, "hoir_mlp_forward")) return err;
  if (B == 0) return HOIR_OK;
  if (!y) return hoir_set_error(HOIR_EINVAL, "null output pointer", "hoir_mlp_forward");
  p.y = y;
  p.post_scale = post_scale;
  if (getenv("HOIR_FOURIER_DBG") != nullptr) p.dbg = atoi(getenv("HOIR_FOURIER_DBG"));
  const void* kern = pick_kernel(false, p.odd_is_sin != 0, g.n0 == 31 && g.n1 == 31);
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_TOTAL));
  const long long tiles = (B + FT - 1) / FT;
  const int sms = hoir_num_sms();
  void* args[] = {&p};
  HOIR_CHECK_CUDA(cudaLaunchKernel(kern, dim3((unsigned)(tiles < sms ? tiles : sms)), dim3(F_THREADS), args, FS_TOTAL, st));
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

int hoir_fourier_train(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh, int target_kind,
                       const void* target, const float* gy_ext, const float* packed, float* const* gw, float* loss_sum,
                       float loss_scale, const OptTail* tail, cudaStream_t st, const char* where) {
  Geo g;
  if (!geometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused Fourier network", where);
  FourierParams p;
  if (int err = fill(m, g, B, x, mesh, packed, &p, where)) return err;
  if (!gw) return hoir_set_error(HOIR_EINVAL, "null gradient pointer table", where);
  for (int l = 0; l < 3; ++l) {
    if (!gw[l]) return hoir_set_error(HOIR_EINVAL, "null gradient pointer", where);
    p.gw[l] = gw[l];
  }
  if (gy_ext == nullptr) {
    if (!target || !loss_sum) return hoir_set_error(HOIR_EINVAL, "target / loss_sum required", where);
    if (target_kind != HOIR_TARGET_F32 && target_kind != HOIR_TARGET_U8_IMAGE)
      return hoir_set_error(HOIR_EINVAL, "unknown target kind", where);
    if (target_kind == HOIR_TARGET_U8_IMAGE && (mesh == nullptr || g.out != 3))
      return hoir_set_error(HOIR_EINVAL, "uint8 image targets need a mesh and 3 outputs", where);
  }
  p.target_kind = target_kind;
  p.target = target;
  p.gy_ext = gy_ext;
  p.loss_sum = loss_sum;
  p.loss_scale = loss_scale;
  const long long tiles = (B + FT - 1) / FT;
  // workspace: a0, a1, a2, g0, g1, gy as [tile][feature][256]
  const long long per_tile = (long long)(g.in0 + 4 * g.hid + 4) * FT;
  void* ws = nullptr;
  if (int err = hoir_get_workspace(0, (size_t)(tiles > 0 ? tiles : 1) * per_tile * sizeof(float), &ws)) return err;
  p.ws = static_cast<float*>(ws);
  p.ws_a0 = 0;
  p.ws_a1 = p.ws_a0 + tiles * g.in0 * FT;
  p.ws_a2 = p.ws_a1 + tiles * g.hid * FT;
  p.ws_g0 = p.ws_a2 + tiles * g.hid * FT;
  p.ws_g1 = p.ws_g0 + tiles * g.hid * FT;
  p.ws_gy = p.ws_g1 + tiles * g.hid * FT;
  void* sy = nullptr;
  if (int err = hoir_get_workspace(2, 64, &sy)) return err;
  p.sync = static_cast<unsigned int*>(sy);
  if (tail != nullptr) {
    p.tail = *tail;
    p.tail.sync = p.sync;
  }
  const int sms = hoir_num_sms();
  const int grid = sms;     // every CTA takes part in the weight-gradient phase and the tail, whatever the batch
  // phase 2: batch slices of >= 8 units (256 samples) per CTA; one private gradient slab per slice
  {
    const long long units_all = tiles * (FT / F_UNIT);
    long long n2 = (units_all + 7) / 8;
    p.p2_ctas = (int)(n2 < 1 ? 1 : (n2 > grid ? grid : n2));
    if (getenv("HOIR_FOURIER_DBG") != nullptr) p.dbg = atoi(getenv("HOIR_FOURIER_DBG"));
    p.woff[0] = 0;
    p.woff[1] = p.woff[0] + (long long)g.hid * g.in0 * g.n0;
    p.woff[2] = p.woff[1] + (long long)g.hid * g.hid * g.n1;
    p.woff[3] = p.woff[2] + (long long)g.out * g.hid * g.n2;
    p.n_params = (p.woff[3] + 3) & ~3LL;       // slab stride (16-byte rows); the pad entries stay zero
    void* sl = nullptr;
    if (int err = hoir_get_workspace(4, (size_t)grid * p.n_params * sizeof(float), &sl)) return err;   // zeroed when allocated
    p.slabs = static_cast<float*>(sl);
  }
  const void* kern = pick_kernel(true, p.odd_is_sin != 0, g.n0 == 31 && g.n1 == 31);
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_TOTAL));
  void* args[] = {&p};
  HOIR_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(F_THREADS), args, FS_TOTAL, st));
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
