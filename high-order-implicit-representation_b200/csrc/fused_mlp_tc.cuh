// fused_mlp_tc.cuh -- tensor-core (tcgen05 / TMEM) kernels for the 40-wide implicit MLP: the whole-network
// training kernel (continuous and discontinuous n = 3) and the forward-only kernel.
// Included by fused_mlp.cu (uses its FusedParams / SegConst / contract_* helpers).  DESIGN.md section 4.1.
//
// Training: one CTA = 8 sample warps + 1 MMA warp + 3 helper warps, persistent cooperative grid (1 CTA per SM).
// A tile is 128 samples; sample s is worked on by a PAIR of threads (s, s + 128) in warps w and w + 4 -- both map
// to TMEM lane s -- each owning 20 of the 40 hidden units.  The hidden layer's three contractions are densified
// GEMMs (K = 40 inputs x 5 / 6 node slots) on tcgen05.mma kind::tf32 with a 3-pass hi/lo split
// (A*B + Alo*B + A*Blo, fp32-level accuracy):
//   FWD  D1[128 x 40]  = Phi1[128 x K] * W1         A: TMEM (chunks of 40 / 24 K-columns, double buffered, written
//                                                     by the sample threads from the layer-0 registers; four of the
//                                                     five chunks of tile n + 1 are staged during tile n)
//                                                    B: smem W1, K-major view
//   GX   T [128 x K]   = G  [128 x 40] * W1^T        A: TMEM, B: the same smem W1, MN-major view
//   GW   gW1[K x 40]  += Phi1^T[K x 8] * G[8 x 40]   A: smem units of 8 samples (MN-major) staged by the HELPER warps,
//                                                    B: smem G^T (K-major); accumulators stay in TMEM for the whole
//                                                     kernel.  The MMA warp issues by readiness (units of tile n - 1
//                                                     between the forward chunks of tile n).
// All shared-memory operands use the 128B_BASE32B layout (rows of 128 B, atoms of 4 rows, 32-byte chunks XOR-ed
// with row % 4), the one layout that tcgen05 accepts for tf32 in BOTH majors (csrc/probe/umma_probe.cu), so W1 and
// G are stored once.  Layer 0, the output layer, the basis, MaxAbs and the loss stay on the CUDA cores; the helper
// warps also own dL/dw2 (register sums for the whole kernel).  The kernel is templated on the threads per sample
// (NQ = 2; NQ = 4 is a tested but slower variant).
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int TS = 128;          // samples per tile
constexpr int NSAMPLE_THREADS = 256;   // forward kernel: 8 sample warps, 2 threads per sample
constexpr int NHELPER = 96;            // 3 helper warps: stage the Phi1 units of the dL/dw1 product
constexpr int NTHREADS_FWD = 288;      // forward-only kernel: 8 sample warps + 1 MMA warp
constexpr int APITCH = 41;       // row pitch (floats) of the saved-activation tiles

// Shape constants of the tensor path (continuous n = 3, 2 hidden/output segments, width 40)
// NQ = threads per sample (2: pairs, 4: quads -- twice the sample warps per SM for the latency-bound CUDA-core
// part; continuous only, the discontinuous chunks would not fit the TMEM plan).
template <class S, int NQ_ = 2>
struct TcPlan {
  static constexpr int NQ = NQ_;
  static constexpr int NSAMPLE = NQ * TS, NT = NSAMPLE + 32 + NHELPER, MMAW = NSAMPLE / 32;   // threads; MMA warp index
  static constexpr int HID = S::HID, ND = S::NODES_H, KH = HID * ND;       // 40, 5, 200
  static constexpr int HH = HID / NQ;                                      // hidden units owned by one thread of a group
  // FWD chunk: QI inputs of each half.  continuous (5 slots per input): 4 inputs, 40 K-columns per chunk;
  // discontinuous (6 slots): 2 inputs, 24 K-columns per chunk -- keeps the double-buffered A staging plus the
  // 240-column T accumulator inside the 512 TMEM columns
  static constexpr int QI = (S::CONT ? 8 : 4) / NQ, CH_IN = NQ * QI, NCH = HH / QI;
  static constexpr int HK = QI * ND, CH_K = NQ * HK;                       // K-columns per thread / per chunk
  static constexpr int KP = (KH + 31) / 32 * 32;                            // 224: W1 columns padded to atoms
  static constexpr int TN = (KH + 15) / 16 * 16;                            // 208: N of the GX product
  static constexpr int UNIT = 8, NUNIT = TS / UNIT;                         // GW: samples per smem unit (smem budget)
  // continuous: the layer-0 dL/dw tiles (dL/dy rows + basis records) get their own shared memory, so the owner
  // walk does not have to wait until the helpers are done with a1 / a2; discontinuous (larger W1): they alias a1 / a2
  static constexpr bool OWN_G0 = S::CONT;   // (basis records only; the dL/dy tile aliases the idle a1 buffer)
  // continuous: the a1 tile is DOUBLE-BUFFERED (the helpers may stage the Phi1 units of tile n while the sample
  // warps already work on tile n + 1 -- measured: 3.4 k of 31 k cycles per tile were spent waiting for them), the
  // layer-0 dL/dy tile of the owner walk aliases the idle a1 buffer, the basis records get 8 KB of their own.
  // discontinuous (W1 is 10 KB larger): same, but the basis records alias the a2 tile (dead by then).
  static constexpr bool A1_DOUBLE = true;
  static constexpr int NA1 = A1_DOUBLE ? 2 : 1;
  static constexpr int GPITCH = A1_DOUBLE ? 40 : 44;                        // layer-0 dL/dy tile [s][GPITCH]
  // TMEM columns
  static constexpr int D1_COL = 0, FA_COL = 64, G_COL = 64, T_COL = FA_COL + 4 * CH_K, GW_COL = T_COL + TN;
  // shared memory (bytes)
  static constexpr uint32_t W1_RS = KP / 32 * 512, W1_BYTES = HID / 4 * W1_RS;          // 3584, 35840
  static constexpr uint32_t GT_RS = TS / 32 * 512, GT_BYTES = HID / 4 * GT_RS;          // 2048, 20480
  static constexpr uint32_t U_RS = 256 / 32 * 512, U_BYTES = UNIT / 4 * U_RS;           // 4096, 16384
  static constexpr uint32_t A_BYTES = TS * APITCH * 4;                                  // 20992
  static constexpr uint32_t W2_BYTES = S::WO * 4;
  static constexpr uint32_t X_BYTES = NQ * TS * 4 * 4;                                  // group exchange slots
  static constexpr uint32_t OFF_W1R = 0, OFF_W1L = OFF_W1R + W1_BYTES, OFF_GTR = OFF_W1L + W1_BYTES,
                            OFF_GTL = OFF_GTR + GT_BYTES, OFF_U = OFF_GTL + GT_BYTES,   // [2][raw|lo]
                            OFF_A1 = OFF_U + 4 * U_BYTES, OFF_A2 = OFF_A1 + NA1 * A_BYTES, OFF_W2 = OFF_A2 + A_BYTES,
                            OFF_ACC2 = OFF_W2 + W2_BYTES, OFF_X = OFF_ACC2 + W2_BYTES, OFF_G0 = OFF_X + X_BYTES;
  static constexpr uint32_t G0_BYTES = TS * GPITCH * 4;                                 // [s][GPITCH] dL/d(layer-0 output)
  static constexpr uint32_t REC_BYTES = S::IN * 16 * TS;                                // [IN][TS] basis records
  static constexpr uint32_t OFF_BAR = OFF_G0 + (OWN_G0 ? REC_BYTES : 0), SMEM_BYTES = OFF_BAR + 256;
  static_assert(S::N == 3 && S::SEGH == 2 && S::SEGO == 2 && S::NH == 1, "tensor path: C2 / C3 family");
  static_assert(HID == 40 && HID % NQ == 0 && QI >= 1 && HH % QI == 0 && HK % 2 == 0 && HH % 2 == 0 && CH_K % 8 == 0 && KH <= 256,
                "tensor path shape");
  static_assert(NQ == 2 || NQ == 4, "pairs or quads");
  static_assert(G_COL + 2 * HID <= T_COL && GW_COL + 2 * HID <= 512 && UNIT % 8 == 0 && (UNIT / 2) % 4 == 0, "TMEM / unit plan");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory plan");
  static_assert(G0_BYTES <= A_BYTES && REC_BYTES <= A_BYTES, "layer-0 dL/dw tiles alias the idle a1 buffer / a2");
  static_assert(GPITCH % 4 == 0 && GPITCH >= HID, "dL/dy rows are read as float4");
  // K-column of node 0 of hidden-layer input i: chunk c = (i % 20) / 4 holds [half 0: 4 inputs | half 1: 4 inputs]
  __host__ __device__ static constexpr int kbase(int i) { return ((i % HH) / QI) * CH_K + (i / HH) * HK + (i % QI) * ND; }
};

enum { B_FFULL0 = 0, B_FFULL1, B_FEMPTY0, B_FEMPTY1, B_D1FULL, B_D1EMPTY, B_GFULL, B_TFULL, B_GWDONE,
       B_UFULL0, B_UFULL1, B_UEMPTY0, B_UEMPTY1, B_A1FULL, B_A1FREE0, B_A1FREE1, B_A2FULL, B_A2FREE, B_OWNDONE, B_COUNT };

// y[O] += sum_j l[j] * W[j * PITCH + 0..O)
template <int N, int O, int PITCH, bool GLOBAL>
__device__ __forceinline__ void contract_fwd_p(const float* __restrict__ wrow, const float* l, float* __restrict__ acc) {
  static_assert(O % 2 == 0 && PITCH % 4 == 0, "vector width");
  if (O % 4 == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(wrow + j * PITCH);
#pragma unroll
      for (int q = 0; q < O / 4; ++q) {
        const float4 w = GLOBAL ? __ldg(wp + q) : wp[q];
        acc[4 * q + 0] = fmaf(l[j], w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(l[j], w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(l[j], w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(l[j], w.w, acc[4 * q + 3]);
      }
    }
  } else {   // the row slice starts on an 8-byte boundary only (quads: 10 outputs per thread)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float2* wp = reinterpret_cast<const float2*>(wrow + j * PITCH);
#pragma unroll
      for (int q = 0; q < O / 2; ++q) {
        const float2 w = GLOBAL ? __ldg(wp + q) : wp[q];
        acc[2 * q + 0] = fmaf(l[j], w.x, acc[2 * q + 0]);
        acc[2 * q + 1] = fmaf(l[j], w.y, acc[2 * q + 1]);
      }
    }
  }
}

// Densified basis row of one input: the N = 3 Lagrange values of segment id (0 / 1) scattered into the ND node
// slots (continuous: the two segments share node 2; discontinuous: slots 0-2 / 3-5), zeros elsewhere.
template <class S, int ND>
__device__ __forceinline__ void densify(int id, const float* __restrict__ l, float* __restrict__ v) {
#pragma unroll
  for (int q = 0; q < ND; ++q) {
    const float lower = q < 3 ? l[q < 3 ? q : 0] : 0.f;
    const float upper = (q >= S::STRIDE && q - S::STRIDE < 3) ? l[(q >= S::STRIDE && q - S::STRIDE < 3) ? q - S::STRIDE : 0] : 0.f;
    v[q] = id ? upper : lower;
  }
}
// the 3 entries of a densified row t[ND] that belong to segment id
template <class S>
__device__ __forceinline__ void window3(int id, const float* __restrict__ t, float& t0, float& t1, float& t2) {
  t0 = id ? t[S::STRIDE + 0] : t[0];
  t1 = id ? t[S::STRIDE + 1] : t[1];
  t2 = id ? t[S::STRIDE + 2] : t[2];
}
// K consecutive columns (K even; any column offset is legal for .x2 / .x4: csrc/probe/tmem_align_probe.cu)
template <int K>
__device__ __forceinline__ void tmem_stK(uint32_t addr, const float* v) {
  static_assert(K % 2 == 0, "column groups of 4 and a tail of 2");
#pragma unroll
  for (int q = 0; q < K / 4; ++q) tmem_st4(addr + 4 * q, v + 4 * q);
  if (K % 4 != 0)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 ::"r"(addr + K - 2), "r"(__float_as_uint(v[K - 2])), "r"(__float_as_uint(v[K - 1])) : "memory");
}
template <int K>
__device__ __forceinline__ void tmem_ldK(uint32_t addr, float* v) {
  static_assert(K % 2 == 0, "column groups of 4 and a tail of 2");
#pragma unroll
  for (int q = 0; q < K / 4; ++q) tmem_ld4(addr + 4 * q, v + 4 * q);
  if (K % 4 != 0) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr + K - 2));
    v[K - 2] = __uint_as_float(r0);
    v[K - 1] = __uint_as_float(r1);
  }
}

// the NQ warps (w, w + 4, ...) that share 32 samples
template <int NQ>
__device__ __forceinline__ void bar_group(int warp) {
  if (NQ == 2) {
    bar_pair(warp);
  } else {
    switch (warp & 3) {
      case 0: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
      case 1: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
      case 2: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
      default: asm volatile("bar.sync 5, 128;" ::: "memory"); break;
    }
  }
}
template <int NSAMPLE>
__device__ __forceinline__ void bar_all_samples() {
  if (NSAMPLE == 256) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 1, 512;" ::: "memory");
}
// sum of one float over the group's threads (exchange slots xsamp[q * TS * 4], q = 0 .. NQ-1, own slot = hf)
template <int NQ>
__device__ __forceinline__ float group_sum(float v, float* xsamp, int hf, int warp) {
  xsamp[hf * TS * 4] = v;
  bar_group<NQ>(warp);
  float t = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) t += xsamp[q * TS * 4];     // same order in every thread of the group
  bar_group<NQ>(warp);
  return t;
}
// MaxAbsNormalization of a 40-wide row split over a thread group (W values each); first index wins ties.
template <int W, int NQ>
__device__ __forceinline__ void maxabs_group(float* v, int hf, float eps, float* xsamp, int warp, float& inv_d, int& kidx) {
  // arg-max of |v| as a balanced tree (depth log2 W instead of a W-long dependent chain); the left operand wins
  // ties, so the first index wins exactly like the sequential scan
  float ma[W], mv[W];
  int mk[W];
#pragma unroll
  for (int i = 0; i < W; ++i) { ma[i] = fabsf(v[i]); mv[i] = v[i]; mk[i] = i; }
#pragma unroll
  for (int n = W; n > 1; n = (n + 1) / 2) {
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const bool right = ma[2 * i + 1] > ma[2 * i];
      ma[i] = right ? ma[2 * i + 1] : ma[2 * i];
      mv[i] = right ? mv[2 * i + 1] : mv[2 * i];
      mk[i] = right ? mk[2 * i + 1] : mk[2 * i];
    }
    if (n & 1) { ma[n / 2] = ma[n - 1]; mv[n / 2] = mv[n - 1]; mk[n / 2] = mk[n - 1]; }
  }
  const float m = ma[0], vk = mv[0];
  const int k = mk[0];
  int kk = hf * W + k;
  kk |= vk > 0.f ? 0 : (vk < 0.f ? 0x10000 : 0x20000);
  xsamp[hf * TS * 4] = m;
  xsamp[hf * TS * 4 + 1] = __int_as_float(kk);
  bar_group<NQ>(warp);
  float mg = -1.f;
  int kg = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {                           // ascending unit index, strict >: the first maximum wins
    const float mo = xsamp[q * TS * 4];
    const int ko = __float_as_int(xsamp[q * TS * 4 + 1]);
    if (mo > mg) { mg = mo; kg = ko; }
  }
  bar_group<NQ>(warp);
  kidx = kg;
  inv_d = __frcp_rn(__fadd_rn(mg, eps));                // the correctly rounded reciprocal, like 1 / d
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] *= inv_d;
}

// MaxAbsNormalization of a 40-wide row split over a thread pair (20 values each); first index wins ties.
template <int W>
__device__ __forceinline__ void maxabs_pair(float* v, int hf, float eps, float* xs, const float* xo, int warp,
                                            float& inv_d, int& kidx) {
  float m = -1.f, vk = 0.f;
  int k = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float a = fabsf(v[i]);
    if (a > m) { m = a; k = i; vk = v[i]; }
  }
  int kk = hf * W + k;
  kk |= vk > 0.f ? 0 : (vk < 0.f ? 0x10000 : 0x20000);
  xs[0] = m;
  xs[1] = __int_as_float(kk);
  bar_pair(warp);
  const float mo = xo[0];
  const int ko = __float_as_int(xo[1]);
  bar_pair(warp);
  const bool other = (mo > m) || (mo == m && hf == 1);
  m = other ? mo : m;
  kidx = other ? ko : kk;
  inv_d = __frcp_rn(__fadd_rn(m, eps));
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] *= inv_d;
}
__device__ __forceinline__ float kidx_sign(int kidx) { return (kidx & 0x10000) ? -1.f : ((kidx & 0x20000) ? 0.f : 1.f); }

template <class S, int NQ>
__global__ void __launch_bounds__((TcPlan<S, NQ>::NT), 1) fused_mlp_tc_kernel(const FusedParams p) {
  using P = TcPlan<S, NQ>;
  constexpr int N = S::N, IN = S::IN, HID = S::HID, OUT = S::OUT, HP = S::HP, OP = S::OP, ND = P::ND, HH = P::HH;
  constexpr int NSAMPLE_THREADS = P::NSAMPLE, NTHREADS = P::NT, MMAW = P::MMAW;   // (shadow the pair constants)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1R = smem + P::OFF_W1R;
  uint8_t* sW1L = smem + P::OFF_W1L;
  uint8_t* sGTR = smem + P::OFF_GTR;
  uint8_t* sGTL = smem + P::OFF_GTL;
  uint8_t* sU = smem + P::OFF_U;
  float* sA1 = reinterpret_cast<float*>(smem + P::OFF_A1);
  float* sA2 = reinterpret_cast<float*>(smem + P::OFF_A2);
  float* sW2 = reinterpret_cast<float*>(smem + P::OFF_W2);
  float* sGY = reinterpret_cast<float*>(smem + P::OFF_ACC2);   // [TS][4] dL/dy rows for the helpers' dL/dw2 pass
  float* sX = reinterpret_cast<float*>(smem + P::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (uniform datapath)

  // ---- prologue: TMEM, barriers, weights -------------------------------------------------
  if (warp == MMAW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[B_FFULL0], NSAMPLE_THREADS); mbar_init(&bars[B_FFULL1], NSAMPLE_THREADS);
    mbar_init(&bars[B_FEMPTY0], 1);  mbar_init(&bars[B_FEMPTY1], 1);
    mbar_init(&bars[B_D1FULL], 1);   mbar_init(&bars[B_D1EMPTY], NSAMPLE_THREADS);
    mbar_init(&bars[B_GFULL], NSAMPLE_THREADS);  mbar_init(&bars[B_TFULL], 1);   mbar_init(&bars[B_GWDONE], 1);
    mbar_init(&bars[B_UFULL0], NHELPER); mbar_init(&bars[B_UFULL1], NHELPER);
    mbar_init(&bars[B_UEMPTY0], 1);  mbar_init(&bars[B_UEMPTY1], 1);
    mbar_init(&bars[B_A1FULL], NSAMPLE_THREADS); mbar_init(&bars[B_A1FREE0], NHELPER); mbar_init(&bars[B_A1FREE1], NHELPER);
    mbar_init(&bars[B_A2FULL], NSAMPLE_THREADS); mbar_init(&bars[B_A2FREE], NHELPER);
    mbar_init(&bars[B_OWNDONE], NSAMPLE_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  {
    // hidden-layer weights: packed [i][node][HP] -> smem rows = o, cols = kbase(i) + node (raw and tf32 residual)
    const float* w1 = p.packed + p.off[1];
    for (uint32_t k = tid; k < (uint32_t)(P::W1_BYTES / 4); k += NTHREADS) {
      reinterpret_cast<float*>(sW1R)[k] = 0.f;
      reinterpret_cast<float*>(sW1L)[k] = 0.f;
    }
    for (uint32_t k = tid; k < (uint32_t)(4 * P::U_BYTES / 4); k += NTHREADS) reinterpret_cast<float*>(sU)[k] = 0.f;
    __syncthreads();
    for (int k = tid; k < P::KH * HID; k += NTHREADS) {
      const int o = k % HID, r = k / HID, i = r / ND, q = r - i * ND;
      const float v = __ldg(w1 + (size_t)r * HP + o);
      const uint32_t off = sw4_off(o, P::kbase(i) + q, P::W1_RS);
      HOIR_DBG_ASSERT(off + 4u <= P::W1_BYTES);
      *reinterpret_cast<float*>(sW1R + off) = v;
      *reinterpret_cast<float*>(sW1L + off) = tf32_lo(v);
    }
    const float* w2 = p.packed + p.off[2];
    for (int k = tid; k < S::WO; k += NTHREADS) sW2[k] = __ldg(w2 + k);
    static_assert(TS * 4 * 4 <= (int)P::W2_BYTES, "dL/dy tile fits the slot behind W2");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;
  const long long tiles = (p.B + TS - 1) / TS;

  if (warp == MMAW) {
    // =========================== MMA warp ===================================================
    const uint32_t leader = elect_one();
    constexpr uint32_t ID_FWD = make_idesc(128, HID, 0, 0);        // A: TMEM (K-major), B: W1 K-major
    constexpr uint32_t ID_GX = make_idesc(128, P::TN, 0, 1);       // A: TMEM, B: W1 MN-major
    constexpr uint32_t ID_GW = make_idesc(128, HID, 1, 0);         // A: Phi unit MN-major, B: G^T K-major
    const uint64_t dW1[2] = {make_desc(smem_u32(sW1R), P::W1_RS), make_desc(smem_u32(sW1L), P::W1_RS)};
    const uint64_t dGT[2] = {make_desc(smem_u32(sGTR), P::GT_RS), make_desc(smem_u32(sGTL), P::GT_RS)};
    // The tensor pipe executes in issue order, so the warp issues whichever product has its operands ready:
    // the dL/dw1 units of tile n - 1 (paced by the helper warps) are interleaved with the forward chunks of
    // tile n instead of blocking them (measured: 4 k of 31 k cycles per tile were spent waiting on FEMPTY).
    uint32_t q_chunk = 0, q_unit = 0, n_tile = 0, gw_pending = 0;
    auto issue_fwd = [&](int c, uint32_t buf) {
      fence_after();
      if (leader) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int j = 0; j < P::CH_K / 8; ++j) {
            const uint32_t ks = (uint32_t)(c * (P::CH_K / 8) + j);
            const uint32_t koff = (ks >> 2) * 512u + (ks & 3u) * 32u;
            mma_ts(tmem + P::D1_COL, tmem + P::FA_COL + buf * 2 * P::CH_K + (pass == 1 ? P::CH_K : 0) + j * 8,
                   dW1[pass == 2] + (uint64_t)(koff >> 4), ID_FWD, (c | pass | j) != 0);
          }
        }
        umma_commit(&bars[B_FEMPTY0 + buf]);
        if (c == P::NCH - 1) umma_commit(&bars[B_D1FULL]);
      }
      __syncwarp();
    };
    // one unit of dL/dw1 (gW1 += Phi1^T * G); q_unit counts units over the whole kernel
    auto issue_gw = [&]() {
      const uint32_t buf = q_unit & 1, u = q_unit % P::NUNIT;
      fence_after();
      if (leader) {
        const uint32_t ubase = smem_u32(sU + buf * 2 * P::U_BYTES);
        const uint64_t dU[2] = {make_desc(ubase, P::U_RS), make_desc(ubase + P::U_BYTES, P::U_RS)};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
            for (int j = 0; j < P::UNIT / 8; ++j) {
              const uint32_t ks = u * (P::UNIT / 8) + j;
              const uint32_t koff = (ks >> 2) * 512u + (ks & 3u) * 32u;
              mma_ss(tmem + P::GW_COL + half * HID,
                     dU[pass == 1] + (uint64_t)((j * 2 * P::U_RS + half * 4 * 512) >> 4),
                     dGT[pass == 2] + (uint64_t)(koff >> 4), ID_GW, (q_unit | pass | j) != 0);
            }
          }
        }
        umma_commit(&bars[B_UEMPTY0 + buf]);
        if (u == P::NUNIT - 1) umma_commit(&bars[B_GWDONE]);
      }
      __syncwarp();
      ++q_unit;
      --gw_pending;
    };
    auto gw_ready = [&]() { return gw_pending > 0 && mbar_test(&bars[B_UFULL0 + (q_unit & 1)], (q_unit >> 1) & 1); };
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      // ---- FWD: D1 = Phi1 * W1 chunk by chunk (priority), pending dL/dw1 units of the previous tile in between
      bool d1_free = n_tile == 0;
      uint32_t spins = 0;
      for (int c = 0; c < P::NCH;) {
        if (!d1_free) d1_free = mbar_test(&bars[B_D1EMPTY], (n_tile - 1) & 1);
        const uint32_t buf = q_chunk & 1, ph = (q_chunk >> 1) & 1;
        if (d1_free && mbar_test(&bars[B_FFULL0 + buf], ph)) {
          issue_fwd(c, buf);
          ++c;
          ++q_chunk;
          spins = 0;
        } else if (gw_ready()) {
          issue_gw();
          spins = 0;
        } else {
          __nanosleep(20);
          if (++spins > (1u << 26)) __trap();
        }
      }
      // ---- GX: T = G * W1^T.  The sample warps stage G only after GWDONE of the previous tile: keep issuing units.
      spins = 0;
      while (!mbar_test(&bars[B_GFULL], n_tile & 1)) {
        if (gw_ready()) {
          issue_gw();
          spins = 0;
        } else {
          __nanosleep(20);
          if (++spins > (1u << 26)) __trap();
        }
      }
      fence_after();
      if (leader) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
          for (int j = 0; j < HID / 8; ++j) {
            mma_ts(tmem + P::T_COL, tmem + P::G_COL + (pass == 1 ? HID : 0) + j * 8,
                   dW1[pass == 2] + (uint64_t)((j * 2 * P::W1_RS) >> 4), ID_GX, (pass | j) != 0);
          }
        }
        umma_commit(&bars[B_TFULL]);
      }
      __syncwarp();
      gw_pending += P::NUNIT;      // this tile's dL/dw1 units: issued as they arrive, here and in the next tile's loops
    }
    while (gw_pending > 0) {
      mbar_wait_backoff(&bars[B_UFULL0 + (q_unit & 1)], (q_unit >> 1) & 1, 32);
      issue_gw();
    }
  } else if (warp > MMAW) {
    // =========================== helper warps ===============================================
    // Stage Phi1 (densified basis of the hidden-layer inputs a1, raw + tf32 residual) for the dL/dw1 product in
    // units of 16 samples, double-buffered.  Thread = (input i, half of the unit): its K-columns are fixed, the
    // row swizzle is a compile-time constant per step.  The sample warps never touch the units.
    const int ht = tid - (NSAMPLE_THREADS + 32);
    // thread = (input i, sh): sh = 0, 1: samples 8 sh .. 8 sh + 7 of a unit; sh = 2: idle.
    // continuous shape (pairs): a warp takes the 32 inputs of chunks 0-3 -- their K-columns kbase(i) + k fall into 32
    // DIFFERENT banks, so a unit store is one wavefront (with inputs 0 .. 31 per warp, chunk 0 and chunk 4 collided:
    // 10.5 M excess wavefronts per launch, 12 % of all shared-memory wavefronts); the third warp takes chunk 4.
    int i = ht % HID, sh = ht / HID;
    if constexpr (S::CONT && NQ == 2 && HID == 40) {
      const int hw = ht >> 5, hl = ht & 31;
      if (hw < 2) {
        sh = hw;
        i = hl < 16 ? hl : hl + 4;                    // inputs 0 .. 15 and 20 .. 35: chunks 0 - 3 of both halves
      } else {
        sh = hl < 16 ? (hl >> 3) : 2;
        i = (hl & 7) < 4 ? 16 + (hl & 7) : 32 + (hl & 7);   // inputs 16 .. 19 and 36 .. 39: chunk 4
      }
    }
    const LagrangeTable& lag = p.lag;
    const int kb = P::kbase(i);
    uint32_t coff[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) coff[k] = ((uint32_t)((kb + k) & 31) << 2) | (((uint32_t)(kb + k) >> 5) << 16);
    // dL/dw2 (output layer): the same thread owns input i for one half of the tile's samples and keeps the
    // 5 nodes x 3 outputs sums in registers for the WHOLE kernel -- no warp shuffles, no per-tile reduction
    float acc2[ND][OUT];
#pragma unroll
    for (int k = 0; k < ND; ++k)
#pragma unroll
      for (int o = 0; o < OUT; ++o) acc2[k][o] = 0.f;
    uint32_t q_unit = 0, n_tile = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      mbar_wait_backoff(&bars[B_A1FULL], n_tile & 1, 128);
      const uint32_t a1buf = P::A1_DOUBLE ? (n_tile & 1) : 0;
      const float* sA1t = sA1 + a1buf * (P::A_BYTES / 4);
      for (int u = 0; u < P::NUNIT; ++u, ++q_unit) {
        if (u == 2) {   // two units are staged ahead; now the hidden activations a2 and dL/dy of this tile are due
          mbar_wait_backoff(&bars[B_A2FULL], n_tile & 1, 64);
          if (sh < 2) {
#pragma unroll 4
            for (int t = 0; t < TS / 2; ++t) {
              const int sl = sh * (TS / 2) + t;
              const float x = sA2[sl * APITCH + i];
              const float4 gy4 = *reinterpret_cast<const float4*>(sGY + sl * 4);
              int id;
              float xi, l[N];
              locate2(x, id, xi);
              hoir::lagrange<N>(lag.X, lag.D, xi, l);
              float c5[ND];
            densify<S, ND>(id, l, c5);
              const float gyv[3] = {gy4.x, gy4.y, gy4.z};
#pragma unroll
              for (int k = 0; k < ND; ++k)
#pragma unroll
                for (int o = 0; o < OUT; ++o) acc2[k][o] = fmaf(c5[k], gyv[o], acc2[k][o]);
            }
          }
          mbar_arrive(&bars[B_A2FREE]);
        }
        const uint32_t buf = q_unit & 1, ph = (q_unit >> 1) & 1;
        mbar_wait_backoff(&bars[B_UEMPTY0 + buf], ph ^ 1, 128);
        if (sh < 2) {
          uint8_t* ur = sU + buf * 2 * P::U_BYTES;
          uint8_t* ul = ur + P::U_BYTES;
#pragma unroll
          for (int t = 0; t < P::UNIT / 2; ++t) {
            const int sl = sh * (P::UNIT / 2) + t;   // (sl & 3) == (t & 3): compile-time swizzle
            const float x = sA1t[(u * P::UNIT + sl) * APITCH + i];
            int id;
            float xi, l[N];
            locate2(x, id, xi);
            hoir::lagrange<N>(lag.X, lag.D, xi, l);
            float c5[ND];
            densify<S, ND>(id, l, c5);
            const uint32_t rowoff = (uint32_t)(t & 3) * 128u + (uint32_t)(sl >> 2) * P::U_RS;
#pragma unroll
            for (int k = 0; k < ND; ++k) {
              const uint32_t off = rowoff + ((coff[k] & 0xffffu) ^ ((uint32_t)(t & 3) << 5)) + (coff[k] >> 16) * 512u;
              HOIR_DBG_ASSERT(off + 4u <= P::U_BYTES);
              *reinterpret_cast<float*>(ur + off) = c5[k];
              *reinterpret_cast<float*>(ul + off) = tf32_lo(c5[k]);
            }
          }
        }
        fence_async_smem();
        mbar_arrive(&bars[B_UFULL0 + buf]);
      }
      mbar_arrive(&bars[B_A1FREE0 + a1buf]);          // done reading this tile's a1 rows
    }
    if (sh < 2) {
#pragma unroll
      for (int k = 0; k < ND; ++k)
#pragma unroll
        for (int o = 0; o < OUT; ++o)
          if (acc2[k][o] != 0.f) hoir::red_add(p.gw[2] + ((size_t)o * HID + i) * S::NODES_O + k, acc2[k][o]);
    }
  } else {
    // =========================== sample warps ===============================================
    const int hf = warp >> 2;                         // which 20 hidden units this thread owns
    const int s = tid & (TS - 1);                     // sample slot == TMEM lane
    const SegConst sc0 = {p.seg0, S::STRIDE, (float)p.seg0, 1.f / (float)p.seg0};
    const float* w0 = p.packed + p.off[0] + hf * HH;
    const LagrangeTable& lag = p.lag;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* a1 = sA1 + s * APITCH;                     // re-pointed per tile when a1 is double-buffered
    float* a2 = sA2 + s * APITCH;
    float* xsamp = sX + s * 4;                        // exchange slots of this sample: [q][TS][4]
    float* xs = xsamp + hf * TS * 4;
    float loss_local = 0.f;
    uint32_t q_chunk = 0, q_unit = 0, n_tile = 0;
    // layer 0 (FFMA, weights through L1) of one tile: this thread's 20 outputs, MaxAbs over the pair
    // (also fetches the tile's targets / external dL/dy: their DRAM latency hides behind the previous tile)
    // p.perm != nullptr: sample slot b reads SOURCE index perm[b] -- a row of x / target / gy_ext (explicit batches sorted
    // by layer-0 cell, hoir_mlp_batch_sort) or an absolute pixel of the mesh (shuffled pixel indices).  The index of the
    // tile after the one being loaded is prefetched one layer0 call ahead, so only one dependent load is exposed.
    int permN = 0;
    if (p.perm != nullptr && (long long)blockIdx.x * TS + s < p.B) permN = __ldg(p.perm + (long long)blockIdx.x * TS + s);
    auto layer0 = [&](long long tile_, float* h_, float4* rec_, float& inv_d_, int& kidx_, float* tg_) {
      const long long b_ = tile_ * TS + s;
      const long long src_ = p.perm != nullptr ? (long long)permN : b_;
      const long long pix_ = p.perm != nullptr ? (long long)permN : p.mesh.first_pixel + b_;
      if (p.perm != nullptr) {
        const long long bn_ = (tile_ + gridDim.x) * TS + s;
        if (bn_ < p.B) permN = __ldg(p.perm + bn_);
      }
      float x0[IN];
#pragma unroll
      for (int i = 0; i < IN; ++i) x0[i] = 0.f;
#pragma unroll
      for (int o = 0; o < OUT; ++o) tg_[o] = 0.f;
      if (b_ < p.B) {
        if (p.gy_ext != nullptr) {
#pragma unroll
          for (int o = 0; o < OUT; ++o) tg_[o] = __ldg(p.gy_ext + src_ * OUT + o);
        } else if (p.target_kind == HOIR_TARGET_U8_IMAGE) {
          const unsigned char* img = static_cast<const unsigned char*>(p.target);
#pragma unroll
          for (int o = 0; o < OUT; ++o) tg_[o] = (float)__ldg(img + pix_ * OUT + o);
        } else {
#pragma unroll
          for (int o = 0; o < OUT; ++o) tg_[o] = __ldg(static_cast<const float*>(p.target) + src_ * OUT + o);
        }
      }
      if (b_ < p.B) {
        if (p.x != nullptr) {
          if constexpr (IN == 4) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x) + src_);
            x0[0] = xv.x; x0[1] = xv.y; x0[2] = xv.z; x0[3] = xv.w;
          } else if constexpr (IN == 2) {
            const float2 xv = __ldg(reinterpret_cast<const float2*>(p.x) + src_);
            x0[0] = xv.x; x0[1] = xv.y;
          } else {
#pragma unroll
            for (int i = 0; i < IN; ++i) x0[i] = __ldg(p.x + src_ * IN + i);
          }
        } else {
          hoir::mesh_position(p.mesh, pix_, x0, IN / 2);
        }
      }
#pragma unroll
      for (int o = 0; o < HH; ++o) h_[o] = 0.f;
#pragma unroll
      for (int i = 0; i < IN; ++i) {
        int id;
        float xi, l[N];
        if (p.pow2_0) locate<true>(x0[i], sc0, id, xi);
        else locate<false>(x0[i], sc0, id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        const int base0 = id * S::STRIDE;
        // pairs: the half-major copy -- this thread's N rows are one contiguous run of N * HH floats
        if constexpr (NQ == 2 && HH % 4 == 0)
          contract_fwd_p<N, HH, (HH % 4 == 0 ? HH : HP), true>(p.packed + p.alt_off + ((size_t)(i * 2 + hf) * p.nodes0 + base0) * HH, l, h_);
        else
          contract_fwd_p<N, HH, HP, true>(w0 + ((size_t)i * p.nodes0 + base0) * HP, l, h_);
        rec_[i] = make_float4(__int_as_float(base0), l[0], l[1], l[2]);
      }
      inv_d_ = 1.f;
      kidx_ = 0;
      if (p.normalize) maxabs_group<HH, NQ>(h_, hf, p.norm_eps, xsamp, warp, inv_d_, kidx_);
    };
    float hN[HH];
    float4 recN[IN];
    float invN = 1.f;
    int kidxN = 0;
    float tgN[OUT];
    constexpr int PRE = P::NCH - 1 < 4 ? P::NCH - 1 : 4;   // forward chunks of the next tile staged ahead (see below)
    auto stage_chunk = [&](int c) {                   // densified basis of this thread's inputs of chunk c (from hN)
      const uint32_t buf = q_chunk & 1, ph = (q_chunk >> 1) & 1;
      float v[P::HK], lo[P::HK];
#pragma unroll
      for (int m = 0; m < P::QI; ++m) {
        const float x = hN[c * P::QI + m];
        int id;
        float xi, l[N];
        locate2(x, id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        densify<S, ND>(id, l, v + m * ND);
      }
#pragma unroll
      for (int j = 0; j < P::HK; ++j) lo[j] = tf32_lo(v[j]);
      mbar_wait(&bars[B_FEMPTY0 + buf], ph ^ 1);       // (after the arithmetic: the buffer is usually free by now)
      fence_after();
      const uint32_t col = tm_lane + P::FA_COL + buf * 2 * P::CH_K + hf * P::HK;
      tmem_stK<P::HK>(col, v);
      tmem_stK<P::HK>(col + P::CH_K, lo);
      tmem_wait_st();
      fence_before();
      mbar_arrive(&bars[B_FFULL0 + buf]);
      ++q_chunk;
    };
    if ((long long)blockIdx.x < tiles) layer0(blockIdx.x, hN, recN, invN, kidxN, tgN);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      const long long b = tile * TS + s;
      const bool live = b < p.B;
      float h[HH];
      float4 rec0[IN];
      float inv_d1 = 1.f;
      int kidx1 = 0;
      const float inv_d0 = invN;
      const int kidx0 = kidxN;
      float tg0[OUT];
#pragma unroll
      for (int o = 0; o < OUT; ++o) tg0[o] = tgN[o];
#pragma unroll
      for (int i = 0; i < IN; ++i) rec0[i] = recN[i];
      // ---- hidden layer forward on the tensor cores: stage this thread's inputs of every chunk (straight from
      //      the layer-0 registers) into TMEM.  The first PRE chunks of every tile but the first were already staged
      //      during the previous tile (right after its T arrived): the forward pipeline is 2 buffers deep, so its
      //      latency (5 x (15 MMAs + hand-overs) ~ 4 k cycles) would otherwise sit on this path.
#pragma unroll
      for (int c = 0; c < P::NCH; ++c) {
        if (n_tile == 0) stage_chunk(c);              // later tiles: all chunks were staged during the previous tile
      }
      // ---- a1 tile for the helpers (dL/dw1 units) and the T epilogue.  Its buffer is free once the helpers have
      //      staged its previous occupant and every owner of the previous tile's layer-0 walk is done (the walk
      //      tiles alias the a1 / a2 buffers) -- both long past by now, so neither wait stalls.
      {
        const uint32_t a1buf = P::A1_DOUBLE ? (n_tile & 1) : 0, use = P::A1_DOUBLE ? (n_tile >> 1) : n_tile;
        a1 = sA1 + a1buf * (P::A_BYTES / 4) + s * APITCH;
        if (use > 0) mbar_wait(&bars[B_A1FREE0 + a1buf], (use - 1) & 1);
        if (n_tile > 0) mbar_wait(&bars[B_OWNDONE], (n_tile - 1) & 1);
      }
#pragma unroll
      for (int j = 0; j < HH; ++j) a1[hf * HH + j] = hN[j];
      mbar_arrive(&bars[B_A1FULL]);                                   // (release) helpers may stage Phi1 units
      // ---- D1 epilogue: hidden pre-activation (own 20 columns) -> MaxAbs -> a2
      mbar_wait(&bars[B_D1FULL], n_tile & 1);
      fence_after();
      tmem_ldK<HH>(tm_lane + P::D1_COL + hf * HH, h);
      tmem_wait_ld();
      fence_before();
      mbar_arrive(&bars[B_D1EMPTY]);
      if (p.normalize) maxabs_group<HH, NQ>(h, hf, p.norm_eps, xsamp, warp, inv_d1, kidx1);
      if (n_tile > 0) mbar_wait(&bars[B_A2FREE], (n_tile - 1) & 1);   // helpers are done with the previous a2 / dL/dy
#pragma unroll
      for (int j = 0; j < HH; ++j) a2[hf * HH + j] = h[j];
      // ---- output layer forward (FFMA) over own 20 inputs, pair sum, loss, dL/dy
      float yo[4] = {0.f, 0.f, 0.f, 0.f}, gy[OUT];
#pragma unroll 4
      for (int j = 0; j < HH; ++j) {
        const int i = hf * HH + j;
        int id;
        float xi, l[N];
        locate2(a2[i], id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        contract_fwd<N, OUT, OP, false>(sW2 + ((size_t)i * S::NODES_O + id * S::STRIDE) * OP, l, yo);
      }
      *reinterpret_cast<float4*>(xs) = make_float4(yo[0], yo[1], yo[2], 0.f);
      bar_group<NQ>(warp);
      {
        float ys[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < NQ; ++q) {                 // same order in every thread of the group
          const float4 o4 = *reinterpret_cast<const float4*>(xsamp + q * TS * 4);
          ys[0] += o4.x; ys[1] += o4.y; ys[2] += o4.z;
        }
        yo[0] = ys[0]; yo[1] = ys[1]; yo[2] = ys[2];
      }
      bar_group<NQ>(warp);
#pragma unroll
      for (int o = 0; o < OUT; ++o) gy[o] = 0.f;
      if (live) {
        if (p.gy_ext != nullptr) {
#pragma unroll
          for (int o = 0; o < OUT; ++o) gy[o] = tg0[o];
        } else {
#pragma unroll
          for (int o = 0; o < OUT; ++o) {
            const float t = p.target_kind == HOIR_TARGET_U8_IMAGE
                                ? __fsub_rn(__fdiv_rn(__fmul_rn(tg0[o], 2.f), 255.f), 1.f) : tg0[o];
            const float diff = yo[o] - t;
            if (hf == 0) loss_local = fmaf(diff, diff, loss_local);
            gy[o] = 2.f * diff * p.loss_scale;
          }
        }
      }
      if (hf == 0) *reinterpret_cast<float4*>(sGY + s * 4) = make_float4(gy[0], gy[1], gy[2], 0.f);
      mbar_arrive(&bars[B_A2FULL]);                   // (release) a2 row + dL/dy: the helpers accumulate dL/dw2
      // ---- output layer backward (FFMA) over own 20 inputs: dL/da2
      float g[HH];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < HH; ++j) {
        const int i = hf * HH + j;
        const float x = a2[i];
        int id;
        float xi, l[N], d[N], t[N];
        locate2(x, id, xi);
        hoir::lagrange_deriv<N>(lag.X, lag.D, xi, l, d);
        contract_bwd<N, OUT, OP>(sW2 + ((size_t)i * S::NODES_O + id * S::STRIDE) * OP, gy, t);
        float gx = 0.f;
#pragma unroll
        for (int q = 0; q < N; ++q) gx = fmaf(d[q], t[q], gx);
        gx *= 2.f;                                   // d xi / d x = segments
        dot = fmaf(gx, x, dot);
        g[j] = gx;                                   // dL/da2 (a2 keeps the activations for the dL/dw2 pass below)
      }
      if (p.normalize) {
        dot = group_sum<NQ>(dot, xsamp, hf, warp);
        const int k = (kidx1 & 0xffff) - hf * HH;
        const float sgn = kidx_sign(kidx1);
#pragma unroll
        for (int j = 0; j < HH; ++j) {
          float gi = g[j];
          if (j == k) gi -= sgn * dot;
          g[j] = gi * inv_d1;
        }
      }
      // ---- stage G (own 20 units): TMEM (A operand of GX) and smem G^T (B operand of GW)
      if (n_tile > 0) mbar_wait(&bars[B_GWDONE], (n_tile - 1) & 1);   // previous tile's GW done with G^T
      fence_after();
      {
        float lo[HH];
#pragma unroll
        for (int j = 0; j < HH; ++j) lo[j] = tf32_lo(g[j]);
        const uint32_t col = tm_lane + P::G_COL + hf * HH;
        tmem_stK<HH>(col, g);
        tmem_stK<HH>(col + HID, lo);
#pragma unroll
        for (int j = 0; j < HH; ++j) {
          const uint32_t off = sw4_off(hf * HH + j, s, P::GT_RS);
          HOIR_DBG_ASSERT(off + 4u <= P::GT_BYTES);
          *reinterpret_cast<float*>(sGTR + off) = g[j];
          *reinterpret_cast<float*>(sGTL + off) = lo[j];
        }
      }
      tmem_wait_st();
      fence_async_smem();
      fence_before();
      mbar_arrive(&bars[B_GFULL]);
      // ---- while the tensor pipe computes T = G * W1^T (the longest product): layer 0 of the NEXT tile
      if (tile + gridDim.x < tiles) layer0(tile + gridDim.x, hN, recN, invN, kidxN, tgN);
      // ---- T epilogue (own 4 inputs per chunk): dL/da1 = slope * sum_j l'_j * T[kbase(i) + window + j]
      dot = 0.f;
      mbar_wait(&bars[B_TFULL], n_tile & 1);
      fence_after();
      // T exists, so the G staging columns (they alias the forward staging buffers) are free again: stage the
      // first chunks of the next tile now (hN already holds its layer-0 output)
      const bool has_next = tile + gridDim.x < tiles;
      if (has_next) {
#pragma unroll
        for (int c = 0; c < 2 && c < PRE; ++c) stage_chunk(c);
      }
#pragma unroll
      for (int c = 0; c < P::NCH; ++c) {
        float t[P::HK];
        tmem_ldK<P::HK>(tm_lane + P::T_COL + c * P::CH_K + hf * P::HK, t);
        tmem_wait_ld();
#pragma unroll
        for (int m = 0; m < P::QI; ++m) {
          const float x = a1[hf * HH + c * P::QI + m];
          int id;
          float xi, l[N], d[N];
          locate2(x, id, xi);
          hoir::lagrange_deriv<N>(lag.X, lag.D, xi, l, d);
          float t0, t1, t2;
          window3<S>(id, t + m * ND, t0, t1, t2);
          const float gx = fmaf(d[2], t2, fmaf(d[1], t1, d[0] * t0)) * 2.f;
          dot = fmaf(gx, x, dot);
          g[c * P::QI + m] = gx;
        }
      }
      fence_before();
      if (has_next && PRE > 2) stage_chunk(2);        // its buffer (chunk 0 of the next tile) was consumed during the T epilogue
      if (p.normalize) {
        dot = group_sum<NQ>(dot, xsamp, hf, warp);
        const int k = (kidx0 & 0xffff) - hf * HH;
        const float sgn = kidx_sign(kidx0);
#pragma unroll
        for (int j = 0; j < HH; ++j) {
          float gi = g[j];
          if (j == k) gi -= sgn * dot;
          g[j] = gi * inv_d0;
        }
      }
      if (has_next && PRE > 3) stage_chunk(3);
      // ---- layer 0 dL/dw.  Incoherent batches (two-pass mode): stream dL/dy rows, basis records and segment
      //      ids to the workspace; gw0_owner_kernel finishes the job.
      if (p.ws_g0 != nullptr) {
        // the tile's dL/dy rows go through shared memory (the idle a1 buffer) so that the 20 KB leave the SM as
        // fully coalesced 16-byte stores instead of 160-byte-strided 8-byte ones
        const uint32_t ob = (n_tile + 1) & 1, use = (n_tile + 1) >> 1;
        if (use > 0) mbar_wait(&bars[B_A1FREE0 + ob], (use - 1) & 1);
        float* sG0 = sA1 + ob * (P::A_BYTES / 4);
        static_assert(P::GPITCH == HID, "the staged tile is copied as one contiguous block");
#pragma unroll
        for (int q = 0; q < HH / 2; ++q)
          *reinterpret_cast<float2*>(sG0 + s * P::GPITCH + hf * HH + 2 * q) = make_float2(g[2 * q], g[2 * q + 1]);
        if (live && hf == 0) {
#pragma unroll
          for (int i = 0; i < IN; ++i) {
            p.ws_rec[(size_t)i * p.B + b] = rec0[i];
            p.ws_seg[(size_t)i * p.B + b] = (unsigned short)(__float_as_int(rec0[i].x) / S::STRIDE);
          }
        }
        bar_all_samples<NSAMPLE_THREADS>();
        {
          const long long b0 = tile * TS;
          const int nrow = (int)((p.B - b0) < TS ? (p.B - b0) : TS);
          float4* dst = reinterpret_cast<float4*>(p.ws_g0 + b0 * HID);
          const float4* src = reinterpret_cast<const float4*>(sG0);
          for (int k = tid; k < nrow * (HID / 4); k += NSAMPLE_THREADS) dst[k] = src[k];
        }
        mbar_arrive(&bars[B_OWNDONE]);
        if (has_next) {
#pragma unroll
          for (int c = PRE; c < P::NCH; ++c) stage_chunk(c);
        }
        continue;
      }
      // ---- otherwise: owners (input, 4 outputs, sample group) walk the tile, run-merged global atomics
      {
        // the dL/dy tile aliases the a1 buffer of tile n + 1: free once the helpers are done with tile n - 1
        const uint32_t ob = (n_tile + 1) & 1, use = (n_tile + 1) >> 1;
        if (use > 0) mbar_wait(&bars[B_A1FREE0 + ob], (use - 1) & 1);
        // without room for their own 8 KB the basis records alias the a2 tile: the helpers are done with it
        // (A2FREE) and every sample thread is past its last a2 read (all 256 arrived on GFULL before T existed)
        if (!P::OWN_G0) mbar_wait(&bars[B_A2FREE], n_tile & 1);
      }
      {
        float* sG0 = sA1 + ((n_tile + 1) & 1) * (P::A_BYTES / 4);                          // [s][GPITCH] dL/d(layer-0 output)
        float4* sRec = P::OWN_G0 ? reinterpret_cast<float4*>(smem + P::OFF_G0) : reinterpret_cast<float4*>(sA2);   // [IN][TS]
#pragma unroll
        for (int q = 0; q < HH / 2; ++q)
          *reinterpret_cast<float2*>(sG0 + s * P::GPITCH + hf * HH + 2 * q) = make_float2(g[2 * q], g[2 * q + 1]);
        if (hf == 0) {
#pragma unroll
          for (int i = 0; i < IN; ++i) sRec[(size_t)i * TS + s] = rec0[i];
        }
        bar_all_samples<NSAMPLE_THREADS>();
        constexpr int NOQ = HID / 4, NGRP = NSAMPLE_THREADS / (IN * NOQ), SPG = (TS + NGRP - 1) / NGRP;
        if (tid < NGRP * IN * NOQ) {
          const int grp = tid / (IN * NOQ), pr = tid - grp * (IN * NOQ), i = pr / NOQ, oq = pr - i * NOQ;
          const int s0 = grp * SPG, s1 = (s0 + SPG < TS) ? s0 + SPG : TS;
          float* dst = p.gw[0] + ((size_t)(4 * oq) * IN + i) * p.nodes0;
          const size_t ostride = (size_t)IN * p.nodes0;
          const float4* rp = sRec + (size_t)i * TS;
          int cur = -1;
          float acc[4][N];
#pragma unroll
          for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int q = 0; q < N; ++q) acc[o][q] = 0.f;
          for (int ss = s0; ss < s1; ++ss) {
            const float4 r = rp[ss];
            const float4 gv = *reinterpret_cast<const float4*>(sG0 + ss * P::GPITCH + 4 * oq);
            const int base = __float_as_int(r.x);
            if (base != cur) {
              if (cur >= 0) {
#pragma unroll
                for (int o = 0; o < 4; ++o)
#pragma unroll
                  for (int q = 0; q < N; ++q)
                    if (acc[o][q] != 0.f) hoir::red_add(dst + o * ostride + cur + q, acc[o][q]);
              }
              cur = base;
#pragma unroll
              for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int q = 0; q < N; ++q) acc[o][q] = 0.f;
            }
            const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              acc[o][0] = fmaf(r.y, gq[o], acc[o][0]);
              acc[o][1] = fmaf(r.z, gq[o], acc[o][1]);
              acc[o][2] = fmaf(r.w, gq[o], acc[o][2]);
            }
          }
          if (cur >= 0) {
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
              for (int q = 0; q < N; ++q)
                if (acc[o][q] != 0.f) hoir::red_add(dst + o * ostride + cur + q, acc[o][q]);
          }
        }
      }
      mbar_arrive(&bars[B_OWNDONE]);                  // (release) this thread is done with the walk tiles; the next
                                                      // tile waits for all 256 before it rewrites a1 / a2
      if (has_next) {                                 // the remaining forward chunks of the next tile
#pragma unroll
        for (int c = PRE; c < P::NCH; ++c) stage_chunk(c);
      }
    }
    // ---- flush: loss, dL/dw2 (smem), dL/dw1 (TMEM accumulators)
    if (p.gy_ext == nullptr && hf == 0) {
      const float sum = hoir::warp_sum(loss_local) * p.loss_scale;
      if (lane == 0 && sum != 0.f) hoir::red_add(p.loss_sum, sum);
    }
    if (n_tile > 0) {
      mbar_wait(&bars[B_GWDONE], (n_tile - 1) & 1);
      fence_after();
      // accumulator row R = hf * 128 + s holds K-column R: chunk c = R / 40, half (R % 40) / 20, input 4c + ..., node R % 5
      // (quads: threads q and q + 2 share accumulator half q & 1 and take 20 of its 40 columns each)
      constexpr int NCP = NQ / 2, WCOLS = HID / NCP;
      const int half = hf % 2, cpart = hf / 2;
      const int R = half * 128 + s;
      float v[WCOLS];
      tmem_ldK<WCOLS>(tm_lane + P::GW_COL + half * HID + cpart * WCOLS, v);
      tmem_wait_ld();
      if (R < P::KH) {
        const int c = R / P::CH_K, rem = R - c * P::CH_K, hfi = rem / P::HK, r2 = rem - hfi * P::HK;
        const int i = hfi * HH + c * P::QI + r2 / ND, q = r2 % ND;
#pragma unroll
        for (int o = 0; o < WCOLS; ++o)
          if (v[o] != 0.f) hoir::red_add(p.gw[1] + ((size_t)(cpart * WCOLS + o) * HID + i) * ND + q, v[o]);
      }
      fence_before();
    }
  }
  __syncthreads();
  if (warp == MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  hoir::opt_tail_in_kernel(p.tail);
}

// ---------------------------------------------------------------------------------------------------------
// Forward-only tensor-core kernel (rendering / inference: hoir_mlp_forward): the forward half of the kernel
// above.  Same thread-pair organisation and FWD product (Phi1 chunks in TMEM x W1 in smem, tf32 x 3 passes);
// activations never leave registers, the only shared memory is W1 (raw + residual), W2 and the pair-exchange
// slots (72 KB) and only 256 TMEM columns are allocated, so TWO CTAs share an SM and one CTA's basis / output
// layer work fills the other's waits on the tensor pipe.
template <class S>
struct TcFwdPlan {
  using P = TcPlan<S>;
  static constexpr uint32_t OFF_W1R = 0, OFF_W1L = OFF_W1R + P::W1_BYTES, OFF_W2 = OFF_W1L + P::W1_BYTES,
                            OFF_X = OFF_W2 + P::W2_BYTES, OFF_BAR = OFF_X + P::X_BYTES, SMEM_BYTES = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 256;
  static_assert(P::FA_COL + 4 * P::CH_K <= TMEM_COLS, "TMEM plan (forward)");
};

template <class S>
__global__ void __launch_bounds__(NTHREADS_FWD, 2) fused_mlp_tc_fwd_kernel(const FusedParams p) {
  using P = TcPlan<S>;
  using F = TcFwdPlan<S>;
  constexpr int N = S::N, IN = S::IN, HID = S::HID, OUT = S::OUT, HP = S::HP, OP = S::OP, ND = P::ND, HH = P::HH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1R = smem + F::OFF_W1R;
  uint8_t* sW1L = smem + F::OFF_W1L;
  float* sW2 = reinterpret_cast<float*>(smem + F::OFF_W2);
  float* sX = reinterpret_cast<float*>(smem + F::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F::OFF_BAR);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(F::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bars[B_FFULL0], NSAMPLE_THREADS); mbar_init(&bars[B_FFULL1], NSAMPLE_THREADS);
    mbar_init(&bars[B_FEMPTY0], 1);  mbar_init(&bars[B_FEMPTY1], 1);
    mbar_init(&bars[B_D1FULL], 1);   mbar_init(&bars[B_D1EMPTY], NSAMPLE_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  {
    const float* w1 = p.packed + p.off[1];
    for (uint32_t k = tid; k < (uint32_t)(P::W1_BYTES / 4); k += NTHREADS_FWD) {
      reinterpret_cast<float*>(sW1R)[k] = 0.f;
      reinterpret_cast<float*>(sW1L)[k] = 0.f;
    }
    __syncthreads();
    for (int k = tid; k < P::KH * HID; k += NTHREADS_FWD) {
      const int o = k % HID, r = k / HID, i = r / ND, q = r - i * ND;
      const float v = __ldg(w1 + (size_t)r * HP + o);
      const uint32_t off = sw4_off(o, P::kbase(i) + q, P::W1_RS);
      *reinterpret_cast<float*>(sW1R + off) = v;
      *reinterpret_cast<float*>(sW1L + off) = tf32_lo(v);
    }
    const float* w2 = p.packed + p.off[2];
    for (int k = tid; k < S::WO; k += NTHREADS_FWD) sW2[k] = __ldg(w2 + k);
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;
  const long long tiles = (p.B + TS - 1) / TS;

  if (warp == 8) {
    const uint32_t leader = elect_one();
    constexpr uint32_t ID_FWD = make_idesc(128, HID, 0, 0);
    const uint64_t dW1[2] = {make_desc(smem_u32(sW1R), P::W1_RS), make_desc(smem_u32(sW1L), P::W1_RS)};
    uint32_t q_chunk = 0, n_tile = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      if (n_tile > 0) mbar_wait(&bars[B_D1EMPTY], (n_tile - 1) & 1);
      for (int c = 0; c < P::NCH; ++c, ++q_chunk) {
        const uint32_t buf = q_chunk & 1, ph = (q_chunk >> 1) & 1;
        mbar_wait(&bars[B_FFULL0 + buf], ph);
        fence_after();
        if (leader) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
            for (int j = 0; j < P::CH_K / 8; ++j) {
              const uint32_t ks = (uint32_t)(c * (P::CH_K / 8) + j);
              const uint32_t koff = (ks >> 2) * 512u + (ks & 3u) * 32u;
              mma_ts(tmem + P::D1_COL, tmem + P::FA_COL + buf * 2 * P::CH_K + (pass == 1 ? P::CH_K : 0) + j * 8,
                     dW1[pass == 2] + (uint64_t)(koff >> 4), ID_FWD, (c | pass | j) != 0);
            }
          }
          umma_commit(&bars[B_FEMPTY0 + buf]);
          if (c == P::NCH - 1) umma_commit(&bars[B_D1FULL]);
        }
        __syncwarp();
      }
    }
  } else {
    const int hf = warp >> 2;
    const int s = tid & (TS - 1);
    const SegConst sc0 = {p.seg0, S::STRIDE, (float)p.seg0, 1.f / (float)p.seg0};
    const float* w0 = p.packed + p.off[0] + hf * HH;
    const LagrangeTable& lag = p.lag;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* xs = sX + (hf * TS + s) * 4;
    const float* xo = sX + ((hf ^ 1) * TS + s) * 4;
    uint32_t q_chunk = 0, n_tile = 0;
    auto layer0 = [&](long long tile_, float* h_) {
      const long long b_ = tile_ * TS + s;
      float x0[IN];
#pragma unroll
      for (int i = 0; i < IN; ++i) x0[i] = 0.f;
      if (b_ < p.B) {
        if (p.x != nullptr) {
#pragma unroll
          for (int i = 0; i < IN; ++i) x0[i] = __ldg(p.x + b_ * IN + i);
        } else {
          hoir::mesh_position(p.mesh, p.mesh.first_pixel + b_, x0, IN / 2);
        }
      }
#pragma unroll
      for (int o = 0; o < HH; ++o) h_[o] = 0.f;
#pragma unroll
      for (int i = 0; i < IN; ++i) {
        int id;
        float xi, l[N];
        if (p.pow2_0) locate<true>(x0[i], sc0, id, xi);
        else locate<false>(x0[i], sc0, id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        contract_fwd_p<N, HH, HP, true>(w0 + ((size_t)i * p.nodes0 + id * S::STRIDE) * HP, l, h_);
      }
      float inv_d_ = 1.f;
      int kidx_ = 0;
      if (p.normalize) maxabs_pair<HH>(h_, hf, p.norm_eps, xs, xo, warp, inv_d_, kidx_);
    };
    float hN[HH];
    if ((long long)blockIdx.x < tiles) layer0(blockIdx.x, hN);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      const long long b = tile * TS + s;
      // ---- hidden layer on the tensor cores: this thread's 4 inputs of every chunk -> TMEM
#pragma unroll
      for (int c = 0; c < P::NCH; ++c, ++q_chunk) {
        const uint32_t buf = q_chunk & 1, ph = (q_chunk >> 1) & 1;
        float v[P::HK], lo[P::HK];
#pragma unroll
        for (int m = 0; m < P::QI; ++m) {
          int id;
          float xi, l[N];
          locate2(hN[c * P::QI + m], id, xi);
          hoir::lagrange<N>(lag.X, lag.D, xi, l);
          densify<S, ND>(id, l, v + m * ND);
        }
#pragma unroll
        for (int j = 0; j < P::HK; ++j) lo[j] = tf32_lo(v[j]);
        mbar_wait(&bars[B_FEMPTY0 + buf], ph ^ 1);
        fence_after();
        const uint32_t col = tm_lane + P::FA_COL + buf * 2 * P::CH_K + hf * P::HK;
        tmem_stK<P::HK>(col, v);
        tmem_stK<P::HK>(col + P::CH_K, lo);
        tmem_wait_st();
        fence_before();
        mbar_arrive(&bars[B_FFULL0 + buf]);
      }
      // ---- while the tensor pipe finishes D1: layer 0 of the next tile
      if (tile + gridDim.x < tiles) layer0(tile + gridDim.x, hN);
      float h[HH];
      mbar_wait(&bars[B_D1FULL], n_tile & 1);
      fence_after();
      tmem_ld20(tm_lane + P::D1_COL + hf * HH, h);
      tmem_wait_ld();
      fence_before();
      mbar_arrive(&bars[B_D1EMPTY]);
      float inv_d1 = 1.f;
      int kidx1 = 0;
      if (p.normalize) maxabs_pair<HH>(h, hf, p.norm_eps, xs, xo, warp, inv_d1, kidx1);
      // ---- output layer (FFMA) over own 20 inputs, pair sum
      float yo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < HH; ++j) {
        const int i = hf * HH + j;
        int id;
        float xi, l[N];
        locate2(h[j], id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        contract_fwd<N, OUT, OP, false>(sW2 + ((size_t)i * S::NODES_O + id * S::STRIDE) * OP, l, yo);
      }
      if (hf == 1) *reinterpret_cast<float4*>(xs) = make_float4(yo[0], yo[1], yo[2], 0.f);
      bar_pair(warp);
      if (hf == 0 && b < p.B) {
        const float4 o4 = *reinterpret_cast<const float4*>(xo);
        const float other[3] = {o4.x, o4.y, o4.z};
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          const float yv = yo[o] + other[o];
          p.y[b * OUT + o] = p.post_scale ? 0.5f * (yv + 1.f) : yv;
        }
      }
      bar_pair(warp);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(F::TMEM_COLS));
}

}  // namespace tc
