// fused_mlp.cu -- whole-network kernels for the coordinate -> RGB implicit MLP
// (HighOrderMLP of piecewise-Lagrange layers with MaxAbsNormalization in between, as built at
// /root/reference/high_order_implicit_representation/networks.py:41-55).
//
// One CTA = 256 threads = one tile of 256 samples; the grid is persistent (one CTA per SM).
//   forward : thread-per-sample.  Layer-0 inputs come from x[B][IN] or are generated from the
//             pixel index (positions_from_mesh semantics); hidden/output weights live in shared
//             memory as [in][node][out_pad] so a thread reads the n*out_pad floats of its node
//             window with LDS.128; activations go layer to layer through registers and
//             thread-private shared-memory columns -- never through HBM.
//   backward: thread-per-sample for dL/dx (basis recomputed from the saved layer inputs);
//             dL/dw as a block-level contraction over the tile's samples: sample threads stage
//             the densified basis rows and dL/dy rows in shared memory, then "owner" threads
//             (input, output group, sample phase) accumulate in registers, reduce the sample
//             phases with warp shuffles and add into shared-memory accumulators that are
//             flushed to HBM once per CTA.  Layer 0 (many segments) merges runs of equal
//             segment and flushes with global atomics.
// Algorithm: oracle/hol_oracle.py.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "hoir_common.cuh"
#include "opt_tail.cuh"

namespace {

constexpr int BT = 256;  // samples per tile == threads per CTA
constexpr int CI = 4;    // inputs per dL/dw chunk
constexpr int RECW = 8;  // floats per densified basis row

template <int KIND_, int N_, int IN_, int HID_, int OUT_, int NH_, int SEGH_, int SEGO_>
struct Shape {
  static constexpr int KIND = KIND_, N = N_, IN = IN_, HID = HID_, OUT = OUT_, NH = NH_;
  static constexpr int SEGH = SEGH_, SEGO = SEGO_;
  static constexpr bool CONT = KIND_ == HOIR_CONTINUOUS;
  static constexpr int STRIDE = CONT ? N_ - 1 : N_;
  static constexpr int NODES_H = STRIDE * SEGH_ + (CONT ? 1 : 0);
  static constexpr int NODES_O = STRIDE * SEGO_ + (CONT ? 1 : 0);
  static constexpr int HP = (HID_ + 3) & ~3;
  static constexpr int OP = (OUT_ + 3) & ~3;
  static constexpr int GP = (HID_ > OUT_ ? HID_ : OUT_) | 1;  // odd row stride of the dL/dy tile
  static constexpr int WH = HID_ * NODES_H * HP;               // floats of one hidden layer
  static constexpr int WO = HID_ * NODES_O * OP;
  static_assert(N_ >= 2 && N_ <= 3, "fused kernel: n in {2,3}");
  static_assert(NODES_H <= RECW && NODES_O <= RECW, "densified row too wide");
  static_assert(IN_ <= 8 && IN_ <= CI * 2, "layer-0 inputs are kept in registers");
  static_assert(IN_ * HID_ <= BT, "layer-0 dL/dw owners: one thread per (in, out) link");
  // shared-memory plan (floats)
  static constexpr int S_W = NH_ * WH + WO;                    // weights of layers 1..NH+1
  static constexpr int S_A = (NH_ + 1) * HID_ * BT;            // saved layer inputs (train only)
  static constexpr int S_G = BT * GP;
  static constexpr int S_REC = CI * BT * RECW;
  static constexpr int S_TRAIN_NOACC = S_W + S_A + S_G + S_REC;
  static constexpr bool ACC_SMEM = (size_t)(S_TRAIN_NOACC + S_W) * 4 <= 227 * 1024;
  static constexpr int S_TRAIN = S_TRAIN_NOACC + (ACC_SMEM ? S_W : 0);
  static constexpr int S_FWD = S_W + HID_ * BT;
};

struct FusedParams {
  long long B;
  const float* x;            // [B][IN] or nullptr (mesh)
  const int* perm;           // nullptr, or [B] source index of every sample slot (tensor-core training kernel only)
  hoir_mesh_desc mesh;
  int target_kind;
  const void* target;
  const float* gy_ext;
  const float* packed;
  long long off[HOIR_MAX_LAYERS];
  float* gw[HOIR_MAX_LAYERS];
  float* loss_sum;
  float loss_scale;
  float* y;
  int post_scale;
  int seg0, nodes0, pow2_0;
  long long alt_off;         // float offset of the [in][half][node][hid/2] copy of the layer-0 weights (tensor-core entries)
  int identity;              // every layer is a Polynomial (single window, xi = x)
  int normalize;
  float norm_eps;
  // two-pass layer-0 dL/dw for incoherent batches (generic x, large B): the main kernel streams
  // dL/d(layer-0 output) rows, (window, basis) records and segment ids to this workspace and a
  // segment-owner kernel accumulates them without atomics on the hot addresses.  NULL: in-kernel.
  float* ws_g0;              // [B][hid]
  float4* ws_rec;            // [IN][B]
  unsigned short* ws_seg;    // [IN][B]
  LagrangeTable lag;
  OptTail tail;              // kind != 0: optimizer tail inside the kernel (behind a grid barrier)
};

struct SegConst {
  int segments, stride;
  float segf, inv_segf;
  int identity;              // Polynomial layer: one window, basis evaluated at x itself (no coordinate map)
};

__device__ __forceinline__ void locate2(float x, int& id, float& xi);
template <bool POW2>
__device__ __forceinline__ void locate(float x, const SegConst& sc, int& id, float& xi) {
  if (POW2 && sc.segments == 2) {   // compile-time after inlining for the hidden / output layers
    locate2(x, id, xi);
    return;
  }
  if (sc.identity) {
    id = 0;
    xi = x;
    return;
  }
  id = hoir::segment_index(x, 1.f, 2.f, sc.segf, sc.segments);
  xi = hoir::local_coordinate<POW2>(x, id, 1.f, 2.f, sc.segf, sc.inv_segf);
}
// two segments on [-1, 1] (hidden / output layers of every reference config): the same roundings
// as the generic form collapse to  id = (fl(x + 1) >= 1),  xi = 2 * (x - (id - 1)) - 1.
__device__ __forceinline__ void locate2(float x, int& id, float& xi) {
  const float t = __fadd_rn(x, 1.f);            // ((x + 1) / 2) * 2: both scalings are exact
  id = t >= 1.f ? 1 : 0;                        // trunc + clamp to {0, 1}; NaN -> 0
  const float q = id ? x : t;                   // x - x_min: x - 0, or x - (-1) = t (the same rounding)
  xi = __fmaf_rn(2.f, q, -1.f);                 // 2 q is exact, so the fused form rounds once exactly like 2 q - 1
}

// y[O] += sum_j l[j] * W[(base + j)][0..OP)   (W row-major, OP floats per node, 16B aligned)
template <int N, int O, int OPAD, bool GLOBAL>
__device__ __forceinline__ void contract_fwd(const float* __restrict__ wrow, const float* l,
                                             float* __restrict__ acc) {
  const float4* wp = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int q = 0; q < OPAD / 4; ++q) {
      float4 w = GLOBAL ? __ldg(wp + j * (OPAD / 4) + q) : wp[j * (OPAD / 4) + q];
      if (4 * q + 0 < O) acc[4 * q + 0] = fmaf(l[j], w.x, acc[4 * q + 0]);
      if (4 * q + 1 < O) acc[4 * q + 1] = fmaf(l[j], w.y, acc[4 * q + 1]);
      if (4 * q + 2 < O) acc[4 * q + 2] = fmaf(l[j], w.z, acc[4 * q + 2]);
      if (4 * q + 3 < O) acc[4 * q + 3] = fmaf(l[j], w.w, acc[4 * q + 3]);
    }
  }
}

// t[j] = sum_o g[o] * W[(base + j)][o]
template <int N, int O, int OPAD>
__device__ __forceinline__ void contract_bwd(const float* __restrict__ wrow, const float* g,
                                             float* __restrict__ t) {
  const float4* wp = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int q = 0; q < OPAD / 4; ++q) {
      float4 w = wp[j * (OPAD / 4) + q];
      if (4 * q + 0 < O) s0 = fmaf(g[4 * q + 0], w.x, s0);
      if (4 * q + 1 < O) s1 = fmaf(g[4 * q + 1], w.y, s1);
      if (4 * q + 2 < O) s0 = fmaf(g[4 * q + 2], w.z, s0);
      if (4 * q + 3 < O) s1 = fmaf(g[4 * q + 3], w.w, s1);
    }
    t[j] = s0 + s1;
  }
}

// MaxAbsNormalization over the thread's W registers; first index wins ties.
template <int W>
__device__ __forceinline__ void maxabs_fwd(float* v, float eps, float& inv_d, int& kidx) {
  float m = -1.f, vk = 0.f;
  int k = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    float a = fabsf(v[i]);
    if (a > m) { m = a; k = i; vk = v[i]; }
  }
  const float d = __fadd_rn(m, eps);
  const float sgn = vk > 0.f ? 1.f : (vk < 0.f ? -1.f : 0.f);
  inv_d = __fdiv_rn(1.f, d);
#pragma unroll
  for (int i = 0; i < W; ++i) v[i] *= inv_d;   // within 1 ulp of x / d (the stand-alone kernel divides)
  kidx = sgn > 0.f ? k : (sgn < 0.f ? (k | 0x10000) : (k | 0x20000));
}

// One hidden/output layer forward for this thread's sample: inputs from the smem column a[i*BT].
template <int N, int I, int O, int OPAD, int NODES, bool POW2>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ a, const float* __restrict__ w,
                                          const SegConst& sc, const LagrangeTable& lag,
                                          float* __restrict__ acc) {
#pragma unroll
  for (int o = 0; o < O; ++o) acc[o] = 0.f;
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    const float x = a[i * BT];
    int id;
    float xi, l[N];
    locate<POW2>(x, sc, id, xi);
    hoir::lagrange<N>(lag.X, lag.D, xi, l);
    contract_fwd<N, O, OPAD, false>(w + ((size_t)i * NODES + id * sc.stride) * OPAD, l, acc);
  }
}

// dL/dx of one hidden/output layer for this thread's sample.  Overwrites a[i*BT] with dL/dx_i and
// returns sum_i dL/dx_i * x_i (needed by the MaxAbs backward).
template <int N, int I, int O, int OPAD, int NODES, bool POW2>
__device__ __forceinline__ float layer_bwd_x(float* __restrict__ a, const float* __restrict__ w,
                                             const SegConst& sc, const LagrangeTable& lag,
                                             const float* __restrict__ g) {
  float dot = 0.f;
#pragma unroll 2
  for (int i = 0; i < I; ++i) {
    const float x = a[i * BT];
    int id;
    float xi, l[N], d[N], t[N];
    locate<POW2>(x, sc, id, xi);
    hoir::lagrange_deriv<N>(lag.X, lag.D, xi, l, d);
    contract_bwd<N, O, OPAD>(w + ((size_t)i * NODES + id * sc.stride) * OPAD, g, t);
    float gx = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) gx = fmaf(d[j], t[j], gx);
    gx *= hoir::local_slope<POW2>(id, 2.f, sc.segf);
    dot = fmaf(gx, x, dot);
    a[i * BT] = gx;
  }
  return dot;
}

// dL/dw of a hidden/output layer for one tile: chunks of CI inputs.  Sample threads write the
// densified basis row (NODES coefficients, zeros outside the active window) of their sample;
// owner threads (il, og, bph) contract it with dL/dy over the samples b = bph (mod BPH).
template <int N, int STRIDE, int I, int O, int OPAD, int NODES, bool POW2, bool ACC_SMEM>
__device__ __forceinline__ void layer_bwd_w(const float* __restrict__ a, const float* __restrict__ G,
                                            int GP, float* __restrict__ rec,
                                            float* __restrict__ gws, float* __restrict__ gwg,
                                            const SegConst& sc, const LagrangeTable& lag) {
  constexpr int OGRP = O <= 8 ? 1 : (O <= 16 ? 2 : (O <= 32 ? 4 : 8));
  constexpr int OG = (O + OGRP - 1) / OGRP;
  static_assert(OG <= 8, "output group too wide");
  constexpr int BPH = (64 / OGRP) < 32 ? (64 / OGRP) : 32;
  constexpr int ACTIVE = CI * OGRP * BPH;
  const int tid = threadIdx.x;
  const int bph = tid % BPH, og = (tid / BPH) % OGRP, il = tid / (BPH * OGRP);
  for (int c = 0; c < (I + CI - 1) / CI; ++c) {
    // ---- stage densified rows of inputs c*CI .. c*CI+CI-1
#pragma unroll
    for (int k = 0; k < CI; ++k) {
      const int i = c * CI + k;
      float row[RECW];
#pragma unroll
      for (int q = 0; q < RECW; ++q) row[q] = 0.f;
      if (i < I) {
        const float x = a[i * BT + tid];
        int id;
        float xi, l[N];
        locate<POW2>(x, sc, id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        const int base = id * STRIDE;
#pragma unroll
        for (int q = 0; q < NODES; ++q) {
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (q - j >= 0 && (q - j) % STRIDE == 0)  // window starts are multiples of STRIDE
              row[q] = (base + j == q) ? l[j] : row[q];
        }
      }
      float4* dst = reinterpret_cast<float4*>(rec + ((size_t)k * BT + tid) * RECW);
      dst[0] = make_float4(row[0], row[1], row[2], row[3]);
      dst[1] = make_float4(row[4], row[5], row[6], row[7]);
    }
    __syncthreads();
    // ---- owners
    if (tid < ACTIVE) {
      const int i = c * CI + il;
      float acc[NODES][OG];
#pragma unroll
      for (int q = 0; q < NODES; ++q)
#pragma unroll
        for (int m = 0; m < OG; ++m) acc[q][m] = 0.f;
#pragma unroll 2
      for (int b = bph; b < BT; b += BPH) {
        const float4* rp = reinterpret_cast<const float4*>(rec + ((size_t)il * BT + b) * RECW);
        const float4 r0 = rp[0];
        float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NODES > 4) r1 = rp[1];
        const float cf[RECW] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        float gv[OG];
#pragma unroll
        for (int m = 0; m < OG; ++m) {
          const int o = og + OGRP * m;
          gv[m] = (o < O) ? G[b * GP + o] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < NODES; ++q)
#pragma unroll
          for (int m = 0; m < OG; ++m) acc[q][m] = fmaf(cf[q], gv[m], acc[q][m]);
      }
#pragma unroll
      for (int q = 0; q < NODES; ++q)
#pragma unroll
        for (int m = 0; m < OG; ++m) {
          float v = acc[q][m];
#pragma unroll
          for (int s = BPH / 2; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
          acc[q][m] = v;
        }
      if (bph == 0 && i < I) {
#pragma unroll
        for (int q = 0; q < NODES; ++q)
#pragma unroll
          for (int m = 0; m < OG; ++m) {
            const int o = og + OGRP * m;
            if (o < O) {
              if (ACC_SMEM) gws[((size_t)i * NODES + q) * OPAD + o] += acc[q][m];
              else if (acc[q][m] != 0.f) hoir::red_add(gwg + ((size_t)o * I + i) * NODES + q, acc[q][m]);
            }
          }
      }
    }
    __syncthreads();
  }
}

// dL/dw of layer 0 (many segments): owner (i, o) walks the tile's samples in order, merges runs
// of equal node window and flushes each run with global atomics (boundary layout).
template <int N, int IN, int HID>
__device__ __forceinline__ void layer0_bwd_w(const float* __restrict__ rec0,
                                             const float* __restrict__ G, int GP,
                                             float* __restrict__ gw0, int nodes0) {
  constexpr int PAIRS = IN * HID;
  constexpr int BPH0 = PAIRS > 128 ? 1 : (PAIRS > 64 ? 2 : (PAIRS > 32 ? 4 : (PAIRS > 16 ? 8 : 16)));
  constexpr int SPAN = BT / BPH0;
  const int tid = threadIdx.x;
  const int p = tid / BPH0, bp = tid % BPH0;
  if (p >= PAIRS) return;
  const int i = p / HID, o = p - i * HID;
  float* dst = gw0 + ((size_t)o * IN + i) * nodes0;
  int cur = -1;
  float a[N];
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = 0.f;
  const float4* rp = reinterpret_cast<const float4*>(rec0) + (size_t)i * BT + bp * SPAN;
  const float* gp = G + (size_t)(bp * SPAN) * GP + o;
#pragma unroll 4
  for (int b = 0; b < SPAN; ++b) {
    const float4 r = rp[b];
    const float gv = gp[b * GP];
    const int base = __float_as_int(r.x);
    if (base != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (a[j] != 0.f) hoir::red_add(dst + cur + j, a[j]);
      }
      cur = base;
#pragma unroll
      for (int j = 0; j < N; ++j) a[j] = 0.f;
    }
    const float l[3] = {r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] = fmaf(l[j], gv, a[j]);
  }
  if (cur >= 0) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (a[j] != 0.f) hoir::red_add(dst + cur + j, a[j]);
  }
}

template <class S, bool TRAIN>
__global__ void __launch_bounds__(BT, 1) fused_mlp_kernel(const FusedParams p) {
  constexpr int N = S::N, IN = S::IN, HID = S::HID, OUT = S::OUT, NH = S::NH;
  constexpr int HP = S::HP, OP = S::OP, GP = S::GP;
  constexpr bool POW2H = (S::SEGH & (S::SEGH - 1)) == 0, POW2O = (S::SEGO & (S::SEGO - 1)) == 0;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                  // hidden layers then output layer
  float* sA = sW + S::S_W;                           // [(NH+1)][HID][BT]   (TRAIN) / [HID][BT]
  float* sG = sA + S::S_A;                           // [BT][GP]            (TRAIN)
  float* sRec = sG + S::S_G;                         // [CI][BT][RECW]      (TRAIN)
  float* sAcc = sRec + S::S_REC;                     // like sW             (TRAIN && ACC_SMEM)
  const int tid = threadIdx.x;

  // ---- stage weights of layers 1..NH+1 (contiguous in the packed buffer) -----------------
  {
    const float4* src = reinterpret_cast<const float4*>(p.packed + p.off[1]);
    float4* dst = reinterpret_cast<float4*>(sW);
    for (int k = tid; k < S::S_W / 4; k += BT) dst[k] = __ldg(src + k);
    if (TRAIN && S::ACC_SMEM)
      for (int k = tid; k < S::S_W; k += BT) sAcc[k] = 0.f;
  }
  __syncthreads();

  const SegConst sc0 = {p.seg0, S::STRIDE, (float)p.seg0, 1.f / (float)p.seg0, p.identity};
  const SegConst sch = {S::SEGH, S::STRIDE, (float)S::SEGH, 1.f / (float)S::SEGH, p.identity};
  const SegConst sco = {S::SEGO, S::STRIDE, (float)S::SEGO, 1.f / (float)S::SEGO, p.identity};
  const float* w0 = p.packed + p.off[0];
  const LagrangeTable& lag = p.lag;
  float loss_local = 0.f;

  const long long tiles = (p.B + BT - 1) / BT;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b = tile * BT + tid;
    const bool live = b < p.B;
    // ---- layer-0 inputs
    float x0[IN];
#pragma unroll
    for (int i = 0; i < IN; ++i) x0[i] = 0.f;
    if (live) {
      if (p.x != nullptr) {
#pragma unroll
        for (int i = 0; i < IN; ++i) x0[i] = __ldg(p.x + b * IN + i);
      } else {
        hoir::mesh_position(p.mesh, p.mesh.first_pixel + b, x0, IN / 2);
      }
    }
    // ---- forward
    float h[HID];
    float inv_d[NH + 1];
    int kidx[NH + 1];
    float4 rec0[IN];  // layer-0 (node window, basis) rows, consumed by its dL/dw pass
    {
#pragma unroll
      for (int o = 0; o < HID; ++o) h[o] = 0.f;
#pragma unroll
      for (int i = 0; i < IN; ++i) {
        int id;
        float xi, l[N];
        if (p.pow2_0) locate<true>(x0[i], sc0, id, xi);
        else locate<false>(x0[i], sc0, id, xi);
        hoir::lagrange<N>(lag.X, lag.D, xi, l);
        const int base0 = id * S::STRIDE;
        contract_fwd<N, HID, HP, true>(w0 + ((size_t)i * p.nodes0 + base0) * HP, l, h);
        rec0[i] = make_float4(__int_as_float(base0), l[0], l[1], N > 2 ? l[2] : 0.f);
      }
    }
    float* aT = sA + tid;
#pragma unroll
    for (int l = 0; l < NH; ++l) {
      if (p.normalize) maxabs_fwd<HID>(h, p.norm_eps, inv_d[l], kidx[l]);
      float* a = aT + (TRAIN ? (size_t)l * HID * BT : 0);
#pragma unroll
      for (int i = 0; i < HID; ++i) a[i * BT] = h[i];
      layer_fwd<N, HID, HID, HP, S::NODES_H, POW2H>(a, sW + (size_t)l * S::WH, sch, lag, h);
    }
    if (p.normalize) maxabs_fwd<HID>(h, p.norm_eps, inv_d[NH], kidx[NH]);
    float yo[OUT];
    {
      float* a = aT + (TRAIN ? (size_t)NH * HID * BT : 0);
#pragma unroll
      for (int i = 0; i < HID; ++i) a[i * BT] = h[i];
      layer_fwd<N, HID, OUT, OP, S::NODES_O, POW2O>(a, sW + (size_t)NH * S::WH, sco, lag, yo);
    }
    if (!TRAIN) {
      if (live) {
#pragma unroll
        for (int o = 0; o < OUT; ++o)
          p.y[b * OUT + o] = p.post_scale ? 0.5f * (yo[o] + 1.f) : yo[o];
      }
      continue;
    }
    // ---- loss and dL/dy
    float g[HID];
#pragma unroll
    for (int o = 0; o < OUT; ++o) g[o] = 0.f;
    if (live) {
      if (p.gy_ext != nullptr) {
#pragma unroll
        for (int o = 0; o < OUT; ++o) g[o] = __ldg(p.gy_ext + b * OUT + o);
      } else {
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
          float t;
          if (p.target_kind == HOIR_TARGET_U8_IMAGE) {
            const unsigned char* img = static_cast<const unsigned char*>(p.target);
            const float u = (float)img[(p.mesh.first_pixel + b) * OUT + o];
            t = __fsub_rn(__fdiv_rn(__fmul_rn(u, 2.f), 255.f), 1.f);
          } else {
            t = __ldg(static_cast<const float*>(p.target) + b * OUT + o);
          }
          const float diff = yo[o] - t;
          loss_local = fmaf(diff, diff, loss_local);
          g[o] = 2.f * diff * p.loss_scale;
        }
      }
    }
    // ---- output layer backward
    {
      float* a = aT + (size_t)NH * HID * BT;
#pragma unroll
      for (int o = 0; o < OUT; ++o) sG[tid * GP + o] = g[o];
      __syncthreads();
      layer_bwd_w<N, S::STRIDE, HID, OUT, OP, S::NODES_O, POW2O, S::ACC_SMEM>(
          sA + (size_t)NH * HID * BT, sG, GP, sRec, sAcc + (size_t)NH * S::WH, p.gw[NH + 1], sco, lag);
      float dot = layer_bwd_x<N, HID, OUT, OP, S::NODES_O, POW2O>(a, sW + (size_t)NH * S::WH, sco, lag, g);
      // MaxAbs backward -> g[HID] = dL/d(hidden output NH-1 .. )
      if (p.normalize) {
        const int k = kidx[NH] & 0xffff;
        const float sgn = (kidx[NH] & 0x10000) ? -1.f : ((kidx[NH] & 0x20000) ? 0.f : 1.f);
#pragma unroll
        for (int i = 0; i < HID; ++i) {
          float gi = a[i * BT];
          if (i == k) gi -= sgn * dot;
          g[i] = gi * inv_d[NH];
        }
      } else {
#pragma unroll
        for (int i = 0; i < HID; ++i) g[i] = a[i * BT];
      }
      // ---- hidden layers, last to first
#pragma unroll
      for (int l = NH - 1; l >= 0; --l) {
        float* al = aT + (size_t)l * HID * BT;
#pragma unroll
        for (int o = 0; o < HID; ++o) sG[tid * GP + o] = g[o];
        __syncthreads();
        layer_bwd_w<N, S::STRIDE, HID, HID, HP, S::NODES_H, POW2H, S::ACC_SMEM>(
            sA + (size_t)l * HID * BT, sG, GP, sRec, sAcc + (size_t)l * S::WH, p.gw[l + 1], sch, lag);
        dot = layer_bwd_x<N, HID, HID, HP, S::NODES_H, POW2H>(al, sW + (size_t)l * S::WH, sch, lag, g);
        if (p.normalize) {
          const int k = kidx[l] & 0xffff;
          const float sgn = (kidx[l] & 0x10000) ? -1.f : ((kidx[l] & 0x20000) ? 0.f : 1.f);
#pragma unroll
          for (int i = 0; i < HID; ++i) {
            float gi = al[i * BT];
            if (i == k) gi -= sgn * dot;
            g[i] = gi * inv_d[l];
          }
        } else {
#pragma unroll
          for (int i = 0; i < HID; ++i) g[i] = al[i * BT];
        }
      }
      // ---- layer 0: dL/dw only
#pragma unroll
      for (int o = 0; o < HID; ++o) sG[tid * GP + o] = g[o];
#pragma unroll
      for (int i = 0; i < IN; ++i) reinterpret_cast<float4*>(sRec)[(size_t)i * BT + tid] = rec0[i];
      __syncthreads();
      layer0_bwd_w<N, IN, HID>(sRec, sG, GP, p.gw[0], p.nodes0);
      __syncthreads();
    }
  }
  if (!TRAIN) return;
  // ---- flush: loss and the shared-memory dL/dw accumulators (boundary layout [out][in][nodes])
  if (p.gy_ext == nullptr) {
    float s = hoir::warp_sum(loss_local) * p.loss_scale;
    if ((tid & 31) == 0 && s != 0.f) hoir::red_add(p.loss_sum, s);
  }
  if (S::ACC_SMEM) {
    __syncthreads();
    for (int l = 0; l < NH; ++l) {
      const float* acc = sAcc + (size_t)l * S::WH;
      for (int k = tid; k < HID * S::NODES_H * HID; k += BT) {
        const int o = k % HID, r = k / HID, q = r % S::NODES_H, i = r / S::NODES_H;
        const float v = acc[((size_t)i * S::NODES_H + q) * HP + o];
        if (v != 0.f) hoir::red_add(p.gw[l + 1] + ((size_t)o * HID + i) * S::NODES_H + q, v);
      }
    }
    const float* acc = sAcc + (size_t)NH * S::WH;
    for (int k = tid; k < HID * S::NODES_O * OUT; k += BT) {
      const int o = k % OUT, r = k / OUT, q = r % S::NODES_O, i = r / S::NODES_O;
      const float v = acc[((size_t)i * S::NODES_O + q) * OP + o];
      if (v != 0.f) hoir::red_add(p.gw[NH + 1] + ((size_t)o * HID + i) * S::NODES_O + q, v);
    }
  }
  hoir::opt_tail_in_kernel(p.tail);
}

}  // namespace
#include "fused_mlp_tc.cuh"
namespace {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ w, float* __restrict__ packed, int in_f,
                            int out_f, int nodes, int out_pad) {
  const long long total = (long long)in_f * nodes * out_pad;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(k % out_pad);
    const long long r = k / out_pad;
    const int q = (int)(r % nodes), i = (int)(r / nodes);
    packed[k] = o < out_f ? w[((size_t)o * in_f + i) * nodes + q] : 0.f;
  }
}

// second copy of the layer-0 weights for the tensor-core training kernel: [in][half][node][hid / 2], so that the three
// node rows a thread (one half of the outputs of a sample) reads are ONE contiguous run -- 2-3 cache lines instead
// of 3-6 when every lane of a warp gathers from a different segment (incoherent batches)
__global__ void pack_alt_kernel(const float* __restrict__ w, float* __restrict__ alt, int in_f, int out_f, int nodes) {
  const int hh = out_f / 2;
  const long long total = (long long)in_f * nodes * out_f;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(k % hh);
    long long r = k / hh;
    const int q = (int)(r % nodes);
    r /= nodes;
    const int half = (int)(r & 1), i = (int)(r >> 1);
    alt[k] = w[((size_t)(half * hh + c) * in_f + i) * nodes + q];
  }
}

bool mlp_desc_ok(const hoir_mlp_desc* m) {
  if (!m || m->num_layers < 2 || m->num_layers > HOIR_MAX_LAYERS) return false;
  for (int l = 0; l < m->num_layers; ++l) {
    if (hoir_layer_nodes(&m->layers[l]) < 1) return false;
    if (m->layers[l].in_features < 1 || m->layers[l].out_features < 1) return false;
    if (l > 0 && m->layers[l].in_features != m->layers[l - 1].out_features) return false;
  }
  return true;
}

// Second pass of the layer-0 dL/dw for incoherent batches.  CTA (chunk of 16384 samples, input i):
//   1. counting sort of the chunk's samples by segment in shared memory;
//   2. every warp reduces blocks of 128 consecutive sorted entries in registers (lane <-> outputs o and
//      o + 32, basis record broadcast, dL/dy row coalesced) and adds a finished segment run into the CTA's
//      shared-memory slab  gw0[:, i, :]  ([node][out], 32 KB for 201 x 40);
//   3. the slab is flushed once: one fire-and-forget global add per weight and CTA, nodes fastest so that
//      a warp's adds fall into consecutive addresses of the boundary layout [out][in][nodes].
// ~0.25 global adds per (sample, input) instead of the 120 of a per-sample scatter, balanced for any
// distribution of the inputs.
constexpr int GW0_CHUNK = 32768, GW0_SHIFT = 15, GW0_MAXSEG = 1024, GW0_BLOCK = 128, GW0_PITCH_PAD = 1, GW0_THREADS = 512, GW0_WARPS = GW0_THREADS / 32;
template <int N>
__global__ void __launch_bounds__(GW0_THREADS) gw0_slab_kernel(long long B, int chunk, int IN, int hid, int seg0, int stride, int nodes0,
                                                       const unsigned short* __restrict__ ws_seg,
                                                       const float4* __restrict__ ws_rec, const float* __restrict__ ws_g0,
                                                       float* __restrict__ gw0) {
  extern __shared__ __align__(16) unsigned char gw0_smem[];
  unsigned int* s_ent = reinterpret_cast<unsigned int*>(gw0_smem);      // sorted: (segment << GW0_SHIFT) | sample offset
  int* s_cnt = reinterpret_cast<int*>(s_ent + chunk);                  // chunk is a multiple of 256
  int* s_pos = s_cnt + GW0_MAXSEG;
  float* slab = reinterpret_cast<float*>(s_pos + GW0_MAXSEG);           // [nodes0][pitch]
  const int pitch = hid + GW0_PITCH_PAD;
  const int i = blockIdx.x;                                             // input fastest: the IN CTAs of a chunk run together
  const long long b0 = (long long)blockIdx.y * chunk;
  const int n = (int)((B - b0) < chunk ? (B - b0) : chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned short* seg = ws_seg + (size_t)i * B + b0;
  const float4* rec = ws_rec + (size_t)i * B + b0;
  for (int k = threadIdx.x; k < seg0; k += GW0_THREADS) s_cnt[k] = 0;
  if (hid != 40)                                                         // (the 40-wide path adds straight to global memory)
    for (int k = threadIdx.x; k < nodes0 * pitch; k += GW0_THREADS) slab[k] = 0.f;
  __syncthreads();
  // histogram with warp-aggregated shared atomics (coherent batches put a whole warp into one bin)
  // (four key loads in flight per lane: the passes over the keys are latency-, not bandwidth-bound)
  for (int t0 = warp * 32; t0 < n; t0 += 4 * GW0_THREADS) {
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * GW0_THREADS + lane;
      keys[u] = t < n ? (int)__ldg(seg + t) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u * GW0_THREADS >= n) break;                    // warp-uniform
      const unsigned peers = __match_any_sync(0xffffffffu, keys[u]);
      if (keys[u] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[keys[u]], __popc(peers));
    }
  }
  __syncthreads();
  if (warp == 0) {
    int running = 0;
    for (int base = 0; base < seg0; base += 32) {
      const int v = base + lane < seg0 ? s_cnt[base + lane] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (base + lane < seg0) s_pos[base + lane] = running + incl - v;
      running += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int t0 = warp * 32; t0 < n; t0 += 4 * GW0_THREADS) {
    int keys[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * GW0_THREADS + lane;
      keys[u] = t < n ? (int)__ldg(seg + t) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u * GW0_THREADS >= n) break;                    // warp-uniform
      const int t = t0 + u * GW0_THREADS + lane, key = keys[u];
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(peers) - 1;
      int base = 0;
      if (key >= 0 && lane == leader) base = atomicAdd(&s_pos[key], __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (key >= 0) s_ent[base + __popc(peers & ((1u << lane) - 1u))] = ((unsigned)key << GW0_SHIFT) | (unsigned)t;
    }
  }
  __syncthreads();
  // reduction over the sorted entries.  hid % 8 == 0 (the 40-wide family): a group of 4 lanes owns a contiguous
  // range of the sorted list, each lane hid/4 outputs -- per entry one broadcast record load, hid/8 8-byte
  // loads of the dL/dy row and 3 * hid/4 FMAs per lane (5 warp-instructions per entry instead of 30); a segment
  // change flushes the group's registers into the CTA's slab.  Otherwise: warp per entry, lane <-> outputs.
  if (hid == 40) {
    constexpr int OPL = 10, DEPTH = 4;                        // outputs per lane; entries in flight per lane
    const int grp = threadIdx.x >> 2, sub = threadIdx.x & 3, ngrp = GW0_THREADS / 4;
    const int per = (n + ngrp - 1) / ngrp;
    const int k0 = grp * per, k1 = min(n, k0 + per);
    if (k0 < k1) {
      int cur = (int)(s_ent[k0] >> GW0_SHIFT);
      float acc[N][OPL];
#pragma unroll
      for (int j = 0; j < N; ++j)
#pragma unroll
        for (int o = 0; o < OPL; ++o) acc[j][o] = 0.f;
      // a finished segment run goes straight to global memory with fire-and-forget adds (the shared-memory float
      // atomics this replaced are compare-and-swap loops: 35 % of the kernel's stall samples)
      float* const gbase = gw0 + ((size_t)(sub * OPL) * IN + i) * nodes0;
      auto flush = [&](int sg) {
        float* dst = gbase + sg * stride;
#pragma unroll
        for (int o = 0; o < OPL; ++o)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            if (acc[j][o] != 0.f) hoir::red_add(dst + (size_t)o * IN * nodes0 + j, acc[j][o]);
            acc[j][o] = 0.f;
          }
      };
      // batches of DEPTH entries: all their record / dL/dy-row loads are issued before the first is consumed
      for (int k = k0; k < k1; k += DEPTH) {
        float4 rc[DEPTH];
        float2 g2[DEPTH][OPL / 2];
        int sgs[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          const bool ok = k + u < k1;
          const unsigned e = ok ? s_ent[k + u] : 0u;
          const int t = (int)(e & ((1u << GW0_SHIFT) - 1u));
          sgs[u] = ok ? (int)(e >> GW0_SHIFT) : -1;
          rc[u] = __ldg(rec + t);
          const float2* grow = reinterpret_cast<const float2*>(ws_g0 + (size_t)(b0 + t) * 40 + sub * OPL);
#pragma unroll
          for (int o = 0; o < OPL / 2; ++o) g2[u][o] = __ldg(grow + o);
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          if (sgs[u] < 0) break;
          if (sgs[u] != cur) {
            flush(cur);
            cur = sgs[u];
          }
          const float l[3] = {rc[u].y, rc[u].z, rc[u].w};
#pragma unroll
          for (int j = 0; j < N; ++j)
#pragma unroll
            for (int o = 0; o < OPL / 2; ++o) {
              acc[j][2 * o] = fmaf(l[j], g2[u][o].x, acc[j][2 * o]);
              acc[j][2 * o + 1] = fmaf(l[j], g2[u][o].y, acc[j][2 * o + 1]);
            }
        }
      }
      flush(cur);
    }
    return;                                                   // nothing went through the shared-memory slab
  } else {
  const bool hi = lane + 32 < hid, lo = lane < hid;
  for (int k0 = warp * GW0_BLOCK; k0 < n; k0 += GW0_WARPS * GW0_BLOCK) {
    const int k1 = k0 + GW0_BLOCK < n ? k0 + GW0_BLOCK : n;
    int cur = (int)(s_ent[k0] >> GW0_SHIFT);
    float acc[N][2];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j][0] = acc[j][1] = 0.f;
    auto flush = [&](int sg) {
      float* dst = slab + (size_t)(sg * stride) * pitch + lane;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (lo && acc[j][0] != 0.f) atomicAdd(dst + j * pitch, acc[j][0]);
        if (hi && acc[j][1] != 0.f) atomicAdd(dst + j * pitch + 32, acc[j][1]);
        acc[j][0] = acc[j][1] = 0.f;
      }
    };
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
      const unsigned e = s_ent[k];
      const int t = (int)(e & ((1u << GW0_SHIFT) - 1u)), sg = (int)(e >> GW0_SHIFT);
      const float4 rc = __ldg(rec + t);
      const float* grow = ws_g0 + (size_t)(b0 + t) * hid;
      const float g0 = lo ? __ldg(grow + lane) : 0.f;
      const float g1 = hi ? __ldg(grow + 32 + lane) : 0.f;
      if (sg != cur) {
        flush(cur);
        cur = sg;
      }
      const float l[3] = {rc.y, rc.z, rc.w};
#pragma unroll
      for (int j = 0; j < N; ++j) {
        acc[j][0] = fmaf(l[j], g0, acc[j][0]);
        acc[j][1] = fmaf(l[j], g1, acc[j][1]);
      }
    }
    flush(cur);
  }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < hid * nodes0; k += GW0_THREADS) {
    const int o = k / nodes0, q = k - o * nodes0;
    const float v = slab[(size_t)q * pitch + o];
    if (v != 0.f) hoir::red_add(gw0 + ((size_t)o * IN + i) * nodes0 + q, v);
  }
}
inline size_t gw0_smem_bytes(int nodes0, int hid, long long chunk = GW0_CHUNK) {
  return (size_t)chunk * 4 + 2 * GW0_MAXSEG * 4 + (hid == 40 ? 0 : (size_t)nodes0 * (hid + GW0_PITCH_PAD) * 4);
}

// ------------------------------------------------------------------------------------------
// Batch sort by layer-0 cell (explicit batches in arbitrary order: the reference's DataLoader(shuffle=True),
// /root/reference/high_order_implicit_representation/single_image_dataset.py:312-320).
// MSE(mean) and the summed gradients do not depend on the order of the samples, so the training kernel may visit them
// in any order; visited by (segment of x0, segment of x1[, low bits of the segments of x2, x3]) the layer-0 weight
// reads of a warp fall into one or two node windows again and the in-kernel run-merged owner walk of the layer-0
// dL/dw works as for contiguous pixel tiles -- no dL/dy workspace (185 MB written + re-read per 2^20 samples), no
// second pass.  Counting sort: (1) key + rank within its bin by one returning atomic per sample, the last CTA
// scans the histogram; (2) scatter.  Keys only need to be coherent, not exact.
// ------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256, SORT_MAX_BINS = 1 << 18;
struct SortPlan {
  int in_f, seg0, shift, cells_1d, minor_bits, bins, bins_pad;   // bins_pad: multiple of 4 * SORT_THREADS (one chunk)
  float segf;
};
__device__ __forceinline__ int sort_key(const SortPlan& sp, const float* xv) {
  const int s0 = hoir::segment_index(xv[0], 1.f, 2.f, sp.segf, sp.seg0) >> sp.shift;
  const int s1 = sp.in_f > 1 ? hoir::segment_index(xv[1], 1.f, 2.f, sp.segf, sp.seg0) >> sp.shift : 0;
  int key = s0 * sp.cells_1d + s1;
  if (sp.minor_bits) {   // rotated coordinates (x2, x3 are functions of x0, x1): 2-3 neighbouring segments per cell
    const int s2 = hoir::segment_index(xv[2], 1.f, 2.f, sp.segf, sp.seg0) & 3;
    const int s3 = hoir::segment_index(xv[3], 1.f, 2.f, sp.segf, sp.seg0) & 3;
    key = (key << 4) | (s2 << 2) | s3;
  }
  return key;
}
// x != nullptr: sample b = row b of x[B][in_f]; else sample b = pixel idx[b] (or first_pixel + b) of the mesh
__device__ __forceinline__ void sort_load(const SortPlan& sp, const float* __restrict__ x, const hoir_mesh_desc& mesh,
                                          const int* __restrict__ idx, long long b, float* xv) {
  if (x != nullptr) {
    if (sp.in_f == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + b);
      xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = i < sp.in_f ? __ldg(x + b * sp.in_f + i) : 0.f;
    }
  } else {
    xv[0] = xv[1] = xv[2] = xv[3] = 0.f;
    hoir::mesh_position(mesh, idx != nullptr ? (long long)__ldg(idx + b) : mesh.first_pixel + b, xv, 2);
  }
}
// One cooperative launch, five phases separated by grid barriers (the whole pre-pass is latency-bound: as three
// launches + a memset + a single-CTA scan it took 94 us at any batch size):
//   P0 zero the histogram   P1 key + rank within its bin (one returning atomic per sample)
//   P2 per-chunk sums (1024 bins per chunk)   P3 exclusive scan (chunk base = sum of the preceding chunk sums)
//   P4 out[offset[key] + rank] = source index of the sample (its row, or its pixel)
struct SortArgs {
  SortPlan sp;
  long long B;
  const float* x;
  hoir_mesh_desc mesh;
  const int* idx;
  long long first;
  unsigned int* sync;        // [2] phase counter / depart counter (zero on entry, re-armed by the last CTA to leave)
  unsigned int* partial;     // [bins_pad / 1024]
  unsigned int* hist;        // [bins_pad]
  uint2* keyrank;            // [B]
  int* out;                  // [B]
};
__device__ __forceinline__ unsigned int block_sum_256(unsigned int v, unsigned int* sh) {   // all threads get the sum
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned int t = 0;
#pragma unroll
  for (int w = 0; w < SORT_THREADS / 32; ++w) t += sh[w];
  return t;
}
__global__ void __launch_bounds__(SORT_THREADS) batch_sort_kernel(const SortArgs a) {
  __shared__ unsigned int sh[SORT_THREADS / 32 + 1];
  const long long gtid = blockIdx.x * (long long)SORT_THREADS + threadIdx.x, gthreads = (long long)gridDim.x * SORT_THREADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint4* hist4 = reinterpret_cast<uint4*>(a.hist);
  for (long long k = gtid; k < a.sp.bins_pad / 4; k += gthreads) hist4[k] = make_uint4(0u, 0u, 0u, 0u);
  hoir::grid_sync_phase(a.sync, 1);
  for (long long b = gtid; b < a.B; b += gthreads) {
    float xv[4];
    sort_load(a.sp, a.x, a.mesh, a.idx, b, xv);
    const int key = sort_key(a.sp, xv);
    a.keyrank[b] = make_uint2((unsigned)key, atomicAdd(&a.hist[key], 1u));
  }
  hoir::grid_sync_phase(a.sync, 2);
  const int nchunks = a.sp.bins_pad / (4 * SORT_THREADS);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const uint4 v = __ldcg(hist4 + (size_t)c * SORT_THREADS + tid);
    const unsigned int t = block_sum_256(v.x + v.y + v.z + v.w, sh);
    if (tid == 0) a.partial[c] = t;
  }
  hoir::grid_sync_phase(a.sync, 3);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    unsigned int pre = 0;
    for (int j = tid; j < c; j += SORT_THREADS) pre += __ldcg(a.partial + j);
    const unsigned int base = block_sum_256(pre, sh);
    uint4 v = __ldcg(hist4 + (size_t)c * SORT_THREADS + tid);
    const unsigned int mine = v.x + v.y + v.z + v.w;
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) wbase += w < warp ? sh[w] : 0u;
    const unsigned int e0 = base + wbase + incl - mine, e1 = e0 + v.x, e2 = e1 + v.y, e3 = e2 + v.z;
    hist4[(size_t)c * SORT_THREADS + tid] = make_uint4(e0, e1, e2, e3);
  }
  hoir::grid_sync_phase(a.sync, 4);
  for (long long b = gtid; b < a.B; b += gthreads) {
    const uint2 kr = a.keyrank[b];
    a.out[__ldcg(a.hist + kr.x) + kr.y] = a.idx != nullptr ? __ldg(a.idx + b) : (int)(a.first + b);
  }
  if (tid == 0 && atomicAdd(&a.sync[1], 1u) == gridDim.x - 1) {   // the last CTA to leave re-arms the counters
    a.sync[0] = 0u;
    a.sync[1] = 0u;
    __threadfence();
  }
}
// coherence probe: fraction of (sample, input) pairs whose layer-0 segment differs from the previous sample's in the
// sorted order (image pixels: ~0.03; independent random inputs: ~0.5)
__global__ void __launch_bounds__(SORT_THREADS) sort_breaks_kernel(const SortPlan sp, long long B, const float* __restrict__ x,
                                                                   const hoir_mesh_desc mesh, const int* __restrict__ order,
                                                                   unsigned long long* __restrict__ breaks) {
  unsigned int local = 0;
  for (long long k = blockIdx.x * (long long)SORT_THREADS + threadIdx.x + 1; k < B; k += (long long)gridDim.x * SORT_THREADS) {
    float a[4], c[4];
    if (x != nullptr) {
      sort_load(sp, x, mesh, nullptr, order[k - 1], a);
      sort_load(sp, x, mesh, nullptr, order[k], c);
    } else {
      sort_load(sp, nullptr, mesh, order, k - 1, a);
      sort_load(sp, nullptr, mesh, order, k, c);
    }
    for (int i = 0; i < sp.in_f && i < 4; ++i)
      local += hoir::segment_index(a[i], 1.f, 2.f, sp.segf, sp.seg0) != hoir::segment_index(c[i], 1.f, 2.f, sp.segf, sp.seg0);
  }
  local = (unsigned int)hoir::warp_sum((float)local);   // < 2^24 per warp: exact in fp32
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(breaks, (unsigned long long)local);
}

bool make_sort_plan(const hoir_layer_desc& l0, SortPlan* sp) {
  if ((l0.kind != HOIR_CONTINUOUS && l0.kind != HOIR_DISCONTINUOUS) || l0.length != 2.f || l0.periodicity > 0.f) return false;
  if (l0.in_features < 1 || l0.in_features > 4) return false;
  sp->in_f = l0.in_features;
  sp->seg0 = l0.segments;
  sp->segf = (float)l0.segments;
  sp->minor_bits = l0.in_features == 4 ? 4 : 0;
  sp->shift = 0;
  auto bins_of = [&](int sh) {
    const long long c = ((l0.segments - 1) >> sh) + 1;
    return (l0.in_features > 1 ? c * c : c) << sp->minor_bits;
  };
  while (bins_of(sp->shift) > SORT_MAX_BINS) ++sp->shift;
  sp->cells_1d = l0.in_features > 1 ? ((l0.segments - 1) >> sp->shift) + 1 : 1;
  sp->bins = (int)bins_of(sp->shift);
  const int q = 4 * SORT_THREADS;
  sp->bins_pad = (sp->bins + q - 1) / q * q;
  return true;
}

// perm_out[0 .. B): the samples' source indices in cell order.  x: rows of x; else pixels idx[b] / first_pixel + b.
int batch_sort_launch(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh, const int* idx,
                      int* perm_out, float* breaks_host, cudaStream_t st, const char* where) {
  SortPlan sp;
  if (!mlp_desc_ok(m) || !make_sort_plan(m->layers[0], &sp))
    return hoir_set_error(HOIR_EUNSUPPORTED, "batch sort: layer 0 must be piecewise with 1..4 inputs, length 2", where);
  if (B < 0 || B > 0x7fffffffLL) return hoir_set_error(HOIR_EINVAL, "bad batch size", where);
  if ((x == nullptr) == (mesh == nullptr)) return hoir_set_error(HOIR_EINVAL, "exactly one of x / mesh must be given", where);
  if (mesh != nullptr && (mesh->rotations * 2 != sp.in_f || mesh->rows < 1 || mesh->cols < 1 ||
                          (long long)mesh->rows * mesh->cols > 0x7fffffffLL))
    return hoir_set_error(HOIR_EINVAL, "mesh descriptor does not match layer 0", where);
  if (B == 0) {
    if (breaks_host) *breaks_host = 0.f;
    return HOIR_OK;
  }
  if (!perm_out) return hoir_set_error(HOIR_EINVAL, "null output pointer", where);
  const size_t n_head = 64 + 1024, n_hist = (size_t)sp.bins_pad * 4, n_kr = (size_t)B * 8;
  void* ws = nullptr;
  if (int err = hoir_get_workspace(3, n_head + n_hist + n_kr, &ws)) return err;   // (zeroed when allocated)
  SortArgs a;
  memset(&a, 0, sizeof(a));
  a.sp = sp;
  a.B = B;
  a.x = x;
  if (mesh) a.mesh = *mesh;
  a.idx = idx;
  a.first = mesh ? mesh->first_pixel : 0;
  a.sync = static_cast<unsigned int*>(ws);
  unsigned long long* brk = reinterpret_cast<unsigned long long*>(static_cast<char*>(ws) + 16);
  a.partial = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + 64);
  a.hist = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + n_head);
  a.keyrank = reinterpret_cast<uint2*>(static_cast<char*>(ws) + n_head + n_hist);
  a.out = perm_out;
  const hoir_mesh_desc& md = a.mesh;
  int occ = 1;
  HOIR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batch_sort_kernel, SORT_THREADS, 0));
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  const long long cap = (long long)hoir_num_sms() * occ;
  long long g = (B + 2 * SORT_THREADS - 1) / (2 * SORT_THREADS);
  if (g < hoir_num_sms()) g = hoir_num_sms();
  const int grid = (int)(g < cap ? g : cap);
  {
    void* args[] = {&a};
    HOIR_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)batch_sort_kernel, dim3(grid), dim3(SORT_THREADS), args, 0, st));
  }
  HOIR_CHECK_CUDA(cudaGetLastError());
  if (breaks_host != nullptr) {
    HOIR_CHECK_CUDA(cudaMemsetAsync(brk, 0, sizeof(unsigned long long), st));
    sort_breaks_kernel<<<grid, SORT_THREADS, 0, st>>>(sp, B, x, md, perm_out, brk);
    unsigned long long h = 0;
    HOIR_CHECK_CUDA(cudaMemcpyAsync(&h, brk, sizeof(h), cudaMemcpyDeviceToHost, st));
    HOIR_CHECK_CUDA(cudaStreamSynchronize(st));
    *breaks_host = (float)((double)h / ((double)B * sp.in_f));
  }
  return HOIR_OK;
}

struct FusedEntry {
  int kind, n, in, hid, out, nh, segh, sego;
  void (*fwd)(const FusedParams);
  void (*train)(const FusedParams);
  size_t smem_fwd, smem_train;
  void (*train_tc)(const FusedParams);   // tensor-core (tcgen05) training kernel, or nullptr
  size_t smem_tc;
  void (*fwd_tc)(const FusedParams);     // tensor-core forward kernel (2 CTAs per SM), or nullptr
  size_t smem_fwd_tc;
  void (*train_tc4)(const FusedParams);  // tensor-core training kernel with four threads per sample, or nullptr
  size_t smem_tc4;
  int threads_tc, threads_tc4;
};

template <class S>
constexpr FusedEntry make_entry() {
  return FusedEntry{S::KIND, S::N, S::IN, S::HID, S::OUT, S::NH, S::SEGH, S::SEGO,
                    fused_mlp_kernel<S, false>, fused_mlp_kernel<S, true>,
                    (size_t)S::S_FWD * 4, (size_t)S::S_TRAIN * 4, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0};
}
// Polynomial networks (the default layer_type of config/images_config.yaml): one window per input, i.e. the
// continuous shape with a single segment everywhere, with the basis evaluated at x itself (FusedParams::identity)
template <class S>
constexpr FusedEntry make_entry_poly() {
  static_assert(S::CONT && S::SEGH == 1 && S::SEGO == 1, "polynomial entry: continuous shape with one segment");
  FusedEntry e = make_entry<S>();
  e.kind = HOIR_POLYNOMIAL;
  return e;
}
template <class S>
constexpr FusedEntry make_entry_tc() {
  return FusedEntry{S::KIND, S::N, S::IN, S::HID, S::OUT, S::NH, S::SEGH, S::SEGO,
                    fused_mlp_kernel<S, false>, fused_mlp_kernel<S, true>,
                    (size_t)S::S_FWD * 4, (size_t)S::S_TRAIN * 4,
                    tc::fused_mlp_tc_kernel<S, 2>, (size_t)tc::TcPlan<S, 2>::SMEM_BYTES,
                    tc::fused_mlp_tc_fwd_kernel<S>, (size_t)tc::TcFwdPlan<S>::SMEM_BYTES,
                    nullptr, 0, tc::TcPlan<S, 2>::NT, 0};
}
template <class S>
constexpr FusedEntry make_entry_tc4() {   // continuous: also the four-threads-per-sample variant
  return FusedEntry{S::KIND, S::N, S::IN, S::HID, S::OUT, S::NH, S::SEGH, S::SEGO,
                    fused_mlp_kernel<S, false>, fused_mlp_kernel<S, true>,
                    (size_t)S::S_FWD * 4, (size_t)S::S_TRAIN * 4,
                    tc::fused_mlp_tc_kernel<S, 2>, (size_t)tc::TcPlan<S, 2>::SMEM_BYTES,
                    tc::fused_mlp_tc_fwd_kernel<S>, (size_t)tc::TcFwdPlan<S>::SMEM_BYTES,
                    tc::fused_mlp_tc_kernel<S, 4>, (size_t)tc::TcPlan<S, 4>::SMEM_BYTES,
                    tc::TcPlan<S, 2>::NT, tc::TcPlan<S, 4>::NT};
}

// Instantiations: BASELINE.json configs C2/C3 (40-wide, 1 hidden layer, rotations 1 and 2) and
// C1 (10-wide, 2 hidden layers), continuous and discontinuous, n = 3; plus n = 2 variants.
const FusedEntry kEntries[] = {
    make_entry_tc4<Shape<HOIR_CONTINUOUS, 3, 4, 40, 3, 1, 2, 2>>(),
    make_entry_tc<Shape<HOIR_DISCONTINUOUS, 3, 4, 40, 3, 1, 2, 2>>(),
    make_entry_tc<Shape<HOIR_DISCONTINUOUS, 3, 2, 40, 3, 1, 2, 2>>(),
    make_entry_tc<Shape<HOIR_CONTINUOUS, 3, 2, 40, 3, 1, 2, 2>>(),
    make_entry<Shape<HOIR_CONTINUOUS, 3, 2, 10, 3, 2, 2, 2>>(),
    make_entry<Shape<HOIR_DISCONTINUOUS, 3, 2, 10, 3, 2, 2, 2>>(),
    make_entry<Shape<HOIR_CONTINUOUS, 3, 4, 10, 3, 2, 2, 2>>(),      // README.md:24 (10-wide with the default rotations=2)
    make_entry<Shape<HOIR_DISCONTINUOUS, 3, 4, 10, 3, 2, 2, 2>>(),
    make_entry<Shape<HOIR_CONTINUOUS, 2, 2, 10, 3, 2, 2, 2>>(),
    make_entry<Shape<HOIR_DISCONTINUOUS, 2, 2, 10, 3, 2, 2, 2>>(),
    make_entry_poly<Shape<HOIR_CONTINUOUS, 3, 4, 10, 3, 2, 1, 1>>(),   // images_config.yaml defaults (rotations 2)
    make_entry_poly<Shape<HOIR_CONTINUOUS, 3, 2, 10, 3, 2, 1, 1>>(),
};

const FusedEntry* find_entry(const hoir_mlp_desc* m) {
  if (!mlp_desc_ok(m) || m->num_layers < 3) return nullptr;
  const int L = m->num_layers;
  const hoir_layer_desc& l0 = m->layers[0];
  const hoir_layer_desc& lo = m->layers[L - 1];
  for (int l = 0; l < L; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    if (d.kind != l0.kind || d.n != l0.n || d.length != 2.f || d.periodicity > 0.f) return nullptr;
    if (d.rescale != 0.f && d.rescale != 1.f) return nullptr;
    if (l > 0 && l < L - 1 && (d.segments != m->layers[1].segments || d.out_features != l0.out_features))
      return nullptr;
  }
  const bool poly = l0.kind == HOIR_POLYNOMIAL;               // `segments` is meaningless for a Polynomial layer
  if (l0.kind != HOIR_CONTINUOUS && l0.kind != HOIR_DISCONTINUOUS && !poly) return nullptr;
  for (const FusedEntry& e : kEntries)
    if (e.kind == l0.kind && e.n == l0.n && e.in == l0.in_features && e.hid == l0.out_features &&
        e.out == lo.out_features && e.nh == L - 2 && e.segh == (poly ? 1 : m->layers[1].segments) &&
        e.sego == (poly ? 1 : lo.segments))
      return &e;
  return nullptr;
}

// floats of the alternative layer-0 copy that follows the per-layer regions (0 without a tensor-core training kernel)
long long alt_floats(const hoir_mlp_desc* m) {
  const FusedEntry* e = find_entry(m);
  if (e == nullptr || e->train_tc == nullptr || (m->layers[0].out_features & 7) != 0) return 0;
  return (long long)m->layers[0].in_features * hoir_layer_nodes(&m->layers[0]) * m->layers[0].out_features;
}

int fill_params(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                const float* packed, FusedParams* p, const char* where) {
  memset(p, 0, sizeof(*p));
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", where);
  if ((x == nullptr) == (mesh == nullptr) && B > 0)
    return hoir_set_error(HOIR_EINVAL, "exactly one of x / mesh must be given", where);
  if (!packed && B > 0) return hoir_set_error(HOIR_EINVAL, "null packed weights", where);
  p->B = B;
  p->x = x;
  if (mesh) {
    if (mesh->rotations * 2 != m->layers[0].in_features || mesh->rows < 1 || mesh->cols < 1 ||
        mesh->normalize_mode < 0 || mesh->normalize_mode > 2 || mesh->first_pixel < 0 ||
        mesh->first_pixel + B > (long long)mesh->rows * mesh->cols)
      return hoir_set_error(HOIR_EINVAL, "mesh descriptor does not match the network / batch", where);
    p->mesh = *mesh;
  }
  p->packed = packed;
  for (int l = 0; l < m->num_layers; ++l) p->off[l] = hoir_mlp_packed_offset(m, l);
  p->alt_off = alt_floats(m) > 0 ? hoir_mlp_packed_offset(m, m->num_layers) : -1;
  p->identity = m->layers[0].kind == HOIR_POLYNOMIAL;
  p->seg0 = p->identity ? 1 : m->layers[0].segments;
  p->nodes0 = hoir_layer_nodes(&m->layers[0]);
  p->pow2_0 = hoir::is_pow2(p->seg0);
  p->normalize = m->normalize;
  p->norm_eps = m->norm_eps;
  hoir_make_lagrange_table(m->layers[0].n, 1.f, &p->lag);
  return HOIR_OK;
}

int num_sms_fused() { return hoir_num_sms(); }

}  // namespace

extern "C" int hoir_mlp_fused_supported(const hoir_mlp_desc* m) {
  return find_entry(m) != nullptr || (mlp_desc_ok(m) && hoir_fourier_supported(m));
}
// 2: the whole training step is fused (hoir_mlp_fused_supported); 1: only the forward (hoir_mlp_forward / hoir_interp_frame:
// the neighborhood-interpolator family, fused_interp.cu -- it trains through the per-layer kernels); 0: neither
extern "C" int hoir_mlp_forward_supported(const hoir_mlp_desc* m) {
  if (hoir_mlp_fused_supported(m)) return 2;
  return mlp_desc_ok(m) && hoir_interp_supported(m) ? 1 : 0;
}
namespace {
bool interp_only(const hoir_mlp_desc* m) { return mlp_desc_ok(m) && find_entry(m) == nullptr && !hoir_fourier_supported(m) && hoir_interp_supported(m); }
}  // namespace

extern "C" long long hoir_mlp_packed_offset(const hoir_mlp_desc* m, int layer) {
  if (!mlp_desc_ok(m) || layer < 0 || layer > m->num_layers) return -1;
  if (hoir_fourier_supported(m)) return hoir_fourier_packed_offset(m, layer);
  if (interp_only(m)) return hoir_interp_packed_offset(m, layer);
  long long off = 0;
  for (int l = 0; l < layer; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    off += (long long)d.in_features * hoir_layer_nodes(&d) * ((d.out_features + 3) & ~3);
  }
  return off;
}

extern "C" long long hoir_mlp_packed_size(const hoir_mlp_desc* m) {
  if (!mlp_desc_ok(m)) return -1;
  return hoir_mlp_packed_offset(m, m->num_layers) + alt_floats(m);
}

extern "C" int hoir_mlp_pack(const hoir_mlp_desc* m, const float* const* w, float* packed,
                             void* stream) {
  if (!mlp_desc_ok(m)) return hoir_set_error(HOIR_EINVAL, "bad network descriptor", "hoir_mlp_pack");
  if (!w || !packed) return hoir_set_error(HOIR_EINVAL, "null pointer", "hoir_mlp_pack");
  if (hoir_fourier_supported(m)) {
    for (int l = 0; l < m->num_layers; ++l)
      if (!w[l]) return hoir_set_error(HOIR_EINVAL, "null weight pointer", "hoir_mlp_pack");
    return hoir_fourier_pack(m, w, packed, (cudaStream_t)stream);
  }
  if (interp_only(m)) {
    for (int l = 0; l < m->num_layers; ++l)
      if (!w[l]) return hoir_set_error(HOIR_EINVAL, "null weight pointer", "hoir_mlp_pack");
    return hoir_interp_pack(m, w, packed, (cudaStream_t)stream);
  }
  for (int l = 0; l < m->num_layers; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    if (!w[l]) return hoir_set_error(HOIR_EINVAL, "null weight pointer", "hoir_mlp_pack");
    const int nodes = hoir_layer_nodes(&d), op = (d.out_features + 3) & ~3;
    const long long total = (long long)d.in_features * nodes * op;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w[l], packed + hoir_mlp_packed_offset(m, l),
                                                        d.in_features, d.out_features, nodes, op);
    if (l == 0 && alt_floats(m) > 0)
      pack_alt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w[0], packed + hoir_mlp_packed_offset(m, m->num_layers),
                                                             d.in_features, d.out_features, nodes);
  }
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_mlp_forward(const hoir_mlp_desc* m, long long B, const float* x,
                                const hoir_mesh_desc* mesh, const float* packed, float* y,
                                int post_scale, void* stream) {
  if (mlp_desc_ok(m) && hoir_fourier_supported(m))
    return hoir_fourier_forward(m, B, x, mesh, packed, y, post_scale, (cudaStream_t)stream);
  if (interp_only(m)) {
    if (mesh != nullptr) return hoir_set_error(HOIR_EUNSUPPORTED, "mesh coordinates: implicit-image networks only", "hoir_mlp_forward");
    return hoir_interp_forward(m, B, x, packed, y, post_scale, nullptr, (cudaStream_t)stream);
  }
  const FusedEntry* e = find_entry(m);
  if (!e) return hoir_set_error(HOIR_EUNSUPPORTED, "no fused instantiation for this network", "hoir_mlp_forward");
  FusedParams p;
  if (int err = fill_params(m, B, x, mesh, packed, &p, "hoir_mlp_forward")) return err;
  if (B == 0) return HOIR_OK;
  if (!y) return hoir_set_error(HOIR_EINVAL, "null output pointer", "hoir_mlp_forward");
  p.y = y;
  p.post_scale = post_scale;
  static const bool no_tc_fwd = getenv("HOIR_NO_TC") != nullptr || getenv("HOIR_NO_TC_FWD") != nullptr;
  if (e->fwd_tc != nullptr && !no_tc_fwd && B >= 4 * tc::TS) {
    HOIR_CHECK_CUDA(cudaFuncSetAttribute(e->fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_fwd_tc));
    const long long tiles_tc = (B + tc::TS - 1) / tc::TS;
    static const bool one_cta = getenv("HOIR_FWD_ONE_CTA") != nullptr;   // A/B: one CTA per SM
    const long long cap = (one_cta ? 1LL : 2LL) * num_sms_fused();
    e->fwd_tc<<<(int)(tiles_tc < cap ? tiles_tc : cap), tc::NTHREADS_FWD, e->smem_fwd_tc, (cudaStream_t)stream>>>(p);
    HOIR_CHECK_CUDA(cudaGetLastError());
    return HOIR_OK;
  }
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(e->fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_fwd));
  const long long tiles = (B + BT - 1) / BT;
  int occ = 1;   // the narrow networks leave room for a second (third) CTA per SM: more warps to hide latency
  HOIR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, e->fwd, BT, e->smem_fwd));
  if (occ < 1 || getenv("HOIR_GRID_ONE_PER_SM") != nullptr) occ = 1;   // the switch is for A/B measurements
  const long long cap = (long long)num_sms_fused() * occ;
  const int grid = (int)(tiles < cap ? tiles : cap);
  e->fwd<<<grid, BT, e->smem_fwd, (cudaStream_t)stream>>>(p);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// Forward of a forward-only-fused network (hoir_mlp_forward_supported == 1) that also stores what the per-layer backward
// kernels need: acts[(2 * (l - 1) + k) * B * hid ...] = input rows of layer l >= 1, k = 0 as the previous layer produced them
// (the input of MaxAbsNormalization), k = 1 normalized (the input of the layer; only written when the network normalizes).
extern "C" int hoir_mlp_forward_acts(const hoir_mlp_desc* m, long long B, const float* x, const float* packed, float* y,
                                     float* acts, void* stream) {
  if (!interp_only(m)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a forward-only-fused network", "hoir_mlp_forward_acts");
  if (B > 0 && !acts) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_mlp_forward_acts");
  return hoir_interp_forward(m, B, x, packed, y, 0, acts, (cudaStream_t)stream);
}

namespace {

__global__ void __launch_bounds__(256) opt_tail_kernel(const OptTail t) {
  if (t.world > 1) hoir::peer_exchange_barrier(t);     // stream order: this GPU's gradient is complete
  hoir::opt_tail_rows(t, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), (int)((gridDim.x * blockDim.x) >> 5),
                      threadIdx.x & 31);
}

long long boundary_offset(const hoir_mlp_desc* m, int layer);
long long boundary_offset_fwd(const hoir_mlp_desc* m) { return boundary_offset(m, m->num_layers); }
long long boundary_offset(const hoir_mlp_desc* m, int layer) {
  long long off = 0;
  for (int l = 0; l < layer; ++l)
    off += (long long)m->layers[l].out_features * m->layers[l].in_features * hoir_layer_nodes(&m->layers[l]);
  return off;
}

long long peer_stride(const hoir_mlp_desc* m) {
  return (boundary_offset_fwd(m) + 1 + 63) / 64 * 64;
}

int fill_tail(const hoir_mlp_desc* m, const hoir_opt_desc* o, OptTail* t, const char* where,
              const hoir_peer_desc* peers = nullptr) {
  memset(t, 0, sizeof(*t));
  if (!mlp_desc_ok(m)) return hoir_set_error(HOIR_EINVAL, "bad network descriptor", where);
  if (!o) return hoir_set_error(HOIR_EINVAL, "null optimizer descriptor", where);
  if (o->kind != HOIR_OPT_ADAM && o->kind != HOIR_OPT_SPARSE_LION)
    return hoir_set_error(HOIR_EINVAL, "optimizer kind not recognized", where);
  if (!o->flat_w || (!o->flat_g && !peers) || !o->exp_avg || !o->packed || (o->kind == HOIR_OPT_ADAM && !o->exp_avg_sq))
    return hoir_set_error(HOIR_EINVAL, "null optimizer state pointer", where);
  if (o->step < 1) return hoir_set_error(HOIR_EINVAL, "step must be >= 1", where);
  t->kind = o->kind;
  t->keep_grad = o->keep_grad;
  t->num_layers = m->num_layers;
  for (int l = 0; l < m->num_layers; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    t->in_f[l] = d.in_features;
    t->nodes[l] = hoir_layer_nodes(&d);
    t->out_pad[l] = (d.out_features + 3) & ~3;
    t->row0[l + 1] = t->row0[l] + d.out_features * d.in_features;
    t->woff[l + 1] = boundary_offset(m, l + 1);
    t->poff[l] = hoir_mlp_packed_offset(m, l);
  }
  t->alt_off = alt_floats(m) > 0 ? hoir_mlp_packed_offset(m, m->num_layers) : -1;
  t->alt_hh = m->layers[0].out_features / 2;
  if (hoir_fourier_supported(m)) hoir_fourier_fill_tail(m, t);
  t->w = o->flat_w;
  t->g = o->flat_g;
  t->m = o->exp_avg;
  t->v = o->exp_avg_sq;
  t->packed = o->packed;
  t->loss_out = o->loss_out;
  t->lr = o->lr;
  t->b1 = o->beta1;
  t->b2 = o->beta2;
  t->eps = o->eps;
  t->wd = o->weight_decay;
  t->bc1 = (float)(1.0 - pow((double)o->beta1, (double)o->step));
  t->bc2_sqrt = (float)sqrt(1.0 - pow((double)o->beta2, (double)o->step));
  if (peers != nullptr) {
    if (peers->world < 2 || peers->world > HOIR_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world)
      return hoir_set_error(HOIR_EINVAL, "bad peer descriptor", where);
    if (o->keep_grad) return hoir_set_error(HOIR_EINVAL, "keep_grad is not available with the peer exchange", where);
    const long long stride = peer_stride(m);
    const int parity = (int)(o->step & 1);
    t->world = peers->world;
    t->rank = peers->rank;
    t->flag_value = (unsigned int)o->step;
    for (int r = 0; r < peers->world; ++r) {
      if (!peers->block[r]) return hoir_set_error(HOIR_EINVAL, "null peer block", where);
      float* base = static_cast<float*>(peers->block[r]);
      t->peer_g[r] = base + parity * stride;
      t->peer_flag[r] = reinterpret_cast<unsigned int*>(base + 2 * stride);
    }
    float* own = static_cast<float*>(peers->block[peers->rank]);
    t->g = own + parity * stride;
    t->zero_buf = own + (parity ^ 1) * stride;
  }
  return HOIR_OK;
}

int launch_tail(const OptTail& t, cudaStream_t stream);
}  // namespace
namespace {
int launch_tail(const OptTail& t, cudaStream_t stream) {
  long long grid = (t.woff[t.num_layers] + 255) / 256;
  if (grid > num_sms_fused() * 8) grid = num_sms_fused() * 8;
  opt_tail_kernel<<<(int)grid, 256, 0, stream>>>(t);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// shared launcher of the training kernels; tail != nullptr: optimizer tail in-kernel when the step is a
// single launch (cooperative), as a separate launch after the second pass otherwise
int train_launch(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                 int target_kind, const void* target, const float* gy_ext, const float* packed,
                 float* const* gw, float* loss_sum, float loss_scale, const OptTail* tail, void* stream,
                 const char* where, const int* perm = nullptr) {
  if (mlp_desc_ok(m) && hoir_fourier_supported(m)) {       // Fourier networks: fused_fourier.cu
    if (perm != nullptr) return hoir_set_error(HOIR_EUNSUPPORTED, "sample permutations: piecewise networks only", where);
    if (B == 0 && !(tail != nullptr && tail->world > 1)) return HOIR_OK;
    return hoir_fourier_train(m, B, x, mesh, target_kind, target, gy_ext, packed, gw, loss_sum, loss_scale, tail,
                              (cudaStream_t)stream, where);
  }
  const FusedEntry* e = find_entry(m);
  if (!e) return hoir_set_error(HOIR_EUNSUPPORTED, "no fused instantiation for this network", where);
  FusedParams p;
  hoir_mesh_desc mesh_all;
  if (perm != nullptr && mesh != nullptr) {   // perm holds absolute pixels: the range check of fill_params does not apply
    mesh_all = *mesh;
    mesh_all.first_pixel = 0;
    if (B > (long long)mesh->rows * mesh->cols) return hoir_set_error(HOIR_EINVAL, "more samples than pixels", where);
    if (int err = fill_params(m, 0, x, &mesh_all, packed, &p, where)) return err;
    p.B = B;
    if (!packed && B > 0) return hoir_set_error(HOIR_EINVAL, "null packed weights", where);
  } else if (int err = fill_params(m, B, x, mesh, packed, &p, where)) return err;
  p.perm = perm;
  if (B == 0) {   // nothing to add; with peers the exchange must still happen (the other ranks wait for it)
    if (tail != nullptr && tail->world > 1) return launch_tail(*tail, (cudaStream_t)stream);
    return HOIR_OK;
  }
  if (!gw) return hoir_set_error(HOIR_EINVAL, "null gradient pointer table", where);
  for (int l = 0; l < m->num_layers; ++l) {
    if (!gw[l]) return hoir_set_error(HOIR_EINVAL, "null gradient pointer", where);
    p.gw[l] = gw[l];
  }
  if (gy_ext == nullptr) {
    if (!target || !loss_sum) return hoir_set_error(HOIR_EINVAL, "target / loss_sum required", where);
    if (target_kind != HOIR_TARGET_F32 && target_kind != HOIR_TARGET_U8_IMAGE)
      return hoir_set_error(HOIR_EINVAL, "unknown target kind", where);
    if (target_kind == HOIR_TARGET_U8_IMAGE && (mesh == nullptr || m->layers[m->num_layers - 1].out_features != 3))
      return hoir_set_error(HOIR_EINVAL, "uint8 image targets need a mesh and 3 outputs", where);
  }
  p.target_kind = target_kind;
  p.target = target;
  p.gy_ext = gy_ext;
  p.loss_sum = loss_sum;
  p.loss_scale = loss_scale;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool no_tc = getenv("HOIR_NO_TC") != nullptr;   // A/B switch: FFMA-only kernel
  static const bool no_2p = getenv("HOIR_NO_TWOPASS") != nullptr;
  static const bool no_fuse = getenv("HOIR_NO_FUSED_TAIL") != nullptr;   // A/B switch: tail as its own launch
  const bool use_tc = e->train_tc != nullptr && !no_tc;
  if (perm != nullptr && !use_tc)
    return hoir_set_error(HOIR_EUNSUPPORTED, "sample permutations need the tensor-core training kernel", where);
  const int in_f = m->layers[0].in_features, hid = m->layers[0].out_features;
  const bool two_pass = use_tc && x != nullptr && perm == nullptr && B >= 32768 && !no_2p && p.seg0 <= GW0_MAXSEG && hid <= 64 &&
                        m->layers[0].n == 3 && gw0_smem_bytes(p.nodes0, hid) <= 200 * 1024;
  if (two_pass) {
    const size_t b_al = ((size_t)B + 63) & ~(size_t)63;
    const size_t n_g0 = b_al * hid * sizeof(float), n_rec = b_al * in_f * sizeof(float4), n_seg = b_al * in_f * sizeof(unsigned short);
    void* ws = nullptr;
    if (int err = hoir_get_workspace(0, n_g0 + n_rec + n_seg, &ws)) return err;
    p.ws_g0 = static_cast<float*>(ws);
    p.ws_rec = reinterpret_cast<float4*>(static_cast<char*>(ws) + n_g0);
    p.ws_seg = reinterpret_cast<unsigned short*>(static_cast<char*>(ws) + n_g0 + n_rec);
  }
  const bool tail_inside = tail != nullptr && !two_pass && !no_fuse;
  if (tail_inside) {
    p.tail = *tail;
    void* ws = nullptr;
    if (int err = hoir_get_workspace(2, 64, &ws)) return err;
    p.tail.sync = static_cast<unsigned int*>(ws);
  }
  static const bool quad = getenv("HOIR_TC_QUAD") != nullptr;   // A/B: four threads per sample
  const bool use_q = use_tc && quad && e->train_tc4 != nullptr;
  auto kern = use_q ? e->train_tc4 : (use_tc ? e->train_tc : e->train);
  const size_t smem = use_q ? e->smem_tc4 : (use_tc ? e->smem_tc : e->smem_train);
  const int block = use_q ? e->threads_tc4 : (use_tc ? e->threads_tc : BT), ts = use_tc ? tc::TS : BT;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (B + ts - 1) / ts;
  int occ = 1;   // persistent grid = every resident CTA slot (the tensor-core kernels fill an SM with one CTA; the
                 // 10-wide FFMA kernels fit two or three); the cooperative launch needs exactly this bound
  HOIR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
  if (occ < 1 || getenv("HOIR_GRID_ONE_PER_SM") != nullptr) occ = 1;   // the switch is for A/B measurements
  const long long cap = (long long)num_sms_fused() * occ;
  int grid = (int)(tiles < cap ? tiles : cap);
  if (tail_inside) {
    // a small batch must not leave the optimizer tail to one or two CTAs (measured, C1 at batch 256: 70 us with the
    // tail on the single tile's CTA, 37 + 4 us as two launches): CTAs beyond the tile count skip the tile loop and
    // join the grid barrier and the walk over the parameters
    static const bool no_extra = getenv("HOIR_NO_TAIL_CTAS") != nullptr;
    long long want = (p.tail.woff[p.tail.num_layers] + 1023) / 1024;
    if (want > cap) want = cap;
    if (!no_extra && grid < want) grid = (int)want;
  }
  if (tail_inside) {
    // the grid barrier needs every CTA resident: cooperative launch (grid <= #SMs, 1 CTA per SM)
    void* args[] = {&p};
    HOIR_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(block), args, smem, st));
  } else {
    kern<<<grid, block, smem, st>>>(p);
  }
  HOIR_CHECK_CUDA(cudaGetLastError());
  if (two_pass) {
    const int stride = hoir::piecewise_stride(m->layers[0].kind, m->layers[0].n);
    // one wave of one CTA per SM (106 registers x 512 threads) when the batch is large enough; chunks of
    // 2048 .. 32768 samples
    // WAVES waves of chunks in sample order (CTA id = chunk * IN + input): the dL/dy rows the resident CTAs gather
    // from are 1/WAVES of the batch, which has to fit in L2 -- every row is read by the IN CTAs of its chunk, in
    // sorted (= random) order; with one wave over a 168 MB buffer the kernel read 1.03 GB from DRAM
    static const int waves_env = getenv("HOIR_GW0_WAVES") ? atoi(getenv("HOIR_GW0_WAVES")) : 0;
    const long long g0_bytes = B * (long long)hid * 4;
    const int waves = waves_env > 0 ? waves_env : (int)((g0_bytes + (96LL << 20) - 1) / (96LL << 20));   // measured at 168 MB: 1 wave 1.242 ms/step, 2: 1.221, 3: 1.225, 4: 1.238, 6: 1.263
    const long long lanes = num_sms_fused() / in_f > 0 ? num_sms_fused() / in_f : 1;
    const long long want = lanes * (waves < 1 ? 1 : waves);
    long long chunk = ((B + want - 1) / want + 255) & ~255LL;
    chunk = chunk < 2048 ? 2048 : (chunk > GW0_CHUNK ? GW0_CHUNK : chunk);
    const dim3 grid2((unsigned)in_f, (unsigned)((B + chunk - 1) / chunk));
    const size_t smem2 = gw0_smem_bytes(p.nodes0, hid, chunk);
    HOIR_CHECK_CUDA(cudaFuncSetAttribute(gw0_slab_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    gw0_slab_kernel<3><<<grid2, GW0_THREADS, smem2, st>>>(B, (int)chunk, in_f, hid, p.seg0, stride, p.nodes0, p.ws_seg, p.ws_rec, p.ws_g0, p.gw[0]);
    HOIR_CHECK_CUDA(cudaGetLastError());
  }
  if (tail != nullptr && !tail_inside) return launch_tail(*tail, st);
  return HOIR_OK;
}

}  // namespace

extern "C" int hoir_mlp_train_step(const hoir_mlp_desc* m, long long B, const float* x,
                                   const hoir_mesh_desc* mesh, int target_kind, const void* target,
                                   const float* gy_ext, const float* packed, float* const* gw,
                                   float* loss_sum, float loss_scale, void* stream) {
  return train_launch(m, B, x, mesh, target_kind, target, gy_ext, packed, gw, loss_sum, loss_scale, nullptr,
                      stream, "hoir_mlp_train_step");
}

extern "C" int hoir_mlp_optimizer_tail(const hoir_mlp_desc* m, const hoir_opt_desc* opt, void* stream) {
  OptTail t;
  if (int err = fill_tail(m, opt, &t, "hoir_mlp_optimizer_tail")) return err;
  return launch_tail(t, (cudaStream_t)stream);
}

extern "C" int hoir_mlp_train_step_fused(const hoir_mlp_desc* m, long long B, const float* x,
                                         const hoir_mesh_desc* mesh, int target_kind, const void* target,
                                         float loss_scale, const hoir_opt_desc* opt, void* stream) {
  OptTail t;
  if (int err = fill_tail(m, opt, &t, "hoir_mlp_train_step_fused")) return err;
  float* gw[HOIR_MAX_LAYERS] = {};
  for (int l = 0; l < m->num_layers; ++l) gw[l] = opt->flat_g + t.woff[l];
  return train_launch(m, B, x, mesh, target_kind, target, nullptr, opt->packed, gw,
                      opt->flat_g + t.woff[m->num_layers], loss_scale, &t, stream, "hoir_mlp_train_step_fused");
}

extern "C" int hoir_mlp_order_supported(const hoir_mlp_desc* m) {
  const FusedEntry* e = find_entry(m);
  SortPlan sp;
  return e != nullptr && e->train_tc != nullptr && getenv("HOIR_NO_TC") == nullptr && make_sort_plan(m->layers[0], &sp);
}

extern "C" int hoir_mlp_batch_sort(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                                   const int* pixels, int* order, float* breaks, void* stream) {
  if (x != nullptr && pixels != nullptr)
    return hoir_set_error(HOIR_EINVAL, "pixel indices need a mesh, not x", "hoir_mlp_batch_sort");
  return batch_sort_launch(m, B, x, mesh, pixels, order, breaks, (cudaStream_t)stream, "hoir_mlp_batch_sort");
}

extern "C" int hoir_mlp_train_step_order(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                                         const int* order, int target_kind, const void* target, const float* gy_ext,
                                         const float* packed, float* const* gw, float* loss_sum, float loss_scale,
                                         void* stream) {
  if (!order && B > 0) return hoir_set_error(HOIR_EINVAL, "null order", "hoir_mlp_train_step_order");
  return train_launch(m, B, x, mesh, target_kind, target, gy_ext, packed, gw, loss_sum, loss_scale, nullptr,
                      stream, "hoir_mlp_train_step_order", order);
}

extern "C" int hoir_mlp_train_step_fused_order(const hoir_mlp_desc* m, long long B, const float* x,
                                               const hoir_mesh_desc* mesh, const int* order, int target_kind,
                                               const void* target, float loss_scale, const hoir_opt_desc* opt,
                                               const hoir_peer_desc* peers, void* stream) {
  const char* where = "hoir_mlp_train_step_fused_order";
  if (!order && B > 0) return hoir_set_error(HOIR_EINVAL, "null order", where);
  OptTail t;
  if (int err = fill_tail(m, opt, &t, where, peers)) return err;
  float* gw[HOIR_MAX_LAYERS] = {};
  float* gbase = peers ? t.g : opt->flat_g;
  for (int l = 0; l < m->num_layers; ++l) gw[l] = gbase + t.woff[l];
  return train_launch(m, B, x, mesh, target_kind, target, nullptr, opt->packed, gw, gbase + t.woff[m->num_layers],
                      loss_scale, &t, stream, where, order);
}

extern "C" long long hoir_peer_stride(const hoir_mlp_desc* m) { return mlp_desc_ok(m) ? peer_stride(m) : -1; }
extern "C" long long hoir_peer_block_floats(const hoir_mlp_desc* m) {
  return mlp_desc_ok(m) ? 2 * peer_stride(m) + 64 : -1;
}

extern "C" int hoir_mlp_train_step_fused_peer(const hoir_mlp_desc* m, long long B, const float* x,
                                              const hoir_mesh_desc* mesh, int target_kind, const void* target,
                                              float loss_scale, const hoir_opt_desc* opt,
                                              const hoir_peer_desc* peers, void* stream) {
  if (!peers) return hoir_set_error(HOIR_EINVAL, "null peer descriptor", "hoir_mlp_train_step_fused_peer");
  OptTail t;
  if (int err = fill_tail(m, opt, &t, "hoir_mlp_train_step_fused_peer", peers)) return err;
  float* gw[HOIR_MAX_LAYERS] = {};
  for (int l = 0; l < m->num_layers; ++l) gw[l] = t.g + t.woff[l];
  return train_launch(m, B, x, mesh, target_kind, target, nullptr, opt->packed, gw, t.g + t.woff[m->num_layers],
                      loss_scale, &t, stream, "hoir_mlp_train_step_fused_peer");
}
