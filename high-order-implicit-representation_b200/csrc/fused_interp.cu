// fused_interp.cu -- whole-network FORWARD kernel for the neighborhood interpolator family
//   in0 -> hid -> ... -> hid -> out,  piecewise-Lagrange layers (continuous / discontinuous, n = 2 or 3), an input layer
//   with any number of segments and 1- or 2-segment hidden / output layers with MaxAbsNormalization in between -- Net(cfg) as built by
//   /root/reference/examples/implicit_neighborhood.py:37-41 from /root/reference/config/neighborhood_config.yaml:25-31
//   (216 -> 50 -> 50 -> 50 -> 27 for width = outside = 3).
//
// One launch = the inference pass of /root/reference/high_order_implicit_representation/rendering.py:28-95:
// ReflectionPad2d + stride-w donut patches + model forward + F.fold + crop + alpha blend (generate_sequence, :98-126),
// or the plain forward y = MLP(x) of an explicit batch.  Per 256-sample tile (16 compute warps + MMA warp + loader warp):
//
//   layer 0   a gather (2.2 MB of weights for 216 x 51 x 50): the packed table [in][node][out] streams through a 3-stage
//             shared-memory ring (cp.async.bulk + mbarrier tx count), one or two inputs per stage.  Warps w and w + 8
//             share a group of 32 samples (16 each).  Lanes = (sample, input of the stage) evaluate segment index, local
//             coordinate and basis values once and leave a 16-byte record in shared memory; for the contraction a
//             quarter-warp owns one sample (8 lanes x 2 x 4 outputs), so one LDS.128 fetches 128 contiguous bytes of
//             the node rows of FOUR samples, conflict-free; the partial sums stay in registers across all inputs.
//             The layer is bound by those shared-memory reads: 200 bytes per (sample, input, node) for 50 FMAs.
//             Patch mode: the input value of (sample, feature) is read straight from the image (reflection and
//             donut order computed in-kernel) -- no [B, 216] feature tensor in HBM.
//   hidden /  the tile's layer-0 output goes to shared memory rows; the main warp of each pair (thread = sample) runs
//   output    MaxAbs and every 1- / 2-segment layer as a dense contraction over its densified basis (ND = 2 .. 6 node slots
//   layers    per input, 32 / ND inputs per 32-column chunk) on tcgen05 (kind::tf32, 3 passes A*B + A_lo*B + A*B_lo):
//             basis chunks are written straight into TMEM, the pre-swizzled hi | lo weight-chunk images stream through
//             a ring of their own, accumulators in TMEM, results back to the shared-memory row.  The helper warp of
//             the pair meanwhile gathers the next tile (two mbarriers per pair hand the rows back and forth).
//   output    y[b][out], or the fold: pixel (c, ph*w + kh - (w+o), pw*w + kw - (w+o)) = alpha*prev + (1-alpha)*value.
//
// Activations never touch HBM in inference.  For training (hoir_mlp_forward_acts) the same kernel also stores the inputs of
// the layers 1 .. L-1 for the per-layer backward kernels (gw0_tc.cu / slab_kernels.cu for the input layer, dense_tc.cu).
// Small batches run 1, 2 or 4 of the eight 32-sample groups per CTA, so that more SMs stream the table in parallel.
// The layer-0 loop is kept small on purpose: an earlier version with the four inputs of a group unrolled (7.7 k SASS
// instructions) lost 25 % to instruction-cache misses.
#include <stdlib.h>
#include <string.h>
#include "hoir_common.cuh"
#include "opt_tail.cuh"
#include "tc_common.cuh"

// in-kernel cycle accounting of CTA 0 (printed by two of its warps): compiled in with -DHOIR_INTERP_TIMERS=1
#ifndef HOIR_INTERP_TIMERS
#define HOIR_INTERP_TIMERS 0
#endif
#if HOIR_INTERP_TIMERS
#define ICLK() clock64()
#else
#define ICLK() 0LL
#endif

namespace {
using namespace tc;

constexpr int IT = 256, I_CWARPS = 16, I_MMA_WARP = 16, I_LOAD_WARP = 17, I_THREADS = 576, I_LSTAGES = 3, I_WSTAGES = 4;
constexpr uint32_t I_LRING_BYTES = 77824, I_WSTAGE_BYTES = 16384;     // layer-0 slab ring (3 stages); weight-chunk ring (4 x 16 KB)
constexpr int I_PITCH = 65, I_MAXW = 64, I_MAXTC = 4;        // activation row pitch; widest row; tensor-core layers
constexpr int I_A_COL = 0, I_D_COL = 256;
enum { IB_LFULL0 = 0, IB_LEMPTY0 = IB_LFULL0 + I_LSTAGES, IB_WFULL0 = IB_LEMPTY0 + I_LSTAGES, IB_WEMPTY0 = IB_WFULL0 + I_WSTAGES,
       IB_AFULL0 = IB_WEMPTY0 + I_WSTAGES, IB_AFULL1, IB_AEMPTY0, IB_AEMPTY1, IB_DFULL, IB_DEMPTY,
       IB_ROWS_FULL0, IB_ROWS_FREE0 = IB_ROWS_FULL0 + 8, IB_COUNT = IB_ROWS_FREE0 + 8 };   // per (main, helper) pair: rows written / rows read
constexpr int I_MAXIN = 1024;                                          // features of a patch (donut-cell table)
constexpr uint32_t IS_LRING = 0, IS_WRING = I_LRING_BYTES, IS_ACT = IS_WRING + I_WSTAGES * I_WSTAGE_BYTES,
                   IS_REC = IS_ACT + IT * I_PITCH * 4, IS_TAB = IS_REC + I_CWARPS * 64 * 16, IS_BAR = IS_TAB + I_MAXIN * 4,
                   IS_TOTAL = IS_BAR + 512;
static_assert(IS_TOTAL <= 227 * 1024, "shared-memory plan");

struct InterpParams {
  long long B;
  // layer-0 inputs: explicit rows, or patches of an image
  const float* x;                  // [B][in0] or nullptr
  const float* image;              // [C][H][W]
  int C, H, W, width, outside, stride, pad, nW;
  // network
  int in0, hid, out, ntc;          // ntc tensor-core layers after layer 0 (hidden ... + output)
  int seg0, nodes0, stride0, op4;  // layer 0: packed [in0][nodes0][op4]
  float segf0, inv_segf0;
  int pow2_0;
  int gi, ngroups;
  uint32_t stage_bytes0;
  int tc_in[I_MAXTC], tc_out[I_MAXTC], tc_outp[I_MAXTC], tc_nch[I_MAXTC], tc_seg[I_MAXTC];
  const float* tc_img[I_MAXTC];    // [chunk][hi | lo][outp x 32] pre-swizzled chunk images
  int normalize;
  float norm_eps;
  const float* w0;
  int gpc;                         // 32-sample groups per CTA tile (1, 2, 4, 8): small batches spread over more SMs
  // output
  float* acts;                     // training: [tc layer][pre-norm | normalized][B][hid] inputs of the tensor-core layers, or nullptr
  float* y;                        // [B][out]
  int post_scale;
  float* fold_out;                 // [C][H][W] (fold mode) or nullptr
  const float* fold_prev;
  float alpha;
  LagrangeTable lag;
};

__device__ __forceinline__ int reflect(int v, int n) {   // nn.ReflectionPad2d index map
  if (v < 0) v = -v;
  if (v >= n) v = 2 * (n - 1) - v;
  return v;
}
// (channel, dh, dw) of feature f of a patch: channel-major, donut cells in row-major order of the T x T window
// (image_neighborhood_dataset, single_image_dataset.py:126-141)
__device__ __forceinline__ void donut_cell(int f, int width, int outside, int& c, int& dh, int& dw) {
  const int T = width + 2 * outside, E = T * T - width * width;
  c = f / E;
  const int e = f - c * E, top = outside * T;
  if (e < top) {
    dh = e / T; dw = e - dh * T;
  } else if (e < top + width * 2 * outside) {
    const int q = e - top;
    dh = outside + q / (2 * outside);
    const int m = q % (2 * outside);
    dw = m < outside ? m : m + width;
  } else {
    const int q = e - top - width * 2 * outside;
    dh = outside + width + q / T; dw = q % T;
  }
}

// One 32-column chunk of the densified basis of a 1- or 2-segment layer for this thread's sample: ND node slots per input,
// 32 / ND inputs per chunk (inputs c * IPC ...), zero elsewhere.  Segment index and local coordinate with the roundings of
// the generic code (hoir_common.cuh: for 1 or 2 segments on [-1, 1] every scaling is exact).
template <int N, bool CONT, int SEG>
__device__ __forceinline__ void basis_chunk(const float* row, float inv_d, int c, int in_f, const LagrangeTable& lag, float* v) {
  constexpr int STRIDE = CONT ? N - 1 : N, ND = STRIDE * SEG + (CONT ? 1 : 0), IPC = 32 / ND;
#pragma unroll
  for (int k = IPC * ND; k < 32; ++k) v[k] = 0.f;
#pragma unroll
  for (int m = 0; m < IPC; ++m) {
    const int i = c * IPC + m;
    const float xa = i < in_f ? row[i] * inv_d : 0.f;
    const float t = __fadd_rn(xa, 1.f);
    const int id = SEG == 2 && t >= 1.f ? 1 : 0;
    const float xi = __fmaf_rn(2.f, SEG == 2 ? (id ? xa : t) : t * 0.5f, -1.f);
    float lv[N];
    hoir::lagrange<N>(lag.X, lag.D, xi, lv);
    const int base = id * STRIDE;
#pragma unroll
    for (int q = 0; q < ND; ++q) {
      float vq = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) vq = (q == base + j) ? lv[j] : vq;
      v[m * ND + q] = i < in_f ? vq : 0.f;
    }
  }
}

// GPC: 32-sample groups per CTA tile -- 8 (full tiles, the constant lets ptxas keep the schedule of the large-batch case:
// a run-time count cost 9 % there) or 0 = p.gpc (1, 2 or 4: small batches spread over more SMs)
template <int N, bool CONT, int GPC>
__global__ void __launch_bounds__(I_THREADS, 1) interp_fwd_kernel(const InterpParams p) {
  const int gpc = GPC ? GPC : p.gpc;
  extern __shared__ __align__(1024) uint8_t ism[];
  uint8_t* sLRing = ism + IS_LRING;
  uint8_t* sWRing = ism + IS_WRING;
  int* sTab = reinterpret_cast<int*>(ism + IS_TAB);               // patch mode: c << 16 | dh << 8 | dw of every feature
  float* sAct = reinterpret_cast<float*>(ism + IS_ACT);
  float4* sRec = reinterpret_cast<float4*>(ism + IS_REC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ism + IS_BAR);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + IB_COUNT);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == I_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < I_LSTAGES; ++s) { mbar_init(&bars[IB_LFULL0 + s], 1); mbar_init(&bars[IB_LEMPTY0 + s], 2 * gpc); }
    for (int s = 0; s < I_WSTAGES; ++s) { mbar_init(&bars[IB_WFULL0 + s], 1); mbar_init(&bars[IB_WEMPTY0 + s], 1); }
    mbar_init(&bars[IB_AFULL0], 32 * gpc); mbar_init(&bars[IB_AFULL1], 32 * gpc);
    mbar_init(&bars[IB_AEMPTY0], 1); mbar_init(&bars[IB_AEMPTY1], 1);
    mbar_init(&bars[IB_DFULL], 1); mbar_init(&bars[IB_DEMPTY], 32 * gpc);
    for (int g = 0; g < 8; ++g) { mbar_init(&bars[IB_ROWS_FULL0 + g], 1); mbar_init(&bars[IB_ROWS_FREE0 + g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (p.x == nullptr) {
    for (int f = tid; f < p.in0; f += I_THREADS) {
      int c, dh, dw;
      donut_cell(f, p.width, p.outside, c, dh, dw);
      sTab[f] = (c << 16) | (dh << 8) | dw;
    }
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_base_s;
  const int TS = 32 * gpc;                                     // samples per tile (the groups >= gpc of the CTA idle)
  const long long tiles = (p.B + TS - 1) / TS;
  const size_t per_input0 = (size_t)p.nodes0 * p.op4;           // floats

  if (warp == I_LOAD_WARP) {
    // =========================== loader: layer-0 groups, then the weight chunks of the tensor-core layers ========
    if (elect_one()) {
      uint32_t ql = 0, qw = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int g = 0; g < p.ngroups; ++g, ++ql) {
          const uint32_t st = ql % I_LSTAGES, ph = (ql / I_LSTAGES) & 1;
          mbar_wait_backoff(&bars[IB_LEMPTY0 + st], ph ^ 1, 64);
          const int gi = min(p.gi, p.in0 - g * p.gi);
          const uint32_t bytes = (uint32_t)(gi * per_input0 * sizeof(float));
          mbar_expect_tx(&bars[IB_LFULL0 + st], bytes);
          bulk_g2s(sLRing + (size_t)st * p.stage_bytes0, p.w0 + (size_t)g * p.gi * per_input0, bytes, &bars[IB_LFULL0 + st]);
        }
        for (int l = 0; l < p.ntc; ++l) {
          const uint32_t bytes = (uint32_t)p.tc_outp[l] * 256u;
          for (int c = 0; c < p.tc_nch[l]; ++c, ++qw) {
            const uint32_t st = qw % I_WSTAGES, ph = (qw / I_WSTAGES) & 1;
            mbar_wait_backoff(&bars[IB_WEMPTY0 + st], ph ^ 1, 64);
            mbar_expect_tx(&bars[IB_WFULL0 + st], bytes);
            bulk_g2s(sWRing + st * I_WSTAGE_BYTES, p.tc_img[l] + (size_t)c * 2 * p.tc_outp[l] * 32, bytes, &bars[IB_WFULL0 + st]);
          }
        }
      }
    }
  } else if (warp == I_MMA_WARP) {
    // =========================== MMA warp ===================================================================
    const uint32_t leader = elect_one();
    uint32_t qw = 0, qa = 0, qd = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int l = 0; l < p.ntc; ++l) {
        const uint32_t idesc = make_idesc(128, p.tc_outp[l], 0, 0);      // A: TMEM, B: W chunk K-major (rows = outputs)
        if (qd > 0) mbar_wait(&bars[IB_DEMPTY], (qd - 1) & 1);
        const int nch = p.tc_nch[l];
        for (int c = 0; c < nch; ++c, ++qw, ++qa) {
          const uint32_t st = qw % I_WSTAGES, wph = (qw / I_WSTAGES) & 1, buf = qa & 1, ph = (qa >> 1) & 1;
          if (l == 0 && c == 0) mbar_wait_backoff(&bars[IB_AFULL0 + buf], ph, 256);   // the tile's whole layer 0 away
          else mbar_wait(&bars[IB_AFULL0 + buf], ph);
          mbar_wait(&bars[IB_WFULL0 + st], wph);
          fence_after();
          if (leader) {
            const uint32_t wbase = smem_u32(sWRing + st * I_WSTAGE_BYTES);
            const uint64_t dW[2] = {make_desc(wbase, 512), make_desc(wbase + (uint32_t)p.tc_outp[l] * 128u, 512)};
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
              if (blk * 4 < gpc)
#pragma unroll
              for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  mma_ts(tmem + I_D_COL + blk * 64, tmem + I_A_COL + blk * 128 + buf * 64 + (pass == 1 ? 32 : 0) + j * 8,
                         dW[pass == 2] + (uint64_t)((j * 32) >> 4), idesc, (c | pass | j) != 0);
            umma_commit(&bars[IB_AEMPTY0 + buf]);
            umma_commit(&bars[IB_WEMPTY0 + st]);
            if (c == nch - 1) umma_commit(&bars[IB_DFULL]);
          }
          __syncwarp();
        }
        ++qd;
      }
    }
  } else if ((warp & 7) >= gpc) {
    // (a 32-sample group beyond the tile: small batches run fewer groups per CTA on more SMs)
  } else {
    // =========================== compute warps ==============================================================
    // warps w and w + 8 share the 32 samples of group w & 7 in layer 0 (16 each: four warps per scheduler keep the
    // shared-memory pipe busy); the tensor-core layers and the output belong to the main warp (w < 8, thread = sample),
    // while its helper already gathers the next tile
    const int grp = warp & 7, hf = warp >> 3;
    const bool main_w = hf == 0;
    const int blk = grp >> 2;
    const uint32_t tm_lane = tmem + ((uint32_t)((grp & 3) * 32) << 16);
    float4* myrec = sRec + warp * 64;                                // [2 buffers][2 inputs][16 samples]
    const int qtr = lane >> 3, l8 = lane & 7, rowb = p.op4 * 4;       // bytes per node row of the layer-0 table
    const bool wide = p.op4 > 32;                                     // more than 32 outputs: two 128-byte reads per row
    const LagrangeTable& lag = p.lag;
    uint32_t ql = 0, qa = 0, qd = 0;
    const int fg = p.width + p.outside;
    long long tm_all = ICLK(), tm_lwait = 0, tm_l0 = 0, tm_tc = 0, tm_awt = 0, tm_dwt = 0, tm_out = 0, tm_rec = 0, tm_ldx = 0, tm_rows = 0, tm_pre = 0;
    uint32_t nt = 0;                                                 // tiles this CTA has started
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      // ---- layer 0: lanes = (sample, input parity) for the basis records, lanes = outputs for the contraction
      {
        const long long bl = tile * TS + grp * 32 + 16 * hf + (lane & 15);
        const bool live = bl < p.B;
        int y0 = 0, x0 = 0;
        if (p.x == nullptr) {
          const long long bq = live ? bl : 0;
          const int ph = (int)(bq / p.nW), pw = (int)(bq - (long long)ph * p.nW);
          y0 = ph * p.stride - p.pad;
          x0 = pw * p.stride - p.pad;
        }
        // a quarter-warp owns the samples 4k + qtr (k = 0..3) of the warp's 16; its 8 lanes hold outputs 4*l8 .. +3 and
        // 32 + 4*l8 .. +3: one LDS.128 fetches 128 contiguous bytes of FOUR samples' node rows (conflict-free), and the
        // broadcast of a basis record costs half a wavefront per sample
        float acc[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[k][c] = 0.f;
        // Software pipeline over the groups of the ring (gi <= 2 inputs each).  Lane = (sample s = lane & 15 of the warp's
        // 16, input h = lane >> 4 of the group): one pass of the record code serves both inputs.  At the top of group g:
        // the records of group g + 1 are computed from values loaded a group earlier, the values of group g + 2 are
        // requested, their donut-cell table entries (patch mode) one group earlier still.
        const int hsel = lane >> 4;
        auto load_tab = [&](int g) {
          const int i = g * p.gi + hsel;
          return (p.x == nullptr && hsel < p.gi && i < p.in0) ? sTab[i] : 0;
        };
        auto load_x = [&](int g, int t) {
          const int i = g * p.gi + hsel;
          float v = 0.f;
          if (live && hsel < p.gi && i < p.in0) {
            if (p.x != nullptr)
              v = __ldg(p.x + bl * p.in0 + i);
            else
              v = __ldg(p.image + ((size_t)(t >> 16) * p.H + reflect(y0 + ((t >> 8) & 255), p.H)) * p.W +
                        reflect(x0 + (t & 255), p.W));
          }
          return v;
        };
        // basis records of a group: rec[il][s] = (first node row in bytes from the slab, N basis values)
        auto make_records = [&](float xval, float4* dst) {
          if (hsel < p.gi) {
            const int id = hoir::segment_index(xval, 1.f, 2.f, p.segf0, p.seg0);
            const float xi = p.pow2_0 ? hoir::local_coordinate<true>(xval, id, 1.f, 2.f, p.segf0, p.inv_segf0)
                                      : hoir::local_coordinate<false>(xval, id, 1.f, 2.f, p.segf0, p.inv_segf0);
            float l[N];
            hoir::lagrange<N>(lag.X, lag.D, xi, l);
            HOIR_DBG_ASSERT(id >= 0 && (uint32_t)((hsel * p.nodes0 + id * p.stride0 + N) * rowb) <= p.stage_bytes0);
            dst[lane] = make_float4(__int_as_float((hsel * p.nodes0 + id * p.stride0) * rowb), l[0], N > 1 ? l[1] : 0.f,
                                    N > 2 ? l[2] : 0.f);
          }
        };
        const long long cp0 = ICLK();
        int tb = load_tab(0);
        float xc = load_x(0, tb), xn;
        make_records(xc, myrec);
        tb = load_tab(1);
        xc = load_x(1, tb);                      // (groups past the last one: every lane predicated off)
        tb = load_tab(2);
        __syncwarp();
        tm_pre += ICLK() - cp0;
        for (int g = 0; g < p.ngroups; ++g, ++ql) {
          const uint32_t st = ql % I_LSTAGES, ph = (ql / I_LSTAGES) & 1;
          const int gi = min(p.gi, p.in0 - g * p.gi);
          const long long cr0 = ICLK();
          make_records(xc, myrec + ((g + 1) & 1) * 32);
          const long long cr1 = ICLK();
          xn = load_x(g + 2, tb);
          tb = load_tab(g + 3);
          const long long c0 = ICLK();
          tm_rec += cr1 - cr0;
          tm_ldx += c0 - cr1;
          mbar_wait(&bars[IB_LFULL0 + st], ph);
          const long long c1 = ICLK();
          tm_lwait += c1 - c0;
          const uint8_t* slab = sLRing + (size_t)st * p.stage_bytes0 + 16 * l8;
          const float4* cur = myrec + (g & 1) * 32 + qtr;
#pragma unroll 1
          for (int il = 0; il < gi; ++il, cur += 16) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 rc = cur[4 * k];                 // 4 distinct records per instruction (one per quarter-warp)
              const float lj[3] = {rc.y, rc.z, rc.w};
              const uint8_t* row = slab + __float_as_int(rc.x);
#pragma unroll
              for (int j = 0; j < N; ++j) {
                // every lane loads (lanes past the row's last output read the next row / stage: in bounds of the ring,
                // never stored) -- no predicates, no zero fills in the inner loop
                const float4 a = *reinterpret_cast<const float4*>(row + j * rowb);
                acc[k][0] = fmaf(lj[j], a.x, acc[k][0]); acc[k][1] = fmaf(lj[j], a.y, acc[k][1]);
                acc[k][2] = fmaf(lj[j], a.z, acc[k][2]); acc[k][3] = fmaf(lj[j], a.w, acc[k][3]);
                if (wide) {                                   // (outputs 32 .. 63: a second wavefront per sample and row)
                  const float4 c = *reinterpret_cast<const float4*>(row + j * rowb + 128);
                  acc[k][4] = fmaf(lj[j], c.x, acc[k][4]); acc[k][5] = fmaf(lj[j], c.y, acc[k][5]);
                  acc[k][6] = fmaf(lj[j], c.z, acc[k][6]); acc[k][7] = fmaf(lj[j], c.w, acc[k][7]);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[IB_LEMPTY0 + st]);
          tm_l0 += ICLK() - c1;
          xc = xn;
        }
        // the layer-0 output of the warp's 16 samples -> shared memory rows (thread-per-sample form for the
        // tensor-core layers); a helper first waits until its main warp is done with the previous tile's rows
        const long long cw0 = ICLK();
        if (!main_w && nt > 0) mbar_wait(&bars[IB_ROWS_FREE0 + grp], (nt - 1) & 1);
        float* rows = sAct + (size_t)(grp * 32 + 16 * hf) * I_PITCH;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int o = (c < 4 ? 0 : 28) + 4 * l8 + c;
            HOIR_DBG_ASSERT(o >= p.hid || (grp * 32 + 16 * hf + 4 * k + qtr) * I_PITCH + o < IT * I_PITCH);
            if (o < p.hid) rows[(4 * k + qtr) * I_PITCH + o] = acc[k][c];
          }
        ++nt;
        if (!main_w) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[IB_ROWS_FULL0 + grp]);
          continue;                                            // on to the next tile's layer 0
        }
        mbar_wait(&bars[IB_ROWS_FULL0 + grp], (nt - 1) & 1);
        __syncwarp();
        tm_rows += ICLK() - cw0;
      }
      // ---- tensor-core layers (main warps): thread = sample, its activation row stays in shared memory
      const long long ct0 = ICLK();
      const long long b = tile * TS + grp * 32 + lane;
      const bool live = b < p.B;
      float* row = sAct + (size_t)(grp * 32 + lane) * I_PITCH;
      for (int l = 0; l < p.ntc; ++l) {
        const int in_f = p.tc_in[l];
        float inv_d = 1.f;
        if (p.normalize) {                                   // MaxAbsNormalization: x / (max |x| + eps)
          float m = 0.f;
#pragma unroll 10
          for (int i = 0; i < in_f; ++i) m = fmaxf(m, fabsf(row[i]));
          inv_d = __fdiv_rn(1.f, __fadd_rn(m, p.norm_eps));
        }
        if (p.acts != nullptr) {
          // training: the layer's input rows (as they arrive, and as the layer sees them) for the backward kernels;
          // the warp writes its 32 rows cooperatively (lanes = consecutive features)
          __syncwarp();                                      // (every lane's row of the previous layer is written)
          const float* grows = sAct + (size_t)(grp * 32) * I_PITCH;
          const long long bw = tile * TS + grp * 32;
          float* ah = p.acts + ((size_t)(2 * l) * p.B + bw) * in_f;
          float* au = p.acts + ((size_t)(2 * l + 1) * p.B + bw) * in_f;
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const float sc = __shfl_sync(0xffffffffu, inv_d, r);
            if (bw + r < p.B) {
              for (int i = lane; i < in_f; i += 32) {
                const float hv = grows[r * I_PITCH + i];
                ah[(size_t)r * in_f + i] = hv;
                if (p.normalize) au[(size_t)r * in_f + i] = hv * sc;
              }
            }
          }
          __syncwarp();
        }
        const int nch = p.tc_nch[l];
#pragma unroll 1
        for (int c = 0; c < nch; ++c, ++qa) {
          const uint32_t buf = qa & 1, ph = (qa >> 1) & 1;
          float v[32], lo[32];
          if (p.tc_seg[l] == 2) basis_chunk<N, CONT, 2>(row, inv_d, c, in_f, lag, v);
          else basis_chunk<N, CONT, 1>(row, inv_d, c, in_f, lag, v);
#pragma unroll
          for (int k = 0; k < 32; ++k) lo[k] = tf32_lo(v[k]);
          const long long ca0 = ICLK();
          mbar_wait(&bars[IB_AEMPTY0 + buf], ph ^ 1);
          tm_awt += ICLK() - ca0;
          fence_after();
          const uint32_t col = tm_lane + I_A_COL + blk * 128 + buf * 64;
          tmem_st32(col, v);
          tmem_st32(col + 32, lo);
          tmem_wait_st();
          fence_before();
          mbar_arrive(&bars[IB_AFULL0 + buf]);
        }
        const long long cd0 = ICLK();
        mbar_wait(&bars[IB_DFULL], qd & 1);
        tm_dwt += ICLK() - cd0;
        fence_after();
#pragma unroll
        for (int g = 0; g < I_MAXW; g += 16)
          if (g < p.tc_outp[l]) {
            float t16[16];
            tmem_ld16(tm_lane + I_D_COL + blk * 64 + g, t16);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) row[g + j] = t16[j];
          }
        fence_before();
        mbar_arrive(&bars[IB_DEMPTY]);
        ++qd;
      }
      // ---- output
      const long long co0 = ICLK();
      tm_tc += co0 - ct0;
      if (live) {
        if (p.fold_out == nullptr) {
          for (int o = 0; o < p.out; ++o) p.y[b * p.out + o] = p.post_scale ? 0.5f * (row[o] + 1.f) : row[o];
        } else {
          const int ph = (int)(b / p.nW), pw = (int)(b - (long long)ph * p.nW);
          int c = 0, kh = 0, kw = 0;
          for (int o = 0; o < p.out; ++o) {
            const int r = ph * p.width + kh - fg, cc = pw * p.width + kw - fg;
            if (r >= 0 && r < p.H && cc >= 0 && cc < p.W) {
              const size_t k = ((size_t)c * p.H + r) * p.W + cc;
              const float hv = row[o];
              p.fold_out[k] = p.fold_prev ? p.alpha * __ldg(p.fold_prev + k) + (1.f - p.alpha) * hv : hv;
            }
            if (++kw == p.width) { kw = 0; if (++kh == p.width) { kh = 0; ++c; } }
          }
        }
      }
      tm_out += ICLK() - co0;
      // the helper may overwrite the rows of this group now
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[IB_ROWS_FREE0 + grp]);
    }
    if (HOIR_INTERP_TIMERS && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15))
      printf("warp %d all %lld  layer0: pre %lld rec %lld ldx %lld wait %lld work %lld rows %lld  tc %lld (a-wait %lld d-wait %lld)  out %lld\n",
             warp, ICLK() - tm_all, tm_pre, tm_rec, tm_ldx, tm_lwait, tm_l0, tm_rows, tm_tc, tm_awt, tm_dwt, tm_out);
  }
  fence_before();
  __syncthreads();
  if (warp == I_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- packing ---------------------------------------------------------------------------------------------------
__global__ void interp_pack_plain_kernel(const float* __restrict__ w, float* __restrict__ packed, int in_f, int out_f,
                                         int nodes, int op) {
  const long long total = (long long)in_f * nodes * op;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(k % op);
    const long long r = k / op;
    const int q = (int)(r % nodes), i = (int)(r / nodes);
    packed[k] = o < out_f ? w[((size_t)o * in_f + i) * nodes + q] : 0.f;
  }
}
// w[out][in][nodes] -> [chunk][hi | lo][outp rows x 32 columns] (SWIZZLE_128B_BASE32B), nd slots x ipc inputs per chunk
__global__ void interp_pack_image_kernel(const float* __restrict__ w, float* __restrict__ img, int in_f, int out_f,
                                         int nodes, int nd, int ipc, int nch, int outp) {
  const long long total = (long long)nch * outp * 32;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(k & 31);
    const long long r = k >> 5;
    const int o = (int)(r % outp), c = (int)(r / outp);
    const int m = col / nd, q = col - m * nd, i = c * ipc + m;
    float v = 0.f;
    if (m < ipc && i < in_f && q < nodes && o < out_f) v = w[((size_t)o * in_f + i) * nodes + q];
    float* chunk = img + (size_t)c * 2 * outp * 32;
    const uint32_t off = (((uint32_t)o >> 2) * 512u + ((uint32_t)o & 3u) * 128u + (((uint32_t)col << 2) ^ (((uint32_t)o & 3u) << 5))) >> 2;
    chunk[off] = v;
    chunk[(size_t)outp * 32 + off] = tf32_lo(v);
  }
}

struct IGeo {
  int kind, n, L, in0, hid, out, ntc;
  int nodes0, op4, gi, ngroups;
  uint32_t stage_bytes0;
  int tc_in[I_MAXTC], tc_out[I_MAXTC], tc_outp[I_MAXTC], tc_nch[I_MAXTC], tc_seg[I_MAXTC], tc_nd[I_MAXTC], tc_ipc[I_MAXTC];
  long long off[HOIR_MAX_LAYERS + 1];          // packed float offsets per layer, then the total
};

bool igeometry(const hoir_mlp_desc* m, IGeo* g) {
  static const bool off = getenv("HOIR_NO_INTERP_FUSED") != nullptr;
  if (off || m == nullptr || m->num_layers < 3 || m->num_layers > 1 + I_MAXTC) return false;
  const hoir_layer_desc& l0 = m->layers[0];
  if (l0.kind != HOIR_CONTINUOUS && l0.kind != HOIR_DISCONTINUOUS) return false;
  if (l0.n != 2 && l0.n != 3) return false;
  const int L = m->num_layers;
  for (int l = 0; l < L; ++l) {
    const hoir_layer_desc& d = m->layers[l];
    if (d.kind != l0.kind || d.n != l0.n || d.length != 2.f || d.periodicity > 0.f) return false;
    if (d.rescale != 1.f) return false;
    if (l > 0 && d.in_features != m->layers[l - 1].out_features) return false;
    if (l > 0 && d.segments != 1 && d.segments != 2) return false;
    if (l < L - 1 && d.out_features != l0.out_features) return false;
  }
  const int hid = l0.out_features, out = m->layers[L - 1].out_features;
  if (hid < 2 || hid > 64 || out < 1 || out > 32 || l0.in_features < 1 || l0.segments < 1) return false;
  g->kind = l0.kind; g->n = l0.n; g->L = L;
  g->in0 = l0.in_features; g->hid = hid; g->out = out; g->ntc = L - 1;
  const int stride = hoir::piecewise_stride(l0.kind, l0.n);
  g->nodes0 = hoir_layer_nodes(&l0);
  g->op4 = (hid + 3) & ~3;
  const size_t per_input = (size_t)g->nodes0 * g->op4 * sizeof(float);
  int gi = (int)((I_LRING_BYTES / I_LSTAGES) / per_input);
  if (gi < 1 || (per_input % 16) != 0) return false;
  g->gi = gi > 2 ? 2 : gi;                              // lanes = 16 samples x 2 inputs for the basis records
  g->ngroups = (g->in0 + g->gi - 1) / g->gi;
  g->stage_bytes0 = (uint32_t)((g->gi * per_input + 127) / 128 * 128);
  if ((size_t)I_LSTAGES * g->stage_bytes0 > I_LRING_BYTES) return false;
  long long o = 0;
  g->off[0] = 0;
  o += (long long)g->in0 * g->nodes0 * g->op4;
  for (int l = 1; l < L; ++l) {
    const int k = l - 1;
    g->tc_in[k] = m->layers[l].in_features;
    g->tc_out[k] = m->layers[l].out_features;
    g->tc_outp[k] = (m->layers[l].out_features + 15) & ~15;
    g->tc_seg[k] = m->layers[l].segments;
    g->tc_nd[k] = stride * g->tc_seg[k] + (l0.kind == HOIR_CONTINUOUS ? 1 : 0);        // == hoir_layer_nodes
    g->tc_ipc[k] = 32 / g->tc_nd[k];
    g->tc_nch[k] = (m->layers[l].in_features + g->tc_ipc[k] - 1) / g->tc_ipc[k];
    g->off[l] = o;
    o += (long long)g->tc_nch[k] * 2 * g->tc_outp[k] * 32;
  }
  g->off[L] = o;
  return true;
}

int ifill(const hoir_mlp_desc* m, const IGeo& g, const float* packed, InterpParams* p) {
  memset(p, 0, sizeof(*p));
  p->in0 = g.in0; p->hid = g.hid; p->out = g.out; p->ntc = g.ntc;
  p->seg0 = m->layers[0].segments;
  p->nodes0 = g.nodes0;
  p->stride0 = hoir::piecewise_stride(g.kind, g.n);
  p->op4 = g.op4;
  p->segf0 = (float)p->seg0;
  p->inv_segf0 = 1.f / (float)p->seg0;
  p->pow2_0 = hoir::is_pow2(p->seg0);
  p->gi = g.gi; p->ngroups = g.ngroups; p->stage_bytes0 = g.stage_bytes0;
  for (int k = 0; k < g.ntc; ++k) {
    p->tc_in[k] = g.tc_in[k]; p->tc_out[k] = g.tc_out[k]; p->tc_outp[k] = g.tc_outp[k]; p->tc_nch[k] = g.tc_nch[k];
    p->tc_seg[k] = g.tc_seg[k];
    p->tc_img[k] = packed + g.off[k + 1];
  }
  p->normalize = m->normalize;
  p->norm_eps = m->norm_eps;
  p->w0 = packed + g.off[0];
  hoir_make_lagrange_table(g.n, 1.f, &p->lag);
  return HOIR_OK;
}

int ilaunch(const IGeo& g, InterpParams& p, cudaStream_t st) {
  const int sms = hoir_num_sms();
  int gpc = 8;                                                  // halve the tile while that puts more SMs to work
  while (gpc > 1 && (p.B + 16 * gpc - 1) / (16 * gpc) <= sms) gpc >>= 1;
  p.gpc = gpc;
  const bool cont = g.kind == HOIR_CONTINUOUS;
  const void* kern;
  if (gpc == 8)
    kern = g.n == 2 ? (cont ? (const void*)interp_fwd_kernel<2, true, 8> : (const void*)interp_fwd_kernel<2, false, 8>)
                    : (cont ? (const void*)interp_fwd_kernel<3, true, 8> : (const void*)interp_fwd_kernel<3, false, 8>);
  else
    kern = g.n == 2 ? (cont ? (const void*)interp_fwd_kernel<2, true, 0> : (const void*)interp_fwd_kernel<2, false, 0>)
                    : (cont ? (const void*)interp_fwd_kernel<3, true, 0> : (const void*)interp_fwd_kernel<3, false, 0>);
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IS_TOTAL));
  const long long tiles = (p.B + 32 * gpc - 1) / (32 * gpc);
  void* args[] = {&p};
  HOIR_CHECK_CUDA(cudaLaunchKernel(kern, dim3((unsigned)(tiles < sms ? tiles : sms)), dim3(I_THREADS), args, IS_TOTAL, st));
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

}  // namespace

bool hoir_interp_supported(const hoir_mlp_desc* m) {
  IGeo g;
  return igeometry(m, &g);
}
long long hoir_interp_packed_offset(const hoir_mlp_desc* m, int layer) {
  IGeo g;
  if (!igeometry(m, &g) || layer < 0 || layer > g.L) return -1;
  return g.off[layer];
}
int hoir_interp_pack(const hoir_mlp_desc* m, const float* const* w, float* packed, cudaStream_t st) {
  IGeo g;
  if (!igeometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused interpolator network", "hoir_mlp_pack");
  const int sms = hoir_num_sms();
  interp_pack_plain_kernel<<<sms * 4, 256, 0, st>>>(w[0], packed + g.off[0], g.in0, g.hid, g.nodes0, g.op4);
  for (int l = 1; l < g.L; ++l) {
    const int k = l - 1;
    interp_pack_image_kernel<<<sms, 256, 0, st>>>(w[l], packed + g.off[l], g.tc_in[k], g.tc_out[k], hoir_layer_nodes(&m->layers[l]),
                                                  g.tc_nd[k], g.tc_ipc[k], g.tc_nch[k], g.tc_outp[k]);
  }
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
int hoir_interp_forward(const hoir_mlp_desc* m, long long B, const float* x, const float* packed, float* y, int post_scale,
                        float* acts, cudaStream_t st) {
  IGeo g;
  if (!igeometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused interpolator network", "hoir_mlp_forward");
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", "hoir_mlp_forward");
  if (B == 0) return HOIR_OK;
  if (!x || !packed || !y) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_mlp_forward");
  InterpParams p;
  ifill(m, g, packed, &p);
  p.B = B;
  p.x = x;
  p.y = y;
  p.post_scale = post_scale;
  p.acts = acts;
  return ilaunch(g, p, st);
}

// One pass of the interpolator over an image in ONE launch (neighborhood_sample_generator, rendering.py:28-95, and the
// blend of generate_sequence, :119): out = alpha * prev + (1 - alpha) * fold(model(patches(reflection_pad(image)))).
extern "C" int hoir_interp_frame(const hoir_mlp_desc* m, const float* image, int C, int H, int W, int width, int outside,
                                 const float* packed, float alpha, const float* prev_image, float* out_image, void* stream) {
  const char* where = "hoir_interp_frame";
  IGeo g;
  if (!igeometry(m, &g)) return hoir_set_error(HOIR_EUNSUPPORTED, "not a fused interpolator network", where);
  if (C < 1 || H < 1 || W < 1 || width < 1 || outside < 1 || width + 2 * outside > 255 || g.in0 > I_MAXIN)
    return hoir_set_error(HOIR_EINVAL, "bad image / patch geometry", where);
  const int T = width + 2 * outside, pad = 2 * outside + width;
  if (g.in0 != C * (T * T - width * width) || g.out != C * width * width)
    return hoir_set_error(HOIR_EINVAL, "network widths do not match the patch geometry", where);
  if (pad >= H || pad >= W) return hoir_set_error(HOIR_EINVAL, "reflection padding must be smaller than the image", where);
  if (!image || !packed || !out_image) return hoir_set_error(HOIR_EINVAL, "null device pointer", where);
  if (out_image == image) return hoir_set_error(HOIR_EINVAL, "the output image must not alias the input image", where);
  int nH = 0, nW = 0;
  const long long total = hoir_neighborhood_patches(H, W, width, outside, width, pad, &nH, &nW);
  if (total <= 0) return hoir_set_error(HOIR_EINVAL, "bad image / patch geometry", where);
  InterpParams p;
  ifill(m, g, packed, &p);
  p.B = total;
  p.image = image;
  p.C = C; p.H = H; p.W = W; p.width = width; p.outside = outside; p.stride = width; p.pad = pad; p.nW = nW;
  p.fold_out = out_image;
  p.fold_prev = prev_image;
  p.alpha = alpha;
  return ilaunch(g, p, (cudaStream_t)stream);
}
