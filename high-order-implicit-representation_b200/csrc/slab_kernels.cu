// slab_kernels.cu -- piecewise-Lagrange layers with MANY segments (the 100-segment input layer of the
// implicit-image MLPs, the 216 x 50-segment input layer of the neighborhood interpolator,
// /root/reference/high_order_implicit_representation/networks.py:41-55).  Such a layer is a gather:
// per (sample, input) only n of the (n-1)*segments+1 nodes are touched, so it stays on the CUDA cores and
// the job is to keep the weight table next to the FMAs.
//
// dL/dw: CTA = (group of up to 8 inputs, slice of the batch).  The group's gradient  gw[:, i, :]  lives in
// shared memory as a slab [input][node][OP] (outputs contiguous: a warp updates one node row with a single
// conflict-free LDS.64 / FFMA / STS.64, lane <-> outputs 2*lane, 2*lane+1).  Warp w owns input w of the
// group, so no two warps ever touch the same slab row: no atomics, no per-warp private copies.  The batch
// slice is walked in blocks of 64 samples whose dL/dy rows (coalesced, read ONCE per CTA) and basis records
// (segment window + Lagrange values, evaluated lane-parallel) are staged in shared memory, double-buffered.
// The slab is flushed once with fire-and-forget global adds, nodes fastest (coalesced in the boundary layout).
// (The forward of such a layer stays in layer_kernels.cu: thread = (sample, 4 outputs), weights through L1/L2.)
// Algorithm: oracle/hol_oracle.py.
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "hoir_common.cuh"
#include "tc_common.cuh"

namespace {

struct SlabParams {
  long long B, per_slice;
  int kind, n, segments, stride, in_f, out_f, nodes, op, gi;
  int ldo;                  // row stride of gy (an output block of a wider layer: > out_f)
  float segf, inv_segf, rescale;
  int pow2;
  const float* x;
  const float* gy;
  float* gw;
  LagrangeTable lag;
};

template <int N, class PARAMS>
__device__ __forceinline__ void slab_eval(const PARAMS& p, float xv, int& base, float* l) {
  const int id = hoir::segment_index(xv, 1.f, 2.f, p.segf, p.segments);
  const float xi = p.pow2 ? hoir::local_coordinate<true>(xv, id, 1.f, 2.f, p.segf, p.inv_segf)
                          : hoir::local_coordinate<false>(xv, id, 1.f, 2.f, p.segf, p.inv_segf);
  hoir::lagrange<N>(p.lag.X, p.lag.D, xi, l);
  base = id * p.stride;
}

constexpr int SLAB_THREADS = 256;

constexpr int GW_BLK = 32;   // samples per staged block (32: slab + staging fit twice per SM -> 16 warps)

// HALF: a per-input slab so large that fewer than five inputs fit -- two warps share an input, each owning 32 of its
// output columns (one float per lane instead of a float2), so that more than gi of the 8 warps have work
template <bool HALF>
__device__ __forceinline__ float2 slab_ld(const float* q) {
  if (HALF) return make_float2(*q, 0.f);
  return *reinterpret_cast<const float2*>(q);
}
template <bool HALF>
__device__ __forceinline__ void slab_st(float* q, float2 v) {
  if (HALF) *q = v.x;
  else *reinterpret_cast<float2*>(q) = v;
}

template <int N, bool HALF>
__global__ void __launch_bounds__(SLAB_THREADS) slab_gw_kernel(const SlabParams p) {
  extern __shared__ __align__(16) float slab[];
  const int i0 = blockIdx.x * p.gi;
  const int gi = min(p.gi, p.in_f - i0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row_f = p.nodes * p.op;
  // after the slab: two staged blocks of dL/dy rows [GW_BLK][op] and basis records [8][GW_BLK] (base, l0..l2)
  float* sG = slab + (((size_t)p.gi * row_f + 3) & ~(size_t)3);
  float4* sRec = reinterpret_cast<float4*>(sG + ((2 * GW_BLK * p.op + 3) & ~3));
  for (int k = threadIdx.x; k < gi * row_f; k += SLAB_THREADS) slab[k] = 0.f;
  for (int k = threadIdx.x; k < 2 * GW_BLK * p.op; k += SLAB_THREADS) sG[k] = 0.f;
  __syncthreads();
  const long long b_lo = (long long)blockIdx.y * p.per_slice;
  const long long b_hi = min(p.B, b_lo + p.per_slice);
  const int nblk = b_hi > b_lo ? (int)((b_hi - b_lo + GW_BLK - 1) / GW_BLK) : 0;
  const int inp = HALF ? warp >> 1 : warp;                     // the input this warp accumulates for
  const int col = HALF ? (warp & 1) * 32 + lane : 2 * lane;    // its first output column
  const bool act = col < p.out_f;
  const int g_elems = GW_BLK * p.out_f;
  const bool dense_rows = p.ldo == p.out_f;
  // 16-byte cp.async staging: a straight copy of the block (contiguous rows, even out), or row by row when the rows
  // are strided and every row start / length is a multiple of 4 floats
  const bool vec = p.op == p.out_f && (dense_rows || (((p.out_f | p.ldo) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gy) & 15) == 0));
  // dL/dy rows of block k -> sG[k & 1] (asynchronous when vec)
  auto stage_g = [&](int k) {
    const long long b0 = b_lo + (long long)k * GW_BLK;
    float* g = sG + (k & 1) * GW_BLK * p.op;
    const long long avail = (b_hi - b0) * p.out_f;
    const float* src = p.gy + b0 * p.ldo;
    if (vec) {
      for (int e = threadIdx.x * 4; e < g_elems; e += SLAB_THREADS * 4) {
        const float* se = src + e;
        if (!dense_rows) {
          const int sl = e / p.out_f;
          se = src + (long long)sl * p.ldo + (e - sl * p.out_f);
        }
        if (e + 3 < avail) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(g + e)), "l"(se) : "memory");
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) g[e + c] = e + c < avail ? __ldg(se + c) : 0.f;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
      for (int e = threadIdx.x; e < g_elems; e += SLAB_THREADS) {
        const int sl = e / p.out_f, o = e - sl * p.out_f;
        g[sl * p.op + o] = e < avail ? __ldg(src + (long long)sl * p.ldo + o) : 0.f;
      }
    }
  };
  // this warp's input values of block k (samples lane, lane + 32), issued early; evaluated by finish_rec
  float xreg[GW_BLK / 32];
  auto load_x = [&](int k) {
    const long long b0 = b_lo + (long long)k * GW_BLK;
#pragma unroll
    for (int h = 0; h < GW_BLK / 32; ++h) {
      const long long b = b0 + h * 32 + lane;
      xreg[h] = (inp < gi && b < b_hi) ? __ldg(p.x + b * p.in_f + i0 + inp) : 0.f;
    }
  };
  auto finish_rec = [&](int k) {
    const long long b0 = b_lo + (long long)k * GW_BLK;
#pragma unroll
    for (int h = 0; h < GW_BLK / 32; ++h) {
      const long long b = b0 + h * 32 + lane;
      int base = 0;
      float l[N];
      slab_eval<N>(p, xreg[h], base, l);
      const float sc = b < b_hi ? p.rescale : 0.f;      // rows past the slice contribute nothing
      sRec[((k & 1) * (SLAB_THREADS / 32) + warp) * GW_BLK + h * 32 + lane] =
          make_float4(__int_as_float(base), l[0] * sc, N > 1 ? l[1] * sc : 0.f, N > 2 ? l[2] * sc : 0.f);
    }
  };
  if (nblk > 0) {
    stage_g(0);
    load_x(0);
  }
  // warp w owns input i0 + w (HALF: 32 columns of input i0 + w/2): exclusive slab cells, plain read-modify-write
  float* sl = slab + (size_t)inp * row_f + col;
  for (int k = 0; k < nblk; ++k) {
    if (inp < gi) finish_rec(k);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                    // block k staged; everyone is done with block k - 1
    if (k + 1 < nblk) {
      stage_g(k + 1);
      load_x(k + 1);
    }
    if (inp < gi) {
      const float* g = sG + (k & 1) * GW_BLK * p.op + col;
      const float4* rec = sRec + ((k & 1) * (SLAB_THREADS / 32) + warp) * GW_BLK;
      constexpr int NJ = N < 3 ? N : 3;
      // K samples whose node windows do not overlap: all loads, then all FMAs, then all stores -- the K
      // read-modify-write chains overlap instead of running back to back (the compiler cannot reorder them itself:
      // it must assume the rows alias)
      auto rmw = [&](auto kc, const float4* rc, int s0) {
        constexpr int K = decltype(kc)::value;
        if (!act) return;
        float2 a[K][NJ], g2[K];
        float* row[K];
#pragma unroll
        for (int u = 0; u < K; ++u) {
          g2[u] = slab_ld<HALF>(g + (s0 + u) * p.op);
          row[u] = sl + __float_as_int(rc[u].x) * p.op;
#pragma unroll
          for (int j = 0; j < NJ; ++j) a[u][j] = slab_ld<HALF>(row[u] + j * p.op);
        }
#pragma unroll
        for (int u = 0; u < K; ++u) {
          const float l[3] = {rc[u].y, rc[u].z, rc[u].w};
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            a[u][j].x = fmaf(l[j], g2[u].x, a[u][j].x);
            a[u][j].y = fmaf(l[j], g2[u].y, a[u][j].y);
          }
        }
#pragma unroll
        for (int u = 0; u < K; ++u)
#pragma unroll
          for (int j = 0; j < NJ; ++j) slab_st<HALF>(row[u] + j * p.op, a[u][j]);
      };
      auto apart = [&](const float4& a, const float4& b) {   // warp-uniform: the records are broadcast reads
        const int d = __float_as_int(a.x) - __float_as_int(b.x);
        return d >= NJ || d <= -NJ;
      };
#pragma unroll 2
      for (int s = 0; s < GW_BLK; s += 4) {
        const float4 rc[4] = {rec[s], rec[s + 1], rec[s + 2], rec[s + 3]};
        const bool p01 = apart(rc[0], rc[1]), p23 = apart(rc[2], rc[3]);
        if (p01 && p23 && apart(rc[0], rc[2]) && apart(rc[0], rc[3]) && apart(rc[1], rc[2]) && apart(rc[1], rc[3])) {
          rmw(std::integral_constant<int, 4>{}, rc, s);
        } else {
          if (p01) rmw(std::integral_constant<int, 2>{}, rc, s);
          else { rmw(std::integral_constant<int, 1>{}, rc, s); rmw(std::integral_constant<int, 1>{}, rc + 1, s + 1); }
          if (p23) rmw(std::integral_constant<int, 2>{}, rc + 2, s + 2);
          else { rmw(std::integral_constant<int, 1>{}, rc + 2, s + 2); rmw(std::integral_constant<int, 1>{}, rc + 3, s + 3); }
        }
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < p.out_f * gi * p.nodes; k += SLAB_THREADS) {
    const int q = k % p.nodes, r = k / p.nodes, il = r % gi, o = r / gi;
    const float v = slab[(il * p.nodes + q) * p.op + o];
    if (v != 0.f) hoir::red_add(p.gw + ((size_t)o * p.in_f + i0 + il) * p.nodes + q, v);
  }
}

// ---- forward of a many-segment layer: weight slabs streamed through shared memory ------------------------
// CTA = tile of 512 samples (16 compute warps x 32 samples) + 1 loader warp, persistent grid.  The packed
// weights [in][node][OP4] are cut into groups of GI inputs; the loader streams group after group into a
// 3-stage shared-memory ring with cp.async.bulk (+ mbarrier transaction count), every compute warp consumes
// each group for its 32 samples: basis records evaluated lane-parallel (one sample per lane), broadcast
// through a per-warp smem slot, node rows read with conflict-free LDS.64 (lane <-> outputs 2*lane, 2*lane+1),
// the 32 x 2 partial sums live in registers across all groups; y is written once, coalesced.  Per tile the
// whole weight table crosses L2 -> smem once (2.2 MB per 512 samples for the interpolator's input layer).
constexpr int FW_WARPS = 16, FW_TILE = FW_WARPS * 32, FW_THREADS = (FW_WARPS + 1) * 32, FW_STAGES = 3;

struct FwdParams {
  long long B;
  int n, segments, stride, in_f, out_f, nodes, op4, gi, ngroups;
  int ldo;              // row stride of y
  float segf, inv_segf, rescale;
  int pow2;
  uint32_t stage_bytes;
  const float* x;
  const float* wp;      // packed [in][node][op4]
  float* y;
  LagrangeTable lag;
};

template <int N>
__global__ void __launch_bounds__(FW_THREADS, 1) slab_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(128) uint8_t fsm[];
  float4* sRec = reinterpret_cast<float4*>(fsm + (size_t)FW_STAGES * p.stage_bytes);     // [FW_WARPS][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRec + FW_WARPS * 32);                    // full[ST], empty[ST]
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int s = 0; s < FW_STAGES; ++s) {
      tc::mbar_init(&bars[s], 1);
      tc::mbar_init(&bars[FW_STAGES + s], FW_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const long long tiles = (p.B + FW_TILE - 1) / FW_TILE;
  const size_t per_input = (size_t)p.nodes * p.op4;          // floats
  if (warp == FW_WARPS) {
    if (tc::elect_one()) {
      uint32_t q = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int g = 0; g < p.ngroups; ++g, ++q) {
          const uint32_t st = q % FW_STAGES, ph = (q / FW_STAGES) & 1;
          tc::mbar_wait_backoff(&bars[FW_STAGES + st], ph ^ 1, 64);
          const int gi = min(p.gi, p.in_f - g * p.gi);
          const uint32_t bytes = (uint32_t)(gi * per_input * sizeof(float));
          tc::mbar_expect_tx(&bars[st], bytes);
          tc::bulk_g2s(fsm + (size_t)st * p.stage_bytes, p.wp + (size_t)g * p.gi * per_input, bytes, &bars[st]);
        }
      }
    }
    return;
  }
  float4* myrec = sRec + warp * 32;
  const bool act = 2 * lane < p.out_f, act1 = 2 * lane + 1 < p.out_f;
  uint32_t q = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long b0 = tile * FW_TILE + warp * 32;
    const long long bl = b0 + lane;
    const bool live = bl < p.B;
    float2 acc[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) acc[s] = make_float2(0.f, 0.f);
    for (int g = 0; g < p.ngroups; ++g, ++q) {
      const uint32_t st = q % FW_STAGES, ph = (q / FW_STAGES) & 1;
      const int i0 = g * p.gi, gi = min(p.gi, p.in_f - i0);
      // this lane's sample: the group's inputs (one 16-byte load when the group is 4 aligned inputs)
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {
        const float* xr = p.x + bl * p.in_f + i0;
        if (gi == 4 && (p.in_f & 3) == 0 && (i0 & 3) == 0) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(xr));
          xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
        } else {
#pragma unroll
          for (int il = 0; il < 4; ++il)
            if (il < gi) xv[il] = __ldg(xr + il);
        }
      }
      tc::mbar_wait(&bars[st], ph);
      const float* slab = reinterpret_cast<const float*>(fsm + (size_t)st * p.stage_bytes) + 2 * lane;
#pragma unroll
      for (int il = 0; il < 4; ++il) {
        if (il < gi) {
          int base = 0;
          float l[N];
          slab_eval<N>(p, xv[il], base, l);
          __syncwarp();
          myrec[lane] = make_float4(__int_as_float((il * p.nodes + base) * p.op4), l[0], N > 1 ? l[1] : 0.f, N > 2 ? l[2] : 0.f);
          __syncwarp();
          if (act) {
#pragma unroll
            for (int s = 0; s < 32; ++s) {
              const float4 rc = myrec[s];
              const float lj[3] = {rc.y, rc.z, rc.w};
              const float* row = slab + __float_as_int(rc.x);
#pragma unroll
              for (int j = 0; j < N; ++j) {
                const float2 w2 = *reinterpret_cast<const float2*>(row + j * p.op4);
                acc[s].x = fmaf(lj[j], w2.x, acc[s].x);
                acc[s].y = fmaf(lj[j], w2.y, acc[s].y);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[FW_STAGES + st]);
    }
    if (act) {
#pragma unroll
      for (int s = 0; s < 32; ++s) {
        if (b0 + s < p.B) {
          float* yo = p.y + (b0 + s) * p.ldo + 2 * lane;
          if (act1 && ((p.out_f | p.ldo) & 1) == 0) {
            *reinterpret_cast<float2*>(yo) = make_float2(acc[s].x * p.rescale, acc[s].y * p.rescale);
          } else {
            yo[0] = acc[s].x * p.rescale;
            if (act1) yo[1] = acc[s].y * p.rescale;
          }
        }
      }
    }
  }
}

int slab_sms() { return hoir_num_sms(); }

constexpr size_t kSlabBudget = 88 * 1024;    // + staging: two CTAs per SM

int slab_fill(const hoir_layer_desc* d, long long B, SlabParams* p, int max_gi) {
  memset(p, 0, sizeof(*p));
  p->B = B;
  p->kind = d->kind;
  p->n = d->n;
  p->segments = d->segments;
  p->stride = hoir::piecewise_stride(d->kind, d->n);
  p->in_f = d->in_features;
  p->out_f = d->out_features;
  p->nodes = hoir_layer_nodes(d);
  p->op = (d->out_features + 1) & ~1;
  p->segf = (float)d->segments;
  p->inv_segf = 1.f / (float)d->segments;
  p->rescale = d->rescale == 0.f ? 1.f : d->rescale;
  p->pow2 = hoir::is_pow2(d->segments);
  hoir_make_lagrange_table(d->n, 1.f, &p->lag);
  const size_t per_input = (size_t)p->nodes * p->op * sizeof(float);
  int gi = (int)(kSlabBudget / per_input);
  if (gi > max_gi) gi = max_gi;
  if (gi > p->in_f) gi = p->in_f;
  p->gi = gi;
  return gi;
}

template <class K>
int slab_launch(K kern, SlabParams& p, cudaStream_t stream, size_t extra_smem = 0) {
  const int groups = (p.in_f + p.gi - 1) / p.gi;
  long long slices = (2LL * slab_sms()) / groups;          // one wave of two CTAs per SM
  const long long max_slices = (p.B + 255) / 256;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  p.per_slice = ((p.B + slices - 1) / slices + 31) / 32 * 32;
  slices = (p.B + p.per_slice - 1) / p.per_slice;
  const size_t smem = (size_t)p.gi * p.nodes * p.op * sizeof(float) + extra_smem;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)groups, (unsigned)slices), SLAB_THREADS, smem, stream>>>(p);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

}  // namespace

// many-segment piecewise layer, big batch: the shared-memory slab kernels apply
bool hoir_slab_eligible(const hoir_layer_desc* d, long long B) {
  static const bool off = getenv("HOIR_NO_SLAB") != nullptr;
  // measured on the interpolator's input layer (216 x 51 x 50): 36 us at batch 256 (the reference's batch size), where the
  // generic tile kernel -- 18-sample tiles on 15 CTAs -- takes 450 us; 70 us at 4096, from where gw0_tc.cu (53 us) takes over
  static const long long min_b = getenv("HOIR_SLAB_MIN_B") ? atoll(getenv("HOIR_SLAB_MIN_B")) : 128;
  if (off || d == nullptr || B < min_b) return false;
  if (d->kind != HOIR_CONTINUOUS && d->kind != HOIR_DISCONTINUOUS) return false;
  if (d->length != 2.f || d->periodicity > 0.f || d->segments <= 8 || d->n < 2 || d->n > 3) return false;
  if (d->out_features > 64) return false;
  const size_t per_input = (size_t)hoir_layer_nodes(d) * ((d->out_features + 1) & ~1) * sizeof(float);
  return per_input <= kSlabBudget;
}

// gw0_tc.cu: the same sum as a tcgen05 contraction (node slots of an input <= 128, <= 64 outputs)
bool hoir_gw0_tc_eligible(const hoir_layer_desc* d, long long B);
int hoir_gw0_tc(const hoir_layer_desc* d, long long B, const float* x, const float* gy, int ldo, float* gw, cudaStream_t stream);

int hoir_slab_grad_w(const hoir_layer_desc* d, long long B, const float* x, const float* gy, int ldo, float* gw,
                     cudaStream_t stream) {
  if (hoir_gw0_tc_eligible(d, B)) return hoir_gw0_tc(d, B, x, gy, ldo, gw, stream);
  SlabParams p;
  slab_fill(d, B, &p, SLAB_THREADS / 32);
  p.x = x;
  p.gy = gy;
  p.ldo = ldo;
  p.gw = gw;
  const size_t extra = 2 * GW_BLK * p.op * sizeof(float) + 2 * (SLAB_THREADS / 32) * GW_BLK * sizeof(float4) + 64;
  // measured (224 -> 52 block, 129 nodes, gi = 3): HALF is 10 % SLOWER than leaving 5 warps idle -- every warp still
  // walks all the samples of a block, so the work per warp does not shrink; kept behind a switch for experiments
  static const bool half_on = getenv("HOIR_SLAB_HALF") != nullptr;
  const bool half = half_on && p.gi <= 4 && p.out_f > 32;
  switch (d->n) {
    case 2: return half ? slab_launch(slab_gw_kernel<2, true>, p, stream, extra) : slab_launch(slab_gw_kernel<2, false>, p, stream, extra);
    default: return half ? slab_launch(slab_gw_kernel<3, true>, p, stream, extra) : slab_launch(slab_gw_kernel<3, false>, p, stream, extra);
  }
}

// forward through the streamed-slab kernel; wp = packed weights [in][node][out_pad4] (layer_pack_kernel)
int hoir_slab_forward(const hoir_layer_desc* d, long long B, const float* x, const float* wp, float* y, int ldo,
                      cudaStream_t stream) {
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.n = d->n;
  p.segments = d->segments;
  p.stride = hoir::piecewise_stride(d->kind, d->n);
  p.in_f = d->in_features;
  p.out_f = d->out_features;
  p.nodes = hoir_layer_nodes(d);
  p.op4 = (d->out_features + 3) & ~3;
  p.segf = (float)d->segments;
  p.inv_segf = 1.f / (float)d->segments;
  p.rescale = d->rescale == 0.f ? 1.f : d->rescale;
  p.pow2 = hoir::is_pow2(d->segments);
  hoir_make_lagrange_table(d->n, 1.f, &p.lag);
  p.x = x;
  p.wp = wp;
  p.y = y;
  p.ldo = ldo;
  const size_t per_input = (size_t)p.nodes * p.op4 * sizeof(float);
  int gi = (int)((48 * 1024) / per_input);
  gi = gi < 1 ? 1 : (gi > 4 ? 4 : gi);
  if (gi == 3) gi = 2;
  p.gi = gi;
  p.ngroups = (p.in_f + gi - 1) / gi;
  p.stage_bytes = (uint32_t)((gi * per_input + 127) / 128 * 128);
  const size_t smem = (size_t)FW_STAGES * p.stage_bytes + FW_WARPS * 32 * sizeof(float4) + 2 * FW_STAGES * sizeof(uint64_t) + 64;
  auto kern = d->n == 2 ? slab_fwd_kernel<2> : slab_fwd_kernel<3>;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (B + FW_TILE - 1) / FW_TILE;
  const int grid = (int)(tiles < slab_sms() ? tiles : slab_sms());
  kern<<<grid, FW_THREADS, smem, stream>>>(p);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

bool hoir_slab_forward_eligible(const hoir_layer_desc* d, long long B) {
  // every 512-sample tile streams the whole packed layer through shared memory: constant cost up to one tile per
  // SM (measured 0.56 ms for 224 x 129 x 52 at both 2^12 and 2^16 samples), so small batches stay on the gather kernel
  if (!hoir_slab_eligible(d, B) || B < 16384) return false;
  const size_t per_input = (size_t)hoir_layer_nodes(d) * ((d->out_features + 3) & ~3) * sizeof(float);
  return per_input <= 64 * 1024 && (per_input % 16) == 0;
}
