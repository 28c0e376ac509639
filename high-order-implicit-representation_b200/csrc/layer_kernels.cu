// layer_kernels.cu -- generic (any n / segments / widths / kind) single-layer kernels.
//
// These back the drop-in nn.Modules returned by high_order_fc_layers(...) and every network the
// fused whole-network kernel (fused_mlp.cu) has no instantiation for.  One CTA handles a tile of
// TB samples: phase 1 evaluates, once per (sample, input), the node window and the n basis
// values (+ derivatives in backward) into shared memory; phase 2 contracts them with the
// gathered weights w[out][in][window].  Algorithm: SURVEY.md A.3-A.5 / oracle/hol_oracle.py.
#include <stdint.h>
#include "hoir_common.cuh"

// dense_tc.cu: tensor-core path for layers whose densified basis expansion is a dense contraction
const void* hoir_dense_tc_find(const hoir_layer_desc* d, long long B);
int hoir_dense_tc_forward(const void* entry, const hoir_layer_desc* d, long long B, const float* x, const float* w,
                          float* y, int ldo, cudaStream_t stream);
int hoir_dense_tc_backward(const void* entry, const hoir_layer_desc* d, long long B, const float* x, const float* w,
                           const float* gy, int ldo, float* gx, int gx_acc, float* gw, cudaStream_t stream);

// slab_kernels.cu: shared-memory slab kernels for piecewise layers with many segments
bool hoir_slab_eligible(const hoir_layer_desc* d, long long B);
bool hoir_slab_forward_eligible(const hoir_layer_desc* d, long long B);
int hoir_slab_forward(const hoir_layer_desc* d, long long B, const float* x, const float* wp, float* y, int ldo, cudaStream_t stream);
int hoir_slab_grad_w(const hoir_layer_desc* d, long long B, const float* x, const float* gy, int ldo, float* gw, cudaStream_t stream);

namespace {

struct LayerParams {
  int kind, n, segments, in_f, out_f, nodes, stride;
  float length, half, rescale, periodicity, segf, inv_segf;
  int pow2, odd_is_sin;
  LagrangeTable lag;
  FourierTable fou;
};

// record layout per (sample, input): [0]=window base (int bits) [1]=d x_in/dx  [2..2+n) values
// [2+n..2+2n) derivatives wrt x_in (backward only)
__device__ __forceinline__ void eval_record(const LayerParams& p, float x, float* rec, bool with_der) {
  float psign = 1.f;
  if (p.periodicity > 0.f) x = hoir::make_periodic(x, p.periodicity, &psign);
  float* val = rec + 2;
  float* der = with_der ? rec + 2 + p.n : nullptr;
  int base = 0;
  float slope = 1.f;
  if (p.kind == HOIR_CONTINUOUS || p.kind == HOIR_DISCONTINUOUS) {
    int id = hoir::segment_index(x, p.half, p.length, p.segf, p.segments);
    float xi;
    if (p.pow2) {
      xi = hoir::local_coordinate<true>(x, id, p.half, p.length, p.segf, p.inv_segf);
      slope = hoir::local_slope<true>(id, p.length, p.segf);
    } else {
      xi = hoir::local_coordinate<false>(x, id, p.half, p.length, p.segf, p.inv_segf);
      slope = hoir::local_slope<false>(id, p.length, p.segf);
    }
    base = id * p.stride;
    hoir::lagrange_rt(p.n, p.lag.X, p.lag.D, xi, val, der);
  } else if (p.kind == HOIR_POLYNOMIAL) {
    hoir::lagrange_rt(p.n, p.lag.X, p.lag.D, x, val, der);
  } else {
    hoir::fourier_rt(p.n, p.fou.c, x, p.length, p.odd_is_sin, val, der);
  }
  rec[0] = __int_as_float(base);
  rec[1] = slope * psign;
}

// ---- weights re-laid out as [in][node][out_pad4] so that a sample's node window x all outputs is one
//      contiguous run: lanes that walk the outputs read coalesced float4s.
__global__ void layer_pack_kernel(const float* __restrict__ w, float* __restrict__ packed, int in_f,
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

// forward: work item = (sample, 4 outputs)
__global__ void __launch_bounds__(256)
layer_fwd_kernel(LayerParams p, long long B, int TB, const float* __restrict__ x,
                 const float* __restrict__ wp, float* __restrict__ y) {
  extern __shared__ float smem[];
  const int recw = p.n + 2, OP = (p.out_f + 3) & ~3, OPQ = OP / 4;
  for (long long tile = blockIdx.x; tile * TB < B; tile += gridDim.x) {
    const long long b0 = tile * TB;
    const int nb = (int)min((long long)TB, B - b0);
    for (int it = threadIdx.x; it < nb * p.in_f; it += blockDim.x)
      eval_record(p, x[b0 * p.in_f + it], smem + (size_t)it * recw, false);
    __syncthreads();
    for (int it = threadIdx.x; it < nb * OPQ; it += blockDim.x) {
      const int bl = it / OPQ, oq = it - bl * OPQ;
      const float* rec = smem + (size_t)bl * p.in_f * recw;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < p.in_f; ++i, rec += recw) {
        const float4* wr = reinterpret_cast<const float4*>(wp + ((size_t)i * p.nodes + __float_as_int(rec[0])) * OP) + oq;
        for (int j = 0; j < p.n; ++j) {
          const float4 w4 = __ldg(wr + (size_t)j * OPQ);
          const float v = rec[2 + j];
          acc.x = fmaf(v, w4.x, acc.x); acc.y = fmaf(v, w4.y, acc.y);
          acc.z = fmaf(v, w4.z, acc.z); acc.w = fmaf(v, w4.w, acc.w);
        }
      }
      float* yo = y + (b0 + bl) * p.out_f + 4 * oq;
      const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
      for (int c = 0; c < 4; ++c)
        if (4 * oq + c < p.out_f) yo[c] = a4[c] * p.rescale;
    }
    __syncthreads();
  }
}

// backward: dL/dx per (sample, input); dL/dw by owners (input, 4 outputs) that walk the tile's samples,
// accumulate the current node window in registers and flush (global atomics) when the window changes.
template <int NT>
__global__ void __launch_bounds__(256)
layer_bwd_kernel(LayerParams p, long long B, int TB, const float* __restrict__ x,
                 const float* __restrict__ wp, const float* __restrict__ gy,
                 float* __restrict__ gx, float* __restrict__ gw) {
  extern __shared__ float smem[];
  const int recw = 2 * p.n + 2, OP = (p.out_f + 3) & ~3, OPQ = OP / 4;
  for (long long tile = blockIdx.x; tile * TB < B; tile += gridDim.x) {
    const long long b0 = tile * TB;
    const int nb = (int)min((long long)TB, B - b0);
    float* sgy = smem + (((size_t)TB * p.in_f * recw + 3) & ~(size_t)3);   // [TB][OP], zero padded, 16 B aligned
    for (int it = threadIdx.x; it < nb * p.in_f; it += blockDim.x)
      eval_record(p, x[b0 * p.in_f + it], smem + (size_t)it * recw, true);
    for (int it = threadIdx.x; it < nb * OP; it += blockDim.x) {
      const int bl = it / OP, o = it - bl * OP;
      sgy[it] = o < p.out_f ? gy[(b0 + bl) * p.out_f + o] * p.rescale : 0.f;
    }
    __syncthreads();
    if (gx != nullptr) {
      for (int it = threadIdx.x; it < nb * p.in_f; it += blockDim.x) {
        const int bl = it / p.in_f, i = it - bl * p.in_f;
        const float* rec = smem + (size_t)it * recw;
        const float4* g4 = reinterpret_cast<const float4*>(sgy + (size_t)bl * OP);
        const float4* wr = reinterpret_cast<const float4*>(wp + ((size_t)i * p.nodes + __float_as_int(rec[0])) * OP);
        float acc = 0.f;
        for (int j = 0; j < p.n; ++j) {
          float t0 = 0.f, t1 = 0.f;
          for (int q = 0; q < OPQ; ++q) {
            const float4 w4 = __ldg(wr + (size_t)j * OPQ + q);
            const float4 g = g4[q];
            t0 = fmaf(g.x, w4.x, fmaf(g.z, w4.z, t0));
            t1 = fmaf(g.y, w4.y, fmaf(g.w, w4.w, t1));
          }
          acc = fmaf(rec[2 + p.n + j], t0 + t1, acc);
        }
        gx[(b0 + bl) * p.in_f + i] = acc * rec[1];
      }
    }
    if (gw != nullptr) {
      for (int it = threadIdx.x; it < p.in_f * OPQ; it += blockDim.x) {
        const int i = it / OPQ, oq = it - i * OPQ;
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        int cur = -1;
        auto flush = [&]() {
          if (cur < 0) return;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (j < p.n) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int o = 4 * oq + c;
                if (o < p.out_f && acc[j][c] != 0.f) atomicAdd(gw + ((size_t)o * p.in_f + i) * p.nodes + cur + j, acc[j][c]);
                acc[j][c] = 0.f;
              }
            }
          }
        };
        for (int bl = 0; bl < nb; ++bl) {
          const float* rec = smem + ((size_t)bl * p.in_f + i) * recw;
          const int base = __float_as_int(rec[0]);
          if (base != cur) {
            flush();
            cur = base;
          }
          const float4 g = *reinterpret_cast<const float4*>(sgy + (size_t)bl * OP + 4 * oq);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            if (j < p.n) {
              const float v = rec[2 + j];
              acc[j][0] = fmaf(v, g.x, acc[j][0]); acc[j][1] = fmaf(v, g.y, acc[j][1]);
              acc[j][2] = fmaf(v, g.z, acc[j][2]); acc[j][3] = fmaf(v, g.w, acc[j][3]);
            }
          }
        }
        flush();
      }
    }
    __syncthreads();
  }
}

// dL/dw of a layer with MANY node windows (piecewise layers with many segments, e.g. the 216 x 51 x 50
// first layer of the neighborhood interpolator): CTA = (input i, slice of the batch); each warp keeps a
// PRIVATE copy of w[:, i, :]'s gradient in shared memory (lanes <-> outputs, no conflicts, no atomics),
// walks its samples, and the copies are summed and flushed once per CTA.
template <int NT>
__global__ void __launch_bounds__(256)
layer_gw_input_owner_kernel(LayerParams p, long long B, int parts, int nwarps, const float* __restrict__ x,
                            const float* __restrict__ gy, float* __restrict__ gw) {
  extern __shared__ float smem[];
  const int i = blockIdx.x / parts, part = blockIdx.x - i * parts;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slab = p.nodes * p.out_f;
  for (int k = threadIdx.x; k < nwarps * slab; k += blockDim.x) smem[k] = 0.f;
  __syncthreads();
  const long long per = (B + parts - 1) / parts, b_lo = part * per, b_hi = min(B, b_lo + per);
  float* mine = smem + (size_t)warp * slab;
  if (warp < nwarps) {
    for (long long b0 = b_lo + (long long)warp * 32; b0 < b_hi; b0 += (long long)nwarps * 32) {
      // lane-parallel basis evaluation of 32 samples (independent loads), then a broadcast walk
      float rec[2 + NT];
      const long long bl = b0 + lane;
      rec[0] = __int_as_float(0);
#pragma unroll
      for (int j = 0; j < NT; ++j) rec[2 + j] = 0.f;
      if (bl < b_hi) {
        float full[2 + HOIR_MAX_N];
        eval_record(p, __ldg(x + bl * p.in_f + i), full, false);
        rec[0] = full[0];
#pragma unroll
        for (int j = 0; j < NT; ++j)
          if (j < p.n) rec[2 + j] = full[2 + j];
      }
      const int cnt = (int)min((long long)32, b_hi - b0);
      for (int o0 = 0; o0 < p.out_f; o0 += 32) {        // warp-uniform trip count: the shuffles need every lane
        const int o = o0 + lane;
        const bool act = o < p.out_f;
#pragma unroll 4
        for (int sidx = 0; sidx < cnt; ++sidx) {
          const float gv = act ? __ldg(gy + (b0 + sidx) * p.out_f + o) * p.rescale : 0.f;
          const int base = __shfl_sync(0xffffffffu, __float_as_int(rec[0]), sidx);
          float* dst = mine + base * p.out_f + (act ? o : 0);
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const float v = __shfl_sync(0xffffffffu, rec[2 + j], sidx);
            if (act && j < p.n) dst[j * p.out_f] = fmaf(v, gv, dst[j * p.out_f]);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < slab; k += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += smem[(size_t)w * slab + k];
    if (v != 0.f) {
      const int q = k / p.out_f, o = k - q * p.out_f;
      atomicAdd(gw + ((size_t)o * p.in_f + i) * p.nodes + q, v);
    }
  }
}

__global__ void segment_index_kernel(LayerParams p, long long total, const float* __restrict__ x,
                                     int* __restrict__ id_min, int* __restrict__ wid_min,
                                     float* __restrict__ x_local) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    float xv = x[k], ps;
    if (p.periodicity > 0.f) xv = hoir::make_periodic(xv, p.periodicity, &ps);
    int id = hoir::segment_index(xv, p.half, p.length, p.segf, p.segments);
    id_min[k] = id;
    wid_min[k] = id * p.stride;
    if (x_local)
      x_local[k] = p.pow2 ? hoir::local_coordinate<true>(xv, id, p.half, p.length, p.segf, p.inv_segf)
                          : hoir::local_coordinate<false>(xv, id, p.half, p.length, p.segf, p.inv_segf);
  }
}

// MaxAbsNormalization: one warp per row; first index wins ties (torch.max(dim) on CPU).
__global__ void __launch_bounds__(256)
maxabs_kernel(long long B, int width, float eps, const float* __restrict__ x,
              const float* __restrict__ gy, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp0; b < B; b += nwarps) {
    const float* xr = x + b * width;
    float m = -1.f;
    int k = 0x7fffffff;
    for (int i = lane; i < width; i += 32) {
      float a = fabsf(xr[i]);
      if (a > m || (a != a && m == m)) { m = a; k = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m, o);
      int k2 = __shfl_xor_sync(0xffffffffu, k, o);
      if (m2 > m || (m2 == m && k2 < k)) { m = m2; k = k2; }
    }
    const float d = __fadd_rn(m, eps);
    if (gy == nullptr) {
      for (int i = lane; i < width; i += 32) out[b * width + i] = __fdiv_rn(xr[i], d);
    } else {
      const float* gr = gy + b * width;
      float dot = 0.f;
      for (int i = lane; i < width; i += 32) dot = fmaf(gr[i], __fdiv_rn(xr[i], d), dot);
      dot = hoir::warp_sum(dot);
      const float xk = xr[k];
      const float sgn = xk > 0.f ? 1.f : (xk < 0.f ? -1.f : 0.f);
      for (int i = lane; i < width; i += 32) {
        float g = gr[i];
        if (i == k) g -= sgn * dot;
        out[b * width + i] = __fdiv_rn(g, d);
      }
    }
  }
}

__global__ void mesh_positions_kernel(hoir_mesh_desc m, long long B, float* __restrict__ x) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < B;
       k += (long long)gridDim.x * blockDim.x) {
    float c[2 * HOIR_MAX_ROTATIONS];
    hoir::mesh_position(m, m.first_pixel + k, c, HOIR_MAX_ROTATIONS);
    for (int r = 0; r < 2 * m.rotations; ++r) x[k * 2 * m.rotations + r] = c[r];
  }
}

int make_params(const hoir_layer_desc* d, LayerParams* p) {
  if (d == nullptr) return hoir_set_error(HOIR_EINVAL, "null layer descriptor", "make_params");
  if (d->kind < HOIR_CONTINUOUS || d->kind > HOIR_FOURIER)
    return hoir_set_error(HOIR_EINVAL, "layer kind not recognized", "make_params");
  if (d->n < 1 || d->n > HOIR_MAX_N)
    return hoir_set_error(HOIR_EINVAL, "n must be in [1, 32]", "make_params");
  if (d->in_features < 1 || d->out_features < 1)
    return hoir_set_error(HOIR_EINVAL, "in_features/out_features must be >= 1", "make_params");
  const bool piecewise = d->kind == HOIR_CONTINUOUS || d->kind == HOIR_DISCONTINUOUS;
  if (piecewise && d->segments < 1)
    return hoir_set_error(HOIR_EINVAL, "segments must be >= 1", "make_params");
  if (d->kind == HOIR_CONTINUOUS && d->n < 2)
    return hoir_set_error(HOIR_EINVAL, "continuous layers need n >= 2", "make_params");
  if (!(d->length > 0.f))
    return hoir_set_error(HOIR_EINVAL, "length must be > 0", "make_params");
  p->kind = d->kind;
  p->n = d->n;
  p->segments = piecewise ? d->segments : 1;
  p->in_f = d->in_features;
  p->out_f = d->out_features;
  p->nodes = hoir_layer_nodes(d);
  p->stride = piecewise ? hoir::piecewise_stride(d->kind, d->n) : 0;
  p->length = d->length;
  p->half = 0.5f * d->length;
  p->rescale = d->rescale == 0.f ? 1.f : d->rescale;
  p->periodicity = d->periodicity;
  p->segf = (float)p->segments;
  p->inv_segf = 1.f / (float)p->segments;
  p->pow2 = hoir::is_pow2(p->segments) && d->length == 2.f;
  p->odd_is_sin = d->fourier_odd_is_sin;
  // piecewise layers interpolate on the unit nodes; Polynomial on (length/2) * nodes
  hoir_make_lagrange_table(d->n, d->kind == HOIR_POLYNOMIAL ? 0.5f * d->length : 1.f, &p->lag);
  hoir_make_fourier_table(d->n, d->fourier_arg_scale == 0.f ? 2.f : d->fourier_arg_scale, &p->fou);
  return HOIR_OK;
}

int num_sms() { return hoir_num_sms(); }

constexpr int kSmemBudget = 96 * 1024;

}  // namespace

extern "C" int hoir_layer_nodes(const hoir_layer_desc* d) {
  if (d == nullptr) return -1;
  switch (d->kind) {
    case HOIR_CONTINUOUS: return (d->n - 1) * d->segments + 1;
    case HOIR_DISCONTINUOUS: return d->n * d->segments;
    case HOIR_POLYNOMIAL:
    case HOIR_FOURIER: return d->n;
    default: return -1;
  }
}

// packed copy of w in the library-owned workspace (slot 1), refreshed per call (w is small)
static int pack_weights(const LayerParams& p, const float* w, float** packed, cudaStream_t stream) {
  const int OP = (p.out_f + 3) & ~3;
  const size_t n = (size_t)p.in_f * p.nodes * OP;
  void* ws = nullptr;
  if (int e = hoir_get_workspace(1, n * sizeof(float), &ws)) return e;
  int grid = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 8);
  layer_pack_kernel<<<grid, 256, 0, stream>>>(w, static_cast<float*>(ws), p.in_f, p.out_f, p.nodes, OP);
  HOIR_CHECK_CUDA(cudaGetLastError());
  *packed = static_cast<float*>(ws);
  return HOIR_OK;
}

// A layer wider than the 64 outputs the tensor-core / slab kernels hold per CTA is cut into output blocks: the
// boundary layout [out][in][nodes] makes every block of w (and of dL/dw) a contiguous sub-tensor, and y / dL/dy
// are addressed with the full row stride.  Returns the number of blocks (0: do not split) and the block width
// (a multiple of 4, so 16-byte accesses stay aligned when out is).  mode 1: dense blocks, 2: slab blocks.
static int output_blocks(const hoir_layer_desc* d, long long B, bool forward, int* width, int* mode) {
  const int out = d->out_features;
  if (out <= 64 || out > 1024) return 0;
  const int nb = (out + 63) / 64;
  const int ob = ((out + nb - 1) / nb + 3) & ~3;
  if (out - (nb - 1) * ob <= 0) return 0;
  hoir_layer_desc d2 = *d;
  d2.out_features = ob;
  *width = ob;
  if (hoir_dense_tc_find(&d2, B) != nullptr) { *mode = 1; return nb; }
  if (forward ? hoir_slab_forward_eligible(&d2, B) : hoir_slab_eligible(&d2, B)) { *mode = 2; return nb; }
  return 0;
}

extern "C" int hoir_layer_forward(const hoir_layer_desc* d, long long B, const float* x,
                                  const float* w, float* y, void* stream) {
  LayerParams p;
  if (int e = make_params(d, &p)) return e;
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", "hoir_layer_forward");
  if (B == 0) return HOIR_OK;
  if (!x || !w || !y) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_layer_forward");
  if (const void* dense = hoir_dense_tc_find(d, B))
    return hoir_dense_tc_forward(dense, d, B, x, w, y, d->out_features, (cudaStream_t)stream);
  {
    int ob = 0, mode = 0;
    const int nb = output_blocks(d, B, true, &ob, &mode);
    for (int k = 0; k < nb; ++k) {
      hoir_layer_desc d2 = *d;
      const int o0 = k * ob;
      d2.out_features = min(ob, d->out_features - o0);
      const float* wk = w + (size_t)o0 * p.in_f * p.nodes;
      if (mode == 1) {
        if (int e = hoir_dense_tc_forward(hoir_dense_tc_find(&d2, B), &d2, B, x, wk, y + o0, d->out_features, (cudaStream_t)stream)) return e;
      } else {
        LayerParams p2;
        if (int e = make_params(&d2, &p2)) return e;
        float* wp2 = nullptr;
        if (int e = pack_weights(p2, wk, &wp2, (cudaStream_t)stream)) return e;
        if (int e = hoir_slab_forward(&d2, B, x, wp2, y + o0, d->out_features, (cudaStream_t)stream)) return e;
      }
    }
    if (nb > 0) return HOIR_OK;
  }
  const size_t per_sample = (size_t)p.in_f * (p.n + 2) * sizeof(float);
  if (per_sample > (size_t)kSmemBudget)
    return hoir_set_error(HOIR_EUNSUPPORTED, "in_features*(n+2) exceeds the shared-memory tile", "hoir_layer_forward");
  float* wp = nullptr;
  if (int e = pack_weights(p, w, &wp, (cudaStream_t)stream)) return e;
  if (hoir_slab_forward_eligible(d, B)) return hoir_slab_forward(d, B, x, wp, y, d->out_features, (cudaStream_t)stream);
  int TB = (int)min((size_t)64, (size_t)kSmemBudget / per_sample);
  if ((long long)TB > B) TB = (int)B;
  const size_t smem = per_sample * TB;
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(layer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  long long tiles = (B + TB - 1) / TB;
  int grid = (int)min(tiles, (long long)num_sms() * 4);
  layer_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p, B, TB, x, wp, y);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_layer_backward(const hoir_layer_desc* d, long long B, const float* x,
                                   const float* w, const float* gy, float* gx, float* gw,
                                   void* stream) {
  LayerParams p;
  if (int e = make_params(d, &p)) return e;
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", "hoir_layer_backward");
  if (B == 0 || (gx == nullptr && gw == nullptr)) return HOIR_OK;
  if (!x || !w || !gy) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_layer_backward");
  if (const void* dense = hoir_dense_tc_find(d, B))
    return hoir_dense_tc_backward(dense, d, B, x, w, gy, d->out_features, gx, 0, gw, (cudaStream_t)stream);
  bool gw_done = false;
  {
    int ob = 0, mode = 0;
    const int nb = output_blocks(d, B, false, &ob, &mode);
    for (int k = 0; k < nb; ++k) {
      hoir_layer_desc d2 = *d;
      const int o0 = k * ob;
      d2.out_features = min(ob, d->out_features - o0);
      const size_t woff = (size_t)o0 * p.in_f * p.nodes;
      if (mode == 1) {
        if (int e = hoir_dense_tc_backward(hoir_dense_tc_find(&d2, B), &d2, B, x, w + woff, gy + o0, d->out_features, gx, k > 0,
                                           gw ? gw + woff : nullptr, (cudaStream_t)stream)) return e;
      } else if (gw != nullptr) {
        if (int e = hoir_slab_grad_w(&d2, B, x, gy + o0, d->out_features, gw + woff, (cudaStream_t)stream)) return e;
      }
    }
    if (nb > 0 && (mode == 1 || gx == nullptr)) return HOIR_OK;
    gw_done = nb > 0;        // slab blocks: dL/dx (if wanted) still goes through the tile kernel below
  }
  const int OP = (p.out_f + 3) & ~3;
  const size_t per_sample = ((size_t)p.in_f * (2 * p.n + 2) + OP) * sizeof(float);
  if (per_sample + 16 > (size_t)kSmemBudget)
    return hoir_set_error(HOIR_EUNSUPPORTED, "in_features*(2n+2) exceeds the shared-memory tile", "hoir_layer_backward");
  float* wp = nullptr;
  if (gx != nullptr)
    if (int e = pack_weights(p, w, &wp, (cudaStream_t)stream)) return e;
  // many node windows and a big batch: dL/dw through the input-owner kernel (no per-sample atomics)
  const size_t slab = (size_t)p.nodes * p.out_f * sizeof(float);
  const bool piecewise = p.kind == HOIR_CONTINUOUS || p.kind == HOIR_DISCONTINUOUS;
  float* gw_tile = gw_done ? nullptr : gw;
  if (gw_done) {
  } else if (gw != nullptr && hoir_slab_eligible(d, B)) {
    if (int e = hoir_slab_grad_w(d, B, x, gy, d->out_features, gw, (cudaStream_t)stream)) return e;
    gw_tile = nullptr;
    if (gx == nullptr) return HOIR_OK;
  } else if (gw != nullptr && piecewise && p.segments > 8 && B >= 4096 && slab <= 200 * 1024 / 2) {
    int nwarps = (int)min((size_t)8, (size_t)(200 * 1024) / slab);
    int parts = max(1, min(64, (num_sms() * 2 + p.in_f - 1) / p.in_f));
    auto okern = p.n <= 4 ? layer_gw_input_owner_kernel<4> : (p.n <= 8 ? layer_gw_input_owner_kernel<8> : layer_gw_input_owner_kernel<HOIR_MAX_N>);
    HOIR_CHECK_CUDA(cudaFuncSetAttribute(okern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    okern<<<p.in_f * parts, 256, nwarps * slab, (cudaStream_t)stream>>>(p, B, parts, nwarps, x, gy, gw);
    HOIR_CHECK_CUDA(cudaGetLastError());
    gw_tile = nullptr;
    if (gx == nullptr) return HOIR_OK;
  }
  int TB = (int)min((size_t)64, ((size_t)kSmemBudget - 16) / per_sample);
  if ((long long)TB > B) TB = (int)B;
  const size_t smem = per_sample * TB + 16;
  long long tiles = (B + TB - 1) / TB;
  int grid = (int)min(tiles, (long long)num_sms() * 4);
  auto kern = p.n <= 4 ? layer_bwd_kernel<4> : (p.n <= 8 ? layer_bwd_kernel<8> : layer_bwd_kernel<HOIR_MAX_N>);
  HOIR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  kern<<<grid, 256, smem, (cudaStream_t)stream>>>(p, B, TB, x, wp, gy, gx, gw_tile);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_segment_index(const hoir_layer_desc* d, long long B, const float* x,
                                  int* id_min, int* wid_min, float* x_local, void* stream) {
  LayerParams p;
  if (int e = make_params(d, &p)) return e;
  if (d->kind != HOIR_CONTINUOUS && d->kind != HOIR_DISCONTINUOUS)
    return hoir_set_error(HOIR_EINVAL, "segment index is defined for piecewise layers only", "hoir_segment_index");
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", "hoir_segment_index");
  if (B == 0) return HOIR_OK;
  if (!x || !id_min || !wid_min) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_segment_index");
  const long long total = B * p.in_f;
  int grid = (int)min((total + 255) / 256, (long long)num_sms() * 8);
  segment_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, total, x, id_min, wid_min, x_local);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// Thread-per-row variant for the common widths (40, 50, 64, ...): the row lives in registers (vector loads of V
// floats, every byte of every sector used), no shuffles; same arithmetic and tie rule as maxabs_kernel.
template <int V> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <int V> __device__ __forceinline__ void vec_unpack(const typename VecT<V>::type& t, float* v);
template <> __device__ __forceinline__ void vec_unpack<4>(const float4& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
template <> __device__ __forceinline__ void vec_unpack<2>(const float2& t, float* v) { v[0] = t.x; v[1] = t.y; }
template <int V> __device__ __forceinline__ typename VecT<V>::type vec_pack(const float* v);
template <> __device__ __forceinline__ float4 vec_pack<4>(const float* v) { return make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ float2 vec_pack<2>(const float* v) { return make_float2(v[0], v[1]); }

template <int W, int V>
__global__ void __launch_bounds__(128)
maxabs_row_kernel(long long B, float eps, const float* __restrict__ x, const float* __restrict__ gy,
                  float* __restrict__ out) {
  using vec = typename VecT<V>::type;
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  float v[W];
  const vec* xr = reinterpret_cast<const vec*>(x + b * W);
#pragma unroll
  for (int q = 0; q < W / V; ++q) vec_unpack<V>(__ldg(xr + q), v + V * q);
  float m = -1.f;
  int k = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float a = fabsf(v[i]);
    if (a > m || (a != a && m == m)) { m = a; k = i; }
  }
  const float d = __fadd_rn(m, eps);
  vec* orow = reinterpret_cast<vec*>(out + b * W);
  if (gy == nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = __fdiv_rn(v[i], d);
#pragma unroll
    for (int q = 0; q < W / V; ++q) orow[q] = vec_pack<V>(v + V * q);
  } else {
    float g[W];
    const vec* gr = reinterpret_cast<const vec*>(gy + b * W);
#pragma unroll
    for (int q = 0; q < W / V; ++q) vec_unpack<V>(__ldg(gr + q), g + V * q);
    float dot = 0.f, xk = 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      dot = fmaf(g[i], __fdiv_rn(v[i], d), dot);
      xk = (i == k) ? v[i] : xk;
    }
    const float sgn = xk > 0.f ? 1.f : (xk < 0.f ? -1.f : 0.f);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float gi = g[i];
      if (i == k) gi -= sgn * dot;
      g[i] = __fdiv_rn(gi, d);
    }
#pragma unroll
    for (int q = 0; q < W / V; ++q) orow[q] = vec_pack<V>(g + V * q);
  }
}

template <int W, int V>
static bool maxabs_row_launch(long long B, int width, float eps, const float* x, const float* gy, float* out,
                              cudaStream_t stream) {
  if (width != W) return false;
  maxabs_row_kernel<W, V><<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(B, eps, x, gy, out);
  return true;
}

static int maxabs_launch(long long B, int width, float eps, const float* x, const float* gy,
                         float* out, void* stream, const char* where) {
  if (B < 0 || width < 1) return hoir_set_error(HOIR_EINVAL, "bad shape", where);
  if (B == 0) return HOIR_OK;
  if (!x || !out) return hoir_set_error(HOIR_EINVAL, "null device pointer", where);
  // the hidden widths of the reference configs (40, 50, 10) and a few round ones; rows must be vector-aligned
  const uintptr_t ptrs = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | (gy ? reinterpret_cast<uintptr_t>(gy) : 0);
  if (B >= 1024 && (ptrs & 15) == 0) {
    cudaStream_t st = (cudaStream_t)stream;
    if (maxabs_row_launch<40, 4>(B, width, eps, x, gy, out, st) || maxabs_row_launch<50, 2>(B, width, eps, x, gy, out, st) ||
        maxabs_row_launch<64, 4>(B, width, eps, x, gy, out, st) || maxabs_row_launch<32, 4>(B, width, eps, x, gy, out, st) ||
        maxabs_row_launch<20, 4>(B, width, eps, x, gy, out, st) || maxabs_row_launch<10, 2>(B, width, eps, x, gy, out, st)) {
      HOIR_CHECK_CUDA(cudaGetLastError());
      return HOIR_OK;
    }
  }
  int grid = (int)min((B + 7) / 8, (long long)num_sms() * 128);   // one row per warp: latency-bound, so many warps
  maxabs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, width, eps, x, gy, out);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
extern "C" int hoir_maxabs_forward(long long B, int width, float eps, const float* x, float* y,
                                   void* stream) {
  return maxabs_launch(B, width, eps, x, nullptr, y, stream, "hoir_maxabs_forward");
}
extern "C" int hoir_maxabs_backward(long long B, int width, float eps, const float* x,
                                    const float* gy, float* gx, void* stream) {
  if (B > 0 && !gy) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_maxabs_backward");
  return maxabs_launch(B, width, eps, x, gy, gx, stream, "hoir_maxabs_backward");
}

extern "C" int hoir_mesh_positions(const hoir_mesh_desc* mesh, long long B, float* x, void* stream) {
  if (!mesh || mesh->rotations < 1 || mesh->rotations > HOIR_MAX_ROTATIONS || mesh->rows < 1 ||
      mesh->cols < 1 || mesh->normalize_mode < 0 || mesh->normalize_mode > 2)
    return hoir_set_error(HOIR_EINVAL, "bad mesh descriptor", "hoir_mesh_positions");
  if (B < 0) return hoir_set_error(HOIR_EINVAL, "negative batch", "hoir_mesh_positions");
  if (B == 0) return HOIR_OK;
  if (!x) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_mesh_positions");
  int grid = (int)min((B + 255) / 256, (long long)num_sms() * 8);
  mesh_positions_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*mesh, B, x);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
