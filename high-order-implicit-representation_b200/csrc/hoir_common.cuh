// hoir_common.cuh -- device-side arithmetic shared by every kernel of libhoir_b200.
//
// The integer part of the path (segment index, node window) follows the op order of the
// package restated in oracle/hol_oracle.py with explicit
// round-to-nearest intrinsics so ptxas cannot contract them into FMAs: it must be bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hoir.h"

#define HOIR_CHECK_CUDA(expr)                                                     \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) return hoir_set_error(HOIR_ECUDA, cudaGetErrorString(_e), #expr); \
  } while (0)

int hoir_set_error(int code, const char* msg, const char* where);
// library-owned per-device scratch (slot 0: two-pass layer-0 gradient, slot 1: packed layer weights,
// slot 2: grid-barrier counters of the in-kernel optimizer tail, slot 3: histogram / keys of the batch sort, slot 4: private gradient slabs of the fused Fourier kernel;
// memory is zeroed when (re)allocated)
int hoir_get_workspace(int slot, size_t bytes, void** out);
void hoir_free_workspaces();
// multiprocessor count of the CURRENT device (cached per device index: one process may drive several GPUs)
int hoir_num_sms();

// Lagrange nodes X_j (fp32, exactly the host table of the package) and
// D_j = 1 / prod_{m != j} (X_j - X_m) (computed in double on the host).
struct LagrangeTable {
  float X[HOIR_MAX_N];
  float D[HOIR_MAX_N];
};
void hoir_make_lagrange_table(int n, float half_length, LagrangeTable* t);

struct FourierTable {
  float c[HOIR_MAX_N];  // c[j] = fl32(arg_scale * pi * ((j+1)/2)) for j>=1
};
void hoir_make_fourier_table(int n, float arg_scale, FourierTable* t);

namespace hoir {

__host__ __device__ inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
__host__ __device__ inline int piecewise_stride(int kind, int n) {
  return kind == HOIR_CONTINUOUS ? n - 1 : n;
}

// id_min = clamp(trunc(((x + half) / length) * segments), 0, segments-1)   (A.3)
// length == 2 (every layer built through Net): the division is an exact multiplication by 0.5,
// bit-identical to the IEEE division and ~10x cheaper.
__device__ __forceinline__ int segment_index(float x, float half, float length, float segf,
                                             int segments) {
  const float s = __fadd_rn(x, half);
  float t = __fmul_rn(length == 2.f ? __fmul_rn(s, 0.5f) : __fdiv_rn(s, length), segf);
  int id = __float2int_rz(t);  // trunc toward zero, saturating; NaN -> 0
  id = min(id, segments - 1);
  id = max(id, 0);
  return id;
}

// x_in = length * ((x - x_min) / (x_max - x_min)) - half,  x_m* = (id / segments) * 2 - 1  (A.3)
// POW2: id/segments, x_max - x_min and the quotient are exact, so the divisions reduce to
// multiplications with bit-identical results.
template <bool POW2>
__device__ __forceinline__ float local_coordinate(float x, int id, float half, float length,
                                                  float segf, float inv_segf) {
  if (POW2) {
    float x_min = __fsub_rn(__fmul_rn(__fmul_rn((float)id, inv_segf), 2.f), 1.f);
    float q = __fmul_rn(__fsub_rn(x, x_min), __fmul_rn(segf, 0.5f));
    return __fsub_rn(__fmul_rn(length, q), half);
  } else {
    float x_min = __fsub_rn(__fmul_rn(__fdiv_rn((float)id, segf), 2.f), 1.f);
    float x_max = __fsub_rn(__fmul_rn(__fdiv_rn((float)(id + 1), segf), 2.f), 1.f);
    float q = __fdiv_rn(__fsub_rn(x, x_min), __fsub_rn(x_max, x_min));
    return __fsub_rn(__fmul_rn(length, q), half);
  }
}
// d x_in / d x
template <bool POW2>
__device__ __forceinline__ float local_slope(int id, float length, float segf) {
  if (POW2) {
    return __fmul_rn(length, __fmul_rn(segf, 0.5f));
  } else {
    float x_min = __fsub_rn(__fmul_rn(__fdiv_rn((float)id, segf), 2.f), 1.f);
    float x_max = __fsub_rn(__fmul_rn(__fdiv_rn((float)(id + 1), segf), 2.f), 1.f);
    return __fdiv_rn(length, __fsub_rn(x_max, x_min));
  }
}

// Reflecting periodic map of the package (A.7); returns x' and d x'/d x in *sign.
__device__ __forceinline__ float make_periodic(float x, float p, float* sign) {
  float xp = x + 0.5f * p;
  float two_p = 2.f * p;
  xp = xp - two_p * floorf(xp / two_p);
  float s = 1.f;
  if (xp > p) { xp = two_p - xp; s = -1.f; }
  *sign = s;
  return xp - 0.5f * p;
}

// Lagrange values (and derivatives) at xi for compile-time n, nodes X / inverse denominators D.
template <int N>
__device__ __forceinline__ void lagrange(const float* __restrict__ X, const float* __restrict__ D,
                                         float xi, float* __restrict__ val) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float p = 1.f;
#pragma unroll
    for (int m = 0; m < N; ++m)
      if (m != j) p *= (xi - X[m]);
    val[j] = p * D[j];
  }
}
template <int N>
__device__ __forceinline__ void lagrange_deriv(const float* __restrict__ X,
                                               const float* __restrict__ D, float xi,
                                               float* __restrict__ val, float* __restrict__ der) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float p = 1.f, s = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m)
      if (m != j) {
        float d = xi - X[m];
        s = fmaf(s, d, p);
        p *= d;
      }
    val[j] = p * D[j];
    der[j] = s * D[j];
  }
}
// n = 3 on the unit Chebyshev-Lobatto nodes X = {-1, 0, 1}, D = {1/2, -1, 1/2} (both exact in fp32; every piecewise
// caller of the templates uses the unit table, hoir_make_lagrange_table(n, 1)): the same products in the same order
// as the generic loops -- bit-identical results -- without the subtractions of zero, the multiplications by one
// and the table loads (12 -> 7 and 27 -> 13 operations; the basis is evaluated ~240 times per training sample).
template <>
__device__ __forceinline__ void lagrange<3>(const float* __restrict__, const float* __restrict__, float xi,
                                            float* __restrict__ val) {
  const float a = xi - 1.f, b = xi + 1.f;
  val[0] = (xi * a) * 0.5f;
  val[1] = (b * a) * -1.f;
  val[2] = (b * xi) * 0.5f;
}
template <>
__device__ __forceinline__ void lagrange_deriv<3>(const float* __restrict__, const float* __restrict__, float xi,
                                                  float* __restrict__ val, float* __restrict__ der) {
  const float a = xi - 1.f, b = xi + 1.f;
  val[0] = (xi * a) * 0.5f;
  val[1] = (b * a) * -1.f;
  val[2] = (b * xi) * 0.5f;
  der[0] = (a + xi) * 0.5f;
  der[1] = (a + b) * -1.f;
  der[2] = (xi + b) * 0.5f;
}
// runtime-n variants (generic per-layer kernels)
__device__ __forceinline__ void lagrange_rt(int n, const float* X, const float* D, float xi,
                                            float* val, float* der) {
  for (int j = 0; j < n; ++j) {
    float p = 1.f, s = 0.f;
    for (int m = 0; m < n; ++m)
      if (m != j) {
        float d = xi - X[m];
        s = fmaf(s, d, p);
        p *= d;
      }
    val[j] = p * D[j];
    if (der) der[j] = s * D[j];
  }
}
// Fourier basis of the package (A.5): phi_0 = 1/2, then sin/cos of c[j]*x/length.
__device__ __forceinline__ void fourier_rt(int n, const float* c, float x, float length,
                                           int odd_is_sin, float* val, float* der) {
  val[0] = 0.5f;
  if (der) der[0] = 0.f;
  for (int j = 1; j < n; ++j) {
    float a = __fdiv_rn(__fmul_rn(c[j], x), length);
    float sn, cs;
    sincosf(a, &sn, &cs);
    bool is_sin = ((j & 1) == 1) == (odd_is_sin != 0);
    val[j] = is_sin ? sn : cs;
    if (der) der[j] = (is_sin ? cs : -sn) * (c[j] / length);
  }
}

// fire-and-forget fp32 add to global memory.  Explicit PTX: once a kernel also contains fences / acquire
// loads (the in-kernel optimizer tail), nvcc turns atomicAdd with an unused result into the RETURNING form
// (ATOMG instead of REDG), which stalls the issuing warp on the round trip (measured: +25% kernel time).
__device__ __forceinline__ void red_add(float* addr, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// positions_from_mesh (A.6) for one pixel; out[2*rotations].
__device__ __forceinline__ void mesh_position(const hoir_mesh_desc& m, long long pixel,
                                              float* __restrict__ out, int max_rot) {
  int row, col;
  if (pixel < 0x7fffffffLL) {   // 32-bit division (a quarter of the instructions of the 64-bit one) whenever it fits
    row = (int)((unsigned)pixel / (unsigned)m.cols);
    col = (int)((unsigned)pixel - (unsigned)row * (unsigned)m.cols);
  } else {
    row = (int)(pixel / m.cols);
    col = (int)(pixel - (long long)row * m.cols);
  }
  float xv = (float)row, yv = (float)col;
  if (m.normalize_mode == 1) {
    const int msi = max(m.rows, m.cols);
    const float ms = (float)msi;
    if ((msi & (msi - 1)) == 0) {   // power of two (4096 ...): the division is an exact scaling, bit-identical
      const float inv = __int_as_float(0x7f000000 - __float_as_int(ms));   // 2^-k from 2^k: exponent arithmetic
      xv = __fmul_rn(__fsub_rn(__fmul_rn(xv, inv), 0.5f), 2.f);
      yv = __fmul_rn(__fsub_rn(__fmul_rn(yv, inv), 0.5f), 2.f);
    } else {
      xv = __fmul_rn(__fsub_rn(__fdiv_rn(xv, ms), 0.5f), 2.f);
      yv = __fmul_rn(__fsub_rn(__fdiv_rn(yv, ms), 0.5f), 2.f);
    }
  } else if (m.normalize_mode == 2) {
    xv = __fsub_rn(__fmul_rn(__fdiv_rn(xv, (float)(m.rows - 1)), 2.f), 1.f);
    yv = __fsub_rn(__fmul_rn(__fdiv_rn(yv, (float)(m.cols - 1)), 2.f), 1.f);
  }
#pragma unroll
  for (int r = 0; r < HOIR_MAX_ROTATIONS; ++r) {
    if (r < max_rot && r < m.rotations) {
      float rx = m.rot_x[r], ry = m.rot_y[r], rs = m.rot_sum[r];
      const float u = __fadd_rn(__fmul_rn(rx, xv), __fmul_rn(ry, yv));
      const float v = __fsub_rn(__fmul_rn(rx, yv), __fmul_rn(ry, xv));
      if (rs == 1.f) {              // rotation 0 (cos 0 + sin 0 = 1): x / 1 is x
        out[2 * r] = u;
        out[2 * r + 1] = v;
      } else {
        out[2 * r] = __fdiv_rn(u, rs);
        out[2 * r + 1] = __fdiv_rn(v, rs);
      }
    }
  }
}

}  // namespace hoir
