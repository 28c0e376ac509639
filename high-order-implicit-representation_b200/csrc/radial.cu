// radial.cu -- the radial random sampler of the random-interpolation path
// (random_symmetric_sample + random_radial_samples_from_image,
//  /root/reference/high_order_implicit_representation/random_sample_dataset.py:176-253, consumed by
//  RadialRandomImageSampleDataset :122-173 and generate_sample_radial, utils.py:20-97).
//
// For every target pixel k of an S x S image, F interpolation points are drawn at random (r, theta):
//   dx = trunc(((0.5 r) cos theta) S), dy = trunc(((0.5 r) sin theta) S)     (fp32, this op order)
//   (px, py) = reflect(target + (dx, dy))            values > S-1 -> 2(S-1) - v, then values < 0 -> -v
//   features[k][f] = (image[:, py, px], stripes[:, py, px] - stripes[:, ty, tx]),  targets[k] = image[:, ty, tx]
// where pixel (x, y) is the element x + y*S of a channel plane (the reference's linear index).
// HBM-bound byte mover: one thread per (target, point), a warp writes 32 consecutive feature rows; the image
// and the stripe planes (S*S*(C+P) floats) stay in L1/L2.  (r, theta) come either from caller tensors (parity
// with the reference's torch.rand stream) or from a counter-based Philox4x32-10 generator in the kernel.
#include "hoir_common.cuh"

namespace {

struct Philox {
  // Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
  static constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static void run(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
      const unsigned long long p0 = (unsigned long long)M0 * c[0], p1 = (unsigned long long)M1 * c[2];
      const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c[1] ^ k0, n2 = (unsigned int)(p0 >> 32) ^ c[3] ^ k1;
      c[1] = (unsigned int)p1;
      c[3] = (unsigned int)p0;
      c[0] = n0;
      c[2] = n2;
      k0 += W0;
      k1 += W1;
    }
  }
};

// the e-th (r, theta) pair of stream (seed, offset): r uniform in [0, 1) on a 2^-24 grid (torch.rand's fp32 grid),
// theta = (u * 2) * fl32(pi), the reference's `torch.rand(...) * 2 * math.pi`
__host__ __device__ inline void radial_uniforms(unsigned long long seed, unsigned long long offset, long long e,
                                                float& r, float& theta) {
  unsigned int c[4] = {(unsigned int)e, (unsigned int)((unsigned long long)e >> 32), (unsigned int)offset,
                       (unsigned int)(offset >> 32)};
  Philox::run(c, (unsigned int)seed, (unsigned int)(seed >> 32));
  r = (float)(c[0] & 0xffffffu) * 5.9604644775390625e-08f;
  const float u = (float)(c[1] & 0xffffffu) * 5.9604644775390625e-08f;
  theta = u * 2.f * 3.14159274101257324f;
}

__device__ __forceinline__ int reflect_point(int v, int S) {
  if (v > S - 1) v = 2 * (S - 1) - v;
  if (v < 0) v = -v;
  return v;
}

// offsets of one interpolation point; the multiplications are separately rounded as in the reference expression
__device__ __forceinline__ void radial_offsets(float r, float theta, float Sf, int& dx, int& dy) {
  float sn, cs;
  sn = sinf(theta);
  cs = cosf(theta);
  const float hr = __fmul_rn(0.5f, r);
  dx = __float2int_rz(__fmul_rn(__fmul_rn(hr, cs), Sf));
  dy = __float2int_rz(__fmul_rn(__fmul_rn(hr, sn), Sf));
}

__global__ void uniforms_kernel(long long n, unsigned long long seed, unsigned long long offset, float* __restrict__ r,
                                float* __restrict__ theta) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float a, b;
    radial_uniforms(seed, offset, e, a, b);
    if (r) r[e] = a;
    if (theta) theta[e] = b;
  }
}

// xy[k][f] = (px, py) for target k at samples[k] (or the k-th grid point: (k / S, k % S) -- indices_from_grid)
__global__ void indices_kernel(int S, int F, long long N, const int* __restrict__ samples, const float* __restrict__ r,
                               const float* __restrict__ theta, unsigned long long seed, unsigned long long offset,
                               int* __restrict__ xy) {
  const float Sf = (float)S;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * F; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / F;
    int tx, ty;
    if (samples) {
      tx = samples[2 * k];
      ty = samples[2 * k + 1];
    } else {
      const int kk = (int)(k % ((long long)S * S));
      tx = kk / S;
      ty = kk - tx * S;
    }
    float a, b;
    if (r) { a = r[e]; b = theta[e]; } else radial_uniforms(seed, offset, e, a, b);
    int dx, dy;
    radial_offsets(a, b, Sf, dx, dy);
    xy[2 * e] = reflect_point(tx + dx, S);
    xy[2 * e + 1] = reflect_point(ty + dy, S);
  }
}

// features [M*S*S][F][C+P], targets [M*S*S][C]; target k of image m is grid point (k / S, k % S).
// A CTA's 256 consecutive (target, point) pairs own a CONTIGUOUS block of 256 * (C+P) output floats: the rows are
// assembled in shared memory and leave as coalesced 16-byte stores (4-byte stores at a 28-byte stride reached 15 %
// of the copy bandwidth).
template <int CP>   // C + P when it is a compile-time width (0: generic loop, direct stores)
__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ img, const float* __restrict__ stripes, int M, int C, int P, int S, int F,
              const float* __restrict__ r, const float* __restrict__ theta, unsigned long long seed,
              unsigned long long offset, float* __restrict__ features, float* __restrict__ targets) {
  __shared__ __align__(16) float srow[CP ? 256 * CP : 4];
  const float Sf = (float)S;
  const long long SS = (long long)S * S, N = (long long)M * SS, total = N * F;
  const int width = CP ? CP : C + P;
  for (long long e0 = blockIdx.x * 256LL; e0 < total; e0 += (long long)gridDim.x * 256) {
    const long long e = e0 + threadIdx.x;
    if (e < total) {
      const long long k = e / F;
      const int f = (int)(e - k * F);
      const long long m = k / SS;
      const int kk = (int)(k - m * SS);
      const int tx = kk / S, ty = kk - tx * S;
      const int tlin = tx + ty * S;
      float a, b;
      if (r) { a = __ldg(r + e); b = __ldg(theta + e); } else radial_uniforms(seed, offset, e, a, b);
      int dx, dy;
      radial_offsets(a, b, Sf, dx, dy);
      const int plin = reflect_point(tx + dx, S) + reflect_point(ty + dy, S) * S;
      const float* im = img + m * C * SS;
      float* dst = CP ? srow + threadIdx.x * CP : features + e * width;
      const int Cc = CP ? 3 : C, Pc = CP ? CP - 3 : P;   // compile-time widths for the RGB + 1/2-rotation cases
#pragma unroll
      for (int c = 0; c < Cc; ++c) dst[c] = __ldg(im + c * SS + plin);
#pragma unroll
      for (int q = 0; q < Pc; ++q) dst[Cc + q] = __fsub_rn(__ldg(stripes + q * SS + plin), __ldg(stripes + q * SS + tlin));
      if (f == 0 && targets) {
        for (int c = 0; c < C; ++c) targets[k * C + c] = __ldg(im + c * SS + tlin);
      }
    }
    if (CP) {
      __syncthreads();
      const long long left = total - e0;
      const int nfl = (int)(left < 256 ? left : 256) * CP;              // floats of this block; e0 * CP * 4 B is 16-aligned
      float* out = features + e0 * CP;
      for (int q = threadIdx.x * 4; q < nfl; q += 256 * 4) {
        if (q + 3 < nfl) *reinterpret_cast<float4*>(out + q) = *reinterpret_cast<const float4*>(srow + q);
        else for (int c = q; c < nfl; ++c) out[c] = srow[c];
      }
      __syncthreads();
    }
  }
}

int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int hoir_radial_uniforms(long long n, unsigned long long seed, unsigned long long offset, float* r,
                                    float* theta, void* stream) {
  if (n < 0) return hoir_set_error(HOIR_EINVAL, "negative count", "hoir_radial_uniforms");
  if (n == 0 || (!r && !theta)) return HOIR_OK;
  uniforms_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, seed, offset, r, theta);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_radial_indices(int S, int F, long long N, const int* samples, const float* r, const float* theta,
                                   unsigned long long seed, unsigned long long offset, int* xy, void* stream) {
  if (S < 1 || F < 0 || N < 0) return hoir_set_error(HOIR_EINVAL, "bad image size / point count", "hoir_radial_indices");
  if ((r == nullptr) != (theta == nullptr))
    return hoir_set_error(HOIR_EINVAL, "r and theta must be given together", "hoir_radial_indices");
  if (N == 0 || F == 0) return HOIR_OK;
  if (!xy) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_radial_indices");
  indices_kernel<<<grid_for(N * F), 256, 0, (cudaStream_t)stream>>>(S, F, N, samples, r, theta, seed, offset, xy);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_radial_sample(const float* images, const float* stripes, int M, int C, int P, int S, int F,
                                  const float* r, const float* theta, unsigned long long seed,
                                  unsigned long long offset, float* features, float* targets, void* stream) {
  if (M < 0 || C < 1 || P < 0 || S < 1 || F < 0)
    return hoir_set_error(HOIR_EINVAL, "bad image batch / point count", "hoir_radial_sample");
  if ((long long)S * S > 0x7fffffffLL / 2) return hoir_set_error(HOIR_EINVAL, "image too large", "hoir_radial_sample");
  if ((r == nullptr) != (theta == nullptr))
    return hoir_set_error(HOIR_EINVAL, "r and theta must be given together", "hoir_radial_sample");
  if (M == 0) return HOIR_OK;
  if (!images || (P > 0 && !stripes) || (F > 0 && !features))
    return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_radial_sample");
  const long long N = (long long)M * S * S;
  if (F == 0) {   // targets only: one point per target, nothing written to features
    if (!targets) return HOIR_OK;
    return hoir_set_error(HOIR_EINVAL, "at least one interpolation point is required", "hoir_radial_sample");
  }
  const int g = grid_for(N * F);
  if (C + P == 7 && C == 3)
    sample_kernel<7><<<g, 256, 0, (cudaStream_t)stream>>>(images, stripes, M, C, P, S, F, r, theta, seed, offset, features, targets);
  else if (C + P == 5 && C == 3)
    sample_kernel<5><<<g, 256, 0, (cudaStream_t)stream>>>(images, stripes, M, C, P, S, F, r, theta, seed, offset, features, targets);
  else
    sample_kernel<0><<<g, 256, 0, (cudaStream_t)stream>>>(images, stripes, M, C, P, S, F, r, theta, seed, offset, features, targets);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
