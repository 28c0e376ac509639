// neighborhood.cu -- patch gather / fold kernels of the neighborhood interpolator
// (image_neighborhood_dataset, /root/reference/high_order_implicit_representation/
// single_image_dataset.py:83-141, and the fold + crop + blend of neighborhood_sample_generator /
// generate_sequence, /root/reference/high_order_implicit_representation/rendering.py:28-126).
// HBM-bound byte movers: one thread per output element, output-coalesced, the image stays in L2.
#include "hoir_common.cuh"

namespace {

__device__ __forceinline__ int reflect(int v, int n) {   // nn.ReflectionPad2d index map
  if (v < 0) v = -v;
  if (v >= n) v = 2 * (n - 1) - v;
  return v;
}

// features[p][c * E + e] (donut cells in row-major order of the T x T window) and
// targets[p][c * w * w + (dh - o) * w + (dw - o)] for patches first .. first + count - 1
__global__ void gather_kernel(const float* __restrict__ img, int C, int H, int W, int width, int outside,
                              int stride, int pad, int nW, long long first, long long count,
                              float* __restrict__ features, float* __restrict__ targets) {
  const int T = width + 2 * outside, E = T * T - width * width, WW = width * width;
  const long long nf = features ? count * C * E : 0, nt = targets ? count * C * WW : 0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nf + nt;
       k += (long long)gridDim.x * blockDim.x) {
    const bool is_t = k >= nf;
    const long long kk = is_t ? k - nf : k;
    const int per = is_t ? C * WW : C * E;
    const long long pl = kk / per;
    const int r = (int)(kk - pl * per);
    const int cells = is_t ? WW : E;
    const int c = r / cells, e = r - c * cells;
    int dh, dw;
    if (is_t) {
      dh = outside + e / width;
      dw = outside + e % width;
    } else {
      // e-th donut cell: rows above the block are full, rows through the block skip `width` cells
      const int top = outside * T;
      if (e < top) {
        dh = e / T; dw = e % T;
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
    const long long p = first + pl;
    const int ph = (int)(p / nW), pw = (int)(p - (long long)ph * nW);
    const int y = reflect(ph * stride + dh - pad, H), x = reflect(pw * stride + dw - pad, W);
    const float v = __ldg(img + ((size_t)c * H + y) * W + x);
    if (is_t) targets[kk] = v; else features[kk] = v;
  }
}

// Same gather, table-driven: the (channel, dh, dw) of every feature / target column is computed ONCE per CTA into
// shared memory (the element order has ~10 runtime divisions per column); a warp then copies one patch at a time,
// lanes over consecutive columns (coalesced stores, the image rows of a patch stay in L1/L2).
__global__ void __launch_bounds__(256)
gather_table_kernel(const float* __restrict__ img, int C, int H, int W, int width, int outside, int stride, int pad,
                    int nW, long long first, long long count, float* __restrict__ features,
                    float* __restrict__ targets) {
  extern __shared__ int tab[];                      // [C * E] features then [C * WW] targets: c << 16 | dh << 8 | dw
  const int T = width + 2 * outside, E = T * T - width * width, WW = width * width;
  const int nfe = C * E, nte = C * WW;
  for (int r = threadIdx.x; r < nfe + nte; r += blockDim.x) {
    const bool is_t = r >= nfe;
    const int rr = is_t ? r - nfe : r, cells = is_t ? WW : E;
    const int c = rr / cells, e = rr - c * cells;
    int dh, dw;
    if (is_t) {
      dh = outside + e / width;
      dw = outside + e % width;
    } else {
      const int top = outside * T;
      if (e < top) {
        dh = e / T; dw = e % T;
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
    tab[r] = (c << 16) | (dh << 8) | dw;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const size_t plane = (size_t)H * W;
  for (long long pl = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); pl < count; pl += (long long)gridDim.x * wpb) {
    const long long p = first + pl;
    const int ph = (int)(p / nW), pw = (int)(p - (long long)ph * nW);
    const int y0 = ph * stride - pad, x0 = pw * stride - pad;
    if (features) {
      float* dst = features + pl * nfe;
      for (int r = lane; r < nfe; r += 32) {
        const int t = tab[r];
        const int y = reflect(y0 + ((t >> 8) & 255), H), x = reflect(x0 + (t & 255), W);
        dst[r] = __ldg(img + (size_t)(t >> 16) * plane + (size_t)y * W + x);
      }
    }
    if (targets) {
      float* dst = targets + pl * nte;
      for (int r = lane; r < nte; r += 32) {
        const int t = tab[nfe + r];
        const int y = reflect(y0 + ((t >> 8) & 255), H), x = reflect(x0 + (t & 255), W);
        dst[r] = __ldg(img + (size_t)(t >> 16) * plane + (size_t)y * W + x);
      }
    }
  }
}

// out[c][r][col] = alpha * prev[c][r][col] + (1 - alpha) * block value predicted for that pixel, where the
// predictions[p][c * w * w + kh * w + kw] come from stride-w patches of the image reflection-padded by 2*o + w
__global__ void fold_kernel(const float* __restrict__ pred, int C, int H, int W, int width, int outside, int nW,
                            float alpha, const float* __restrict__ prev, float* __restrict__ out) {
  const int fg = width + outside, WW = width * width;
  const long long total = (long long)C * H * W;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
       k += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(k % W);
    const long long t = k / W;
    const int r = (int)(t % H), c = (int)(t / H);
    const int yy = r + fg, xx = col + fg;
    const int ph = yy / width, kh = yy - ph * width, pw = xx / width, kw = xx - pw * width;
    const float v = __ldg(pred + ((size_t)ph * nW + pw) * (C * WW) + c * WW + kh * width + kw);
    out[k] = prev ? alpha * __ldg(prev + k) + (1.f - alpha) * v : v;
  }
}

}  // namespace

extern "C" long long hoir_neighborhood_patches(int H, int W, int width, int outside, int stride, int pad,
                                               int* n_rows, int* n_cols) {
  const int T = width + 2 * outside;
  if (H < 1 || W < 1 || width < 1 || outside < 0 || stride < 1 || pad < 0) return -1;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  if (Hp < T || Wp < T) return 0;
  const int nH = (Hp - T) / stride + 1, nW = (Wp - T) / stride + 1;
  if (n_rows) *n_rows = nH;
  if (n_cols) *n_cols = nW;
  return (long long)nH * nW;
}

extern "C" int hoir_neighborhood_gather(const float* image, int C, int H, int W, int width, int outside,
                                        int stride, int pad, long long first_patch, long long count,
                                        float* features, float* targets, void* stream) {
  int nH = 0, nW = 0;
  const long long total = hoir_neighborhood_patches(H, W, width, outside, stride, pad, &nH, &nW);
  if (total < 0 || C < 1) return hoir_set_error(HOIR_EINVAL, "bad image / patch geometry", "hoir_neighborhood_gather");
  if (pad > 0 && (pad >= H || pad >= W))
    return hoir_set_error(HOIR_EINVAL, "reflection padding must be smaller than the image", "hoir_neighborhood_gather");
  if (first_patch < 0 || count < 0 || first_patch + count > total)
    return hoir_set_error(HOIR_EINVAL, "patch range out of bounds", "hoir_neighborhood_gather");
  if (count == 0 || (!features && !targets)) return HOIR_OK;
  if (!image) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_neighborhood_gather");
  const int T = width + 2 * outside;
  const size_t tab_bytes = (size_t)C * T * T * sizeof(int);
  if (T <= 255 && C <= 32767 && tab_bytes <= 48 * 1024) {
    long long g2 = (count + 7) / 8;
    if (g2 > 148 * 16) g2 = 148 * 16;
    gather_table_kernel<<<(int)g2, 256, tab_bytes, (cudaStream_t)stream>>>(image, C, H, W, width, outside, stride, pad, nW,
                                                                         first_patch, count, features, targets);
    HOIR_CHECK_CUDA(cudaGetLastError());
    return HOIR_OK;
  }
  const long long n = count * C * ((features ? T * T - width * width : 0) + (targets ? width * width : 0));
  long long grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  gather_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(image, C, H, W, width, outside, stride, pad, nW,
                                                            first_patch, count, features, targets);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_neighborhood_fold(const float* predictions, int C, int H, int W, int width, int outside,
                                      float alpha, const float* prev_image, float* out_image, void* stream) {
  const int pad = 2 * outside + width;
  int nH = 0, nW = 0;
  const long long total = hoir_neighborhood_patches(H, W, width, outside, width, pad, &nH, &nW);
  if (total <= 0 || C < 1) return hoir_set_error(HOIR_EINVAL, "bad image / patch geometry", "hoir_neighborhood_fold");
  if (!predictions || !out_image) return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_neighborhood_fold");
  long long grid = ((long long)C * H * W + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  fold_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(predictions, C, H, W, width, outside, nW, alpha, prev_image,
                                                          out_image);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
