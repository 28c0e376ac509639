// hoir_common.cu -- error state, basis tables, optimizers and the FFMA peak micro-kernel.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include "hoir_common.cuh"

static thread_local char g_err[512] = "";

int hoir_set_error(int code, const char* msg, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where ? where : "hoir", msg ? msg : "error");
  return code;
}

extern "C" const char* hoir_last_error(void) { return g_err; }

namespace {
struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};
constexpr int kSlots = 6;
Workspace g_ws[16][kSlots];
// forward calls arrive on the caller's thread, backward calls on autograd's device thread: the table is shared
std::mutex g_ws_mutex;
}  // namespace

int hoir_get_workspace(int slot, size_t bytes, void** out) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  int dev = 0;
  HOIR_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || slot < 0 || slot >= kSlots)
    return hoir_set_error(HOIR_EUNSUPPORTED, "device index >= 16", "hoir_get_workspace");
  Workspace& w = g_ws[dev][slot];
  if (w.bytes < bytes) {
    if (w.ptr) {
      HOIR_CHECK_CUDA(cudaDeviceSynchronize());
      HOIR_CHECK_CUDA(cudaFree(w.ptr));
      w.ptr = nullptr;
      w.bytes = 0;
    }
    const size_t grow = bytes + bytes / 4;
    HOIR_CHECK_CUDA(cudaMalloc(&w.ptr, grow));
    HOIR_CHECK_CUDA(cudaMemset(w.ptr, 0, grow));   // slot 2 holds the grid-barrier counters
    // the memset runs on the legacy default stream; callers launch on non-blocking streams (PyTorch side streams,
    // graph capture warm-ups): make it visible to all of them before the first kernel can read the counters
    HOIR_CHECK_CUDA(cudaDeviceSynchronize());
    w.bytes = grow;
  }
  *out = w.ptr;
  return HOIR_OK;
}

int hoir_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int sms = cache[dev].load(std::memory_order_relaxed);
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    cache[dev].store(sms, std::memory_order_relaxed);
  }
  return sms;
}

void hoir_free_workspaces() {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 16; ++d)
    for (int s = 0; s < kSlots; ++s)
      if (g_ws[d][s].ptr) {
        cudaSetDevice(d);
        cudaFree(g_ws[d][s].ptr);
        g_ws[d][s].ptr = nullptr;
        g_ws[d][s].bytes = 0;
      }
  cudaSetDevice(cur);
}

extern "C" int hoir_shutdown(void) {
  hoir_free_workspaces();
  return HOIR_OK;
}
extern "C" int hoir_version(void) { return HOIR_VERSION; }

// Chebyshev-Lobatto nodes the way the package builds them: -cos(k*pi/(n-1))
// evaluated in fp32, |X| < 1e-15 snapped to 0.  D_j = 1 / prod_{m != j} (X_j - X_m).
void hoir_make_lagrange_table(int n, float half_length, LagrangeTable* t) {
  memset(t, 0, sizeof(*t));
  if (n < 1 || n > HOIR_MAX_N) return;
  if (n == 1) {
    t->X[0] = 0.f;
    t->D[0] = 1.f;
    return;
  }
  for (int k = 0; k < n; ++k) {
    float arg = (float)((double)k * M_PI / (double)(n - 1));
    float x = -cosf(arg);
    if (fabsf(x) < 1e-15f) x = 0.f;
    t->X[k] = half_length * x;
  }
  for (int j = 0; j < n; ++j) {
    double d = 1.0;
    for (int m = 0; m < n; ++m)
      if (m != j) d *= (double)t->X[j] - (double)t->X[m];
    t->D[j] = (float)(1.0 / d);
  }
}

void hoir_make_fourier_table(int n, float arg_scale, FourierTable* t) {
  memset(t, 0, sizeof(*t));
  for (int j = 1; j < n && j < HOIR_MAX_N; ++j)
    t->c[j] = (float)((double)arg_scale * M_PI * (double)((j + 1) / 2));
}

// ------------------------------------------------------------------------------------------
// Optimizers on flat fp32 vectors.
// torch.optim.Adam defaults as used at /root/reference/high_order_implicit_representation/
// networks.py:94-97 (no amsgrad, L2 weight decay added to the gradient).
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(long long n, float* __restrict__ w, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                            float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    float gk = g[k];
    float wk = w[k];
    if (wd != 0.f) gk = fmaf(wd, wk, gk);
    float mk = m[k] + (1.f - b1) * (gk - m[k]);          // lerp, as torch does
    float vk = b2 * v[k] + (1.f - b2) * gk * gk;
    m[k] = mk;
    v[k] = vk;
    float denom = sqrtf(vk) / bc2_sqrt + eps;
    w[k] = wk - (lr / bc1) * (mk / denom);
  }
}

// SparseLion: Lion restricted to entries whose gradient is non-zero.
__global__ void sparse_lion_kernel(long long n, float* __restrict__ w, const float* __restrict__ g,
                                   float* __restrict__ ea, float lr, float b1, float b2, float wd) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    float gk = g[k];
    if (gk != 0.f) {
      float e = ea[k];
      float wk = w[k] * (1.f - lr * wd);
      float u = e * b1 + gk * (1.f - b1);
      float s = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
      w[k] = wk - lr * s;
      ea[k] = e * b2 + gk * (1.f - b2);
    }
  }
}

static int flat_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return (int)g;
}

extern "C" int hoir_adam_step(long long n, float* w, const float* g, float* exp_avg,
                              float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                              float weight_decay, long long step, void* stream) {
  if (n < 0 || step < 1) return hoir_set_error(HOIR_EINVAL, "bad n/step", "hoir_adam_step");
  if (n == 0) return HOIR_OK;
  if (!w || !g || !exp_avg || !exp_avg_sq)
    return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_adam_step");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(
      n, w, g, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, (float)bc1,
      (float)sqrt(bc2));
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_sparse_lion_step(long long n, float* w, const float* g, float* exp_avg,
                                     float lr, float beta1, float beta2, float weight_decay,
                                     void* stream) {
  if (n < 0) return hoir_set_error(HOIR_EINVAL, "bad n", "hoir_sparse_lion_step");
  if (n == 0) return HOIR_OK;
  if (!w || !g || !exp_avg)
    return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_sparse_lion_step");
  sparse_lion_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, w, g, exp_avg, lr, beta1,
                                                                     beta2, weight_decay);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// The same updates with their scalars read from DEVICE memory, so that a step captured in a CUDA graph follows
// the step count (bias corrections) and the lr schedule without being re-captured: the host refreshes the 8
// floats (lr, beta1, beta2, eps, weight_decay, bias_correction1, sqrt(bias_correction2), unused) before a replay.
__global__ void adam_dev_kernel(long long n, float* __restrict__ w, const float* __restrict__ g,
                                float* __restrict__ m, float* __restrict__ v, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5], bc2_sqrt = hyper[6];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    float gk = g[k];
    float wk = w[k];
    if (wd != 0.f) gk = fmaf(wd, wk, gk);
    float mk = m[k] + (1.f - b1) * (gk - m[k]);
    float vk = b2 * v[k] + (1.f - b2) * gk * gk;
    m[k] = mk;
    v[k] = vk;
    float denom = sqrtf(vk) / bc2_sqrt + eps;
    w[k] = wk - (lr / bc1) * (mk / denom);
  }
}
__global__ void sparse_lion_dev_kernel(long long n, float* __restrict__ w, const float* __restrict__ g,
                                       float* __restrict__ ea, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], wd = hyper[4];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    float gk = g[k];
    if (gk != 0.f) {
      float e = ea[k];
      float wk = w[k] * (1.f - lr * wd);
      float u = e * b1 + gk * (1.f - b1);
      float s = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
      w[k] = wk - lr * s;
      ea[k] = e * b2 + gk * (1.f - b2);
    }
  }
}

struct Floats8 { float v[8]; };
__global__ void write_floats_kernel(float* __restrict__ dst, int n, Floats8 f) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = f.v[threadIdx.x];
}
// dst[0..n) = values[0..n) (n <= 8), stream-ordered, the values travel as kernel arguments: no host buffer has to
// stay alive or unchanged until the copy executes
extern "C" int hoir_write_floats(float* dst, int n, const float* values, void* stream) {
  if (n < 0 || n > 8) return hoir_set_error(HOIR_EINVAL, "n must be in 0..8", "hoir_write_floats");
  if (n == 0) return HOIR_OK;
  if (!dst || !values) return hoir_set_error(HOIR_EINVAL, "null pointer", "hoir_write_floats");
  Floats8 f;
  for (int k = 0; k < 8; ++k) f.v[k] = k < n ? values[k] : 0.f;
  write_floats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst, n, f);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_adam_step_dev(long long n, float* w, const float* g, float* exp_avg, float* exp_avg_sq,
                                  const float* hyper, void* stream) {
  if (n < 0) return hoir_set_error(HOIR_EINVAL, "bad n", "hoir_adam_step_dev");
  if (n == 0) return HOIR_OK;
  if (!w || !g || !exp_avg || !exp_avg_sq || !hyper)
    return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_adam_step_dev");
  adam_dev_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, w, g, exp_avg, exp_avg_sq, hyper);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

extern "C" int hoir_sparse_lion_step_dev(long long n, float* w, const float* g, float* exp_avg, const float* hyper,
                                         void* stream) {
  if (n < 0) return hoir_set_error(HOIR_EINVAL, "bad n", "hoir_sparse_lion_step_dev");
  if (n == 0) return HOIR_OK;
  if (!w || !g || !exp_avg || !hyper)
    return hoir_set_error(HOIR_EINVAL, "null device pointer", "hoir_sparse_lion_step_dev");
  sparse_lion_dev_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(n, w, g, exp_avg, hyper);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}

// ------------------------------------------------------------------------------------------
// FFMA-only micro-kernel: the measured FP32 roofline denominator.  16 independent
// register-register FFMA chains per thread, 1024 threads per SM-resident CTA pair.
// ------------------------------------------------------------------------------------------
template <bool REG_FORM>
__global__ void __launch_bounds__(512, 2)
ffma_peak_kernel(int iters, float a, float b, float* __restrict__ sink) {
  if (REG_FORM) {  // force a, b into registers: FFMA R, R, R, R (what a weight-from-LDS kernel issues)
    a += 1e-9f * (float)threadIdx.x;
    b += 1e-9f * (float)threadIdx.x;
  }
  float r[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) r[k] = (float)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < 16; ++k) r[k] = fmaf(r[k], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += r[k];
  if (s == 123.456f) sink[0] = s;
}

extern "C" int hoir_ffma_peak(int iters, int reg_form, float* ms, double* flops, void* stream) {
  if (iters < 1 || !ms || !flops) return hoir_set_error(HOIR_EINVAL, "bad arguments", "hoir_ffma_peak");
  int dev = 0, sms = 0;
  HOIR_CHECK_CUDA(cudaGetDevice(&dev));
  HOIR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* sink = nullptr;
  HOIR_CHECK_CUDA(cudaMalloc(&sink, sizeof(float)));
  cudaEvent_t e0, e1;
  HOIR_CHECK_CUDA(cudaEventCreate(&e0));
  HOIR_CHECK_CUDA(cudaEventCreate(&e1));
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = sms * 2 * 4;
  auto kern = reg_form ? ffma_peak_kernel<true> : ffma_peak_kernel<false>;
  kern<<<grid, 512, 0, s>>>(iters / 8 + 1, 1.0001f, 0.5f, sink);  // warm-up
  HOIR_CHECK_CUDA(cudaEventRecord(e0, s));
  kern<<<grid, 512, 0, s>>>(iters, 1.0001f, 0.5f, sink);
  HOIR_CHECK_CUDA(cudaEventRecord(e1, s));
  HOIR_CHECK_CUDA(cudaEventSynchronize(e1));
  HOIR_CHECK_CUDA(cudaEventElapsedTime(ms, e0, e1));
  *flops = 2.0 * 16.0 * 8.0 * (double)iters * 512.0 * (double)grid;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  HOIR_CHECK_CUDA(cudaGetLastError());
  return HOIR_OK;
}
