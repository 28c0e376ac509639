/* hoir.h -- C ABI of libhoir_b200.so (hand-written CUDA for sm_100a).
 *
 * The reference (jloveric/high-order-implicit-representation) has no native code and no FFI:
 * its hot path sits behind the Python constructors of the un-vendored package
 * high-order-layers-torch (star-imported at
 * /root/reference/high_order_implicit_representation/networks.py:7-8).  The drop-in boundary is
 * therefore (a) this C ABI -- plain pointers and sizes, no torch types -- and (b) the Python
 * mirror of the package's constructors in high-order-implicit-representation_b200/, which binds
 * these entry points with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer named x, w, y, g*, target, ... is a DEVICE pointer owned by the caller,
 *     row-major contiguous fp32 unless stated; the library never allocates or frees user memory;
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it;
 *   - return 0 on success; otherwise a HOIR_E* code, message via hoir_last_error() (thread local);
 *   - B == 0 is legal everywhere (the reference render loops feed one empty batch,
 *     /root/reference/high_order_implicit_representation/rendering.py:200);
 *   - weight layout at this boundary is the package's  w[out_features][in_features][nodes];
 *   - there is NO CPU fallback: a host pointer is a caller error.
 */
#ifndef HOIR_H
#define HOIR_H

#ifdef __cplusplus
extern "C" {
#endif

#define HOIR_VERSION 106
#define HOIR_MAX_LAYERS 8
#define HOIR_MAX_N 32          /* max basis functions per segment / Fourier terms          */
#define HOIR_MAX_ROTATIONS 8

enum hoir_error {
  HOIR_OK = 0,
  HOIR_EINVAL = 1,             /* bad descriptor / shape / null pointer -> ValueError       */
  HOIR_EUNSUPPORTED = 2,       /* no fused instantiation for this network -> use per-layer   */
  HOIR_ECUDA = 3               /* CUDA runtime error -> RuntimeError                         */
};

/* layer_type strings of high_order_fc_layers(...),
 * /root/reference/high_order_implicit_representation/networks.py:148-161 */
enum hoir_kind {
  HOIR_CONTINUOUS = 0,         /* PiecewisePolynomial: nodes = (n-1)*segments+1             */
  HOIR_DISCONTINUOUS = 1,      /* PiecewiseDiscontinuousPolynomial: nodes = n*segments      */
  HOIR_POLYNOMIAL = 2,         /* Polynomial: nodes = n, no segment map                      */
  HOIR_FOURIER = 3             /* FourierSeries: nodes = n (1/2, then sin/cos pairs)         */
};

/* One high-order layer.  Replaces the kwargs of high_order_fc_layers as forwarded by
 * HighOrderMLP (/root/reference/high_order_implicit_representation/networks.py:41-55). */
typedef struct hoir_layer_desc {
  int kind;                    /* enum hoir_kind                                             */
  int n;                       /* basis functions per segment (Fourier: number of terms)     */
  int segments;                /* piecewise kinds only; >= 1                                  */
  int in_features;
  int out_features;
  float length;                /* `scale` kwarg, 2.0 through Net                             */
  float rescale;               /* output multiplier: 1/in_features if rescale_output else 1  */
  float periodicity;           /* <= 0: none                                                  */
  float fourier_arg_scale;     /* basis argument = fourier_arg_scale*pi*i*x/length           */
  int fourier_odd_is_sin;      /* 1: odd j -> sin, even j -> cos                             */
} hoir_layer_desc;

/* HighOrderMLP = L_0, (MaxAbs, L_k) x hidden_layers, MaxAbs, L_last. */
typedef struct hoir_mlp_desc {
  int num_layers;              /* hidden_layers + 2                                          */
  hoir_layer_desc layers[HOIR_MAX_LAYERS];
  int normalize;               /* 1: MaxAbsNormalization between layers, 0: none             */
  float norm_eps;              /* 1e-6                                                        */
} hoir_mlp_desc;

/* In-kernel coordinate source: positions_from_mesh(width=rows, height=cols, rotations,
 * normalize) followed by stack(dim=2).reshape(-1, 2*rotations)
 * (/root/reference/high_order_implicit_representation/single_image_dataset.py:39-47).
 * Sample k of a call is pixel first_pixel + k in row-major (row, col) order. */
typedef struct hoir_mesh_desc {
  int rows, cols;
  int rotations;               /* in_features of layer 0 must equal 2*rotations              */
  int normalize_mode;          /* 0: raw integers; 1: (v/max(rows,cols)-0.5)*2;
                                  2: v/(size-1)*2-1                                           */
  long long first_pixel;
  float rot_x[HOIR_MAX_ROTATIONS], rot_y[HOIR_MAX_ROTATIONS], rot_sum[HOIR_MAX_ROTATIONS];
} hoir_mesh_desc;

/* Target source for the fused training step. */
enum hoir_target_kind {
  HOIR_TARGET_F32 = 0,         /* float target[B][out]                                        */
  HOIR_TARGET_U8_IMAGE = 1     /* uint8 image[rows*cols][3], indexed by pixel; the kernel
                                  applies  u*2/255-1  (single_image_dataset.py:49)            */
};

const char* hoir_last_error(void);
int hoir_version(void);
/* frees the library-owned per-device workspace (two-pass layer-0 gradient of large generic batches);
 * one training step may be in flight per device at a time because of that workspace */
int hoir_shutdown(void);
/* number of weight nodes per (out, in) link for a layer descriptor, or -1 */
int hoir_layer_nodes(const hoir_layer_desc* d);

/* ---- single layer (replaces <layer>.forward / autograd backward of the package) -------- */
int hoir_layer_forward(const hoir_layer_desc* d, long long B, const float* x, const float* w,
                       float* y, void* stream);
/* gx may be NULL (no input gradient); gw may be NULL; gw is ACCUMULATED into. */
int hoir_layer_backward(const hoir_layer_desc* d, long long B, const float* x, const float* w,
                        const float* gy, float* gx, float* gw, void* stream);
/* Debug/parity entry: the integer part of the path.  id_min[B*in], wid_min[B*in] (first node
 * of the gathered window), x_local[B*in] (may be NULL).  Piecewise kinds only. */
int hoir_segment_index(const hoir_layer_desc* d, long long B, const float* x, int* id_min,
                       int* wid_min, float* x_local, void* stream);

/* ---- MaxAbsNormalization over dim 1 ------------------------------------------------------ */
int hoir_maxabs_forward(long long B, int width, float eps, const float* x, float* y, void* stream);
int hoir_maxabs_backward(long long B, int width, float eps, const float* x, const float* gy,
                         float* gx, void* stream);

/* ---- coordinates --------------------------------------------------------------------------*/
/* x[B][2*rotations] for pixels first_pixel .. first_pixel+B-1 (bit-identical to the host mesh) */
int hoir_mesh_positions(const hoir_mesh_desc* mesh, long long B, float* x, void* stream);

/* ---- neighborhood interpolator: patch gather / fold ------------------------------------------
 * image [C][H][W] fp32 in [-1, 1].  Patches are the T x T windows (T = width + 2*outside) of the image
 * reflection-padded by `pad`, taken every `stride` pixels in row-major order; patch p yields
 *   features[p][c*(T*T - w*w) + e]   e = rank of the cell among the donut cells, row-major
 *   targets [p][c*w*w + kh*w + kw]   the inner w x w block
 * (image_neighborhood_dataset, /root/reference/high_order_implicit_representation/
 *  single_image_dataset.py:83-141).  features / targets may be NULL. */
long long hoir_neighborhood_patches(int H, int W, int width, int outside, int stride, int pad,
                                    int* n_rows, int* n_cols);
int hoir_neighborhood_gather(const float* image, int C, int H, int W, int width, int outside, int stride,
                             int pad, long long first_patch, long long count, float* features,
                             float* targets, void* stream);
/* predictions[p][C*w*w] of the stride-w patches of the image padded by 2*outside + width ->
 * out = alpha*prev + (1-alpha)*fold(predictions) cropped to [C][H][W]  (prev NULL: out = fold);
 * neighborhood_sample_generator / generate_sequence,
 * /root/reference/high_order_implicit_representation/rendering.py:76-95,119 */
int hoir_neighborhood_fold(const float* predictions, int C, int H, int W, int width, int outside,
                           float alpha, const float* prev_image, float* out_image, void* stream);

/* ---- random interpolation: radial sampler -------------------------------------------------------
 * For every target pixel k of an S x S image, F interpolation points at random (r, theta):
 *   dx = trunc(((0.5 r) cos theta) S), dy = trunc(((0.5 r) sin theta) S), reflected at the image borders
 * (random_symmetric_sample, /root/reference/high_order_implicit_representation/random_sample_dataset.py:211-244).
 * r / theta [N][F] are caller tensors (r = rand, theta = rand * 2 * pi: the reference's stream) or, when both are
 * NULL, element e = k*F + f of the Philox4x32-10 stream (seed, offset) -- hoir_radial_uniforms writes that stream
 * out.  samples [N][2] int32 (x, y) target points, or NULL for the grid order of indices_from_grid (:17-21):
 * target k is the point (k / S, k % S), pixel x + y*S of a channel plane. */
int hoir_radial_uniforms(long long n, unsigned long long seed, unsigned long long offset, float* r, float* theta,
                         void* stream);
int hoir_radial_indices(int S, int F, long long N, const int* samples, const float* r, const float* theta,
                        unsigned long long seed, unsigned long long offset, int* xy /* [N][F][2] */, void* stream);
/* images [M][C][S][S], stripes [P][S][S] (the position planes of positions_from_mesh) ->
 *   features[m*S*S + k][f] = (image[:, p], stripes[:, p] - stripes[:, t])      [M*S*S][F][C+P]
 *   targets [m*S*S + k]    = image[:, t]                                      [M*S*S][C]     (may be NULL)
 * with t the target pixel and p the f-th interpolation point of target k
 * (random_radial_samples_from_image, random_sample_dataset.py:176-208; r / theta are [M*S*S][F] or NULL) */
int hoir_radial_sample(const float* images, const float* stripes, int M, int C, int P, int S, int F,
                       const float* r, const float* theta, unsigned long long seed, unsigned long long offset,
                       float* features, float* targets, void* stream);

/* ---- fused whole-network path ---------------------------------------------------------------
 * Packed weights: one flat fp32 buffer, layer l at float offset hoir_mlp_packed_offset(l) with
 * layout [in][node][out_padded] (out_padded = out rounded up to 4).  hoir_mlp_pack fills it from
 * the boundary-layout tensors; it must be refreshed whenever a w changes.  Networks trained by the tensor-core
 * kernel carry a second copy of the layer-0 weights behind the last layer, [in][half][node][out/2] (a thread's node
 * rows are then one contiguous run): hoir_mlp_packed_size includes it, hoir_mlp_pack and the optimizer tail
 * write both. */
int hoir_mlp_fused_supported(const hoir_mlp_desc* m);           /* 1 / 0 */
/* 2: forward AND training step are fused (== hoir_mlp_fused_supported); 1: the forward only -- the neighborhood
 * interpolator family (3-5 piecewise layers, n = 2 or 3, 1- or 2-segment hidden / output layers, <= 64 wide, <= 32
 * outputs: the stock /root/reference/config/neighborhood_config.yaml and the 216 -> 50 -> 50 -> 50 -> 27 network of
 * /root/reference/README.md:57): hoir_mlp_pack / hoir_mlp_forward /
 * hoir_interp_frame apply; a training step = hoir_mlp_forward_acts + the per-layer backward entry points; 0: neither */
int hoir_mlp_forward_supported(const hoir_mlp_desc* m);
long long hoir_mlp_packed_size(const hoir_mlp_desc* m);         /* floats, or -1 */
long long hoir_mlp_packed_offset(const hoir_mlp_desc* m, int layer);
int hoir_mlp_pack(const hoir_mlp_desc* m, const float* const* w, float* packed, void* stream);

/* y[B][out] = MLP(x).  Exactly one of x / mesh is non-NULL.  post_scale: 0 -> y, 1 -> 0.5*(y+1)
 * (the render loop, /root/reference/high_order_implicit_representation/rendering.py:213). */
int hoir_mlp_forward(const hoir_mlp_desc* m, long long B, const float* x,
                     const hoir_mesh_desc* mesh, const float* packed, float* y, int post_scale,
                     void* stream);

/* The same forward for TRAINING a forward-only-fused network (hoir_mlp_forward_supported == 1): additionally stores the
 * inputs of the layers l = 1 .. L-1 for the per-layer backward entry points (hoir_layer_backward,
 * hoir_maxabs_backward): acts is [L-1][2][B][hidden width]; [l-1][0] = the rows as layer l-1 produced them (input of
 * MaxAbsNormalization), [l-1][1] = the normalized rows (input of layer l; not written when normalize == 0).
 * Replaces the forward half of Net.eval_step for the interpolator
 * (/root/reference/high_order_implicit_representation/networks.py:64-70). */
int hoir_mlp_forward_acts(const hoir_mlp_desc* m, long long B, const float* x, const float* packed, float* y,
                          float* acts, void* stream);

/* One pass of the neighborhood interpolator over an image in ONE launch: reflection pad by 2*outside + width,
 * stride-width donut patches, model forward, fold, crop and blend
 *   out = alpha * prev + (1 - alpha) * new        (prev NULL: out = new)
 * -- neighborhood_sample_generator + the blend of generate_sequence,
 * /root/reference/high_order_implicit_representation/rendering.py:28-95, 119.  image / prev / out are [C][H][W];
 * the network must map C*((w+2o)^2 - w^2) features to C*w*w outputs; out must not alias image.  No feature tensor and
 * no activation touches HBM. */
int hoir_interp_frame(const hoir_mlp_desc* m, const float* image, int C, int H, int W, int width, int outside,
                      const float* packed, float alpha, const float* prev_image, float* out_image, void* stream);

/* One fused forward + MSE + backward pass (Net.eval_step + loss.backward(),
 * /root/reference/high_order_implicit_representation/networks.py:64-70):
 *   loss_sum[0] += sum_b sum_o (y_hat - target)^2 * loss_scale       (loss_scale = 1/(B_global*out))
 *   gw[l]       += dL/dw_l   in boundary layout [out][in][nodes], with L = MSE(mean) over B_global
 * Exactly one of x / mesh non-NULL; target per target_kind.  If gy_ext != NULL it is used as
 * dL/dy_hat[B][out] instead of the MSE (autograd backward) and target/loss_sum are ignored. */
int hoir_mlp_train_step(const hoir_mlp_desc* m, long long B, const float* x,
                        const hoir_mesh_desc* mesh, int target_kind, const void* target,
                        const float* gy_ext, const float* packed, float* const* gw,
                        float* loss_sum, float loss_scale, void* stream);

/* ---- fused training step with the optimizer in the same launch -------------------------------
 * State of one optimizer over the FLAT parameter vector of a network (all w tensors in boundary
 * layout, concatenated in layer order).  Replaces optimizer.zero_grad() / loss.backward() /
 * optimizer.step() of the reference's training loop
 * (/root/reference/high_order_implicit_representation/networks.py:64-70, 93-102). */
enum hoir_opt_kind { HOIR_OPT_ADAM = 1, HOIR_OPT_SPARSE_LION = 2 };
typedef struct hoir_opt_desc {
  int kind;                    /* enum hoir_opt_kind                                          */
  float lr, beta1, beta2, eps, weight_decay;
  long long step;              /* 1-based index of this update (Adam bias correction)         */
  float* flat_w;               /* [n_params]                                                   */
  float* flat_g;               /* [n_params + 1]: dL/dw then the loss accumulator; must be zero
                                  on entry of a training step and is left zeroed by the tail   */
  float* exp_avg;              /* [n_params]                                                   */
  float* exp_avg_sq;           /* [n_params] (Adam only)                                       */
  float* packed;               /* kernel-side weights (hoir_mlp_packed_size floats); kept in
                                  sync with flat_w by the tail                                 */
  float* loss_out;             /* [1] receives the loss of the step                            */
  int keep_grad;               /* 1: leave flat_g as computed (inspection / parity tests); the
                                  caller zeroes it before the next step                        */
} hoir_opt_desc;

/* optimizer update + packed-weight refresh + gradient zeroing + loss hand-over: ONE launch.
 * Multi-GPU: call after the all-reduce of flat_g. */
int hoir_mlp_optimizer_tail(const hoir_mlp_desc* m, const hoir_opt_desc* opt, void* stream);
/* hoir_mlp_train_step with gw = views of opt->flat_g, packed = opt->packed, loss_sum =
 * opt->flat_g + n_params, followed by the tail.  For coherent batches the tail runs INSIDE the
 * whole-network kernel behind a grid-wide barrier: one launch per training step. */
int hoir_mlp_train_step_fused(const hoir_mlp_desc* m, long long B, const float* x,
                              const hoir_mesh_desc* mesh, int target_kind, const void* target,
                              float loss_scale, const hoir_opt_desc* opt, void* stream);

/* ---- multi-GPU: gradient exchange over NVLink / NVSwitch peer memory INSIDE the training step ----------
 * Data-parallel training sums the flat gradient (tens of KB) over the ranks once per step.  Instead of a
 * separate NCCL all-reduce, every rank keeps its gradient in a SYMMETRIC block that all peers of the node
 * have mapped (CUDA IPC / fabric handles; the host side obtains the mappings, e.g. through
 * torch.distributed._symmetric_memory), and the optimizer tail of the step
 *   1. publishes "my gradient of step s is complete" into every peer's flag array (st.release.sys),
 *   2. waits for the same flag from every peer (ld.acquire.sys),
 *   3. reads all ranks' gradients in rank order (bit-identical sums on every rank => replicated weights
 *      stay bit-identical without a broadcast), applies the update, re-packs the weights,
 *   4. zeroes its OTHER parity buffer for the step after next (two gradient buffers alternate, so no
 *      second barrier is needed before a buffer is reused).
 * Block layout (floats): [gradient parity 0: stride][gradient parity 1: stride][flags: 64 x u32],
 * stride = hoir_peer_stride(m) >= n_params + 1; the block must be zero before the first step. */
#define HOIR_MAX_PEERS 16
typedef struct hoir_peer_desc {
  int world, rank;
  void* block[HOIR_MAX_PEERS];  /* block[r]: rank r's block as mapped into THIS process            */
} hoir_peer_desc;
long long hoir_peer_stride(const hoir_mlp_desc* m);        /* floats per gradient buffer, or -1   */
long long hoir_peer_block_floats(const hoir_mlp_desc* m);  /* 2 * stride + 64, or -1              */
/* hoir_mlp_train_step_fused with the exchange above; opt->flat_g is ignored (the gradient lives in
 * block[rank], parity opt->step & 1) and opt->keep_grad must be 0.  Every rank must call it for every
 * step with the same opt->step. */
int hoir_mlp_train_step_fused_peer(const hoir_mlp_desc* m, long long B, const float* x,
                                   const hoir_mesh_desc* mesh, int target_kind, const void* target,
                                   float loss_scale, const hoir_opt_desc* opt,
                                   const hoir_peer_desc* peers, void* stream);

/* ---- batches in arbitrary sample order (the reference's DataLoader(shuffle=True),
 * /root/reference/high_order_implicit_representation/single_image_dataset.py:312-320) -------------------------------
 * MSE(mean) and the summed gradients do not depend on the order in which a batch's samples are visited.
 * hoir_mlp_batch_sort writes into order[B] the samples' SOURCE indices sorted by layer-0 cell (segment of x0, segment
 * of x1, low bits of the segments of x2 / x3; a counting sort, 3 small launches):
 *   x != NULL              : sample b = row b of x[B][in];        order holds row indices 0 .. B-1
 *   mesh != NULL           : sample b = pixel pixels[b] of the mesh (pixels == NULL: first_pixel + b);
 *                            order holds absolute pixel indices
 * The *_order training entries then visit sample slot k as source order[k] (row of x / target / gy_ext, or pixel of the
 * mesh / uint8 image): a warp's layer-0 weight reads fall into one or two node windows again and the layer-0 dL/dw is
 * accumulated by the in-kernel run-merged walk, as for contiguous pixel tiles -- no workspace, no second pass.
 * breaks != NULL (HOST pointer): the call synchronises the stream and returns the fraction of (sample, input) pairs
 * whose layer-0 segment differs from the previous sample's in sorted order (pixel coordinates: ~0.03; independent
 * random inputs: ~0.5 -- such batches are better served by hoir_mlp_train_step_fused's two-pass mode).
 * Layer 0 must be piecewise with <= 4 inputs; networks trained by the tensor-core kernel only (HOIR_EUNSUPPORTED
 * otherwise).  order is a DEVICE pointer owned by the caller. */
int hoir_mlp_order_supported(const hoir_mlp_desc* m);          /* 1 / 0: the *_order entries accept this network */
int hoir_mlp_batch_sort(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                        const int* pixels, int* order, float* breaks, void* stream);
/* hoir_mlp_train_step / hoir_mlp_train_step_fused(_peer) with an explicit visiting order (peers may be NULL) */
int hoir_mlp_train_step_order(const hoir_mlp_desc* m, long long B, const float* x, const hoir_mesh_desc* mesh,
                              const int* order, int target_kind, const void* target, const float* gy_ext,
                              const float* packed, float* const* gw, float* loss_sum, float loss_scale, void* stream);
int hoir_mlp_train_step_fused_order(const hoir_mlp_desc* m, long long B, const float* x,
                                    const hoir_mesh_desc* mesh, const int* order, int target_kind,
                                    const void* target, float loss_scale, const hoir_opt_desc* opt,
                                    const hoir_peer_desc* peers, void* stream);

/* ---- optimizers on flat fp32 vectors --------------------------------------------------------
 * torch.optim.Adam (networks.py:94-97) and SparseLion (networks.py:99-102). */
int hoir_adam_step(long long n, float* w, const float* g, float* exp_avg, float* exp_avg_sq,
                   float lr, float beta1, float beta2, float eps, float weight_decay,
                   long long step, void* stream);
int hoir_sparse_lion_step(long long n, float* w, const float* g, float* exp_avg, float lr,
                          float beta1, float beta2, float weight_decay, void* stream);
/* The same updates with the scalars in device memory -- hyper[8] = {lr, beta1, beta2, eps, weight_decay,
 * 1 - beta1^step, sqrt(1 - beta2^step), unused} -- so that a training step captured in a CUDA graph follows the
 * step count and the lr schedule (networks.py:104-118) without re-capture: the host rewrites hyper before a replay. */
int hoir_adam_step_dev(long long n, float* w, const float* g, float* exp_avg, float* exp_avg_sq,
                       const float* hyper, void* stream);
int hoir_sparse_lion_step_dev(long long n, float* w, const float* g, float* exp_avg, const float* hyper,
                              void* stream);
/* dst[0..n) = values[0..n), n <= 8; `values` is HOST memory read before the call returns (the floats travel as
 * kernel arguments), the write is ordered on `stream` */
int hoir_write_floats(float* dst, int n, const float* values, void* stream);

/* FFMA-only micro-kernel: runs `iters` rounds of 128 independent FFMAs per thread on every SM and
 * returns the elapsed milliseconds in *ms and the FLOPs executed in *flops (the measured FP32
 * roofline denominator of bench.py).  reg_form 1: FFMA with three register sources (what a
 * kernel whose weights come from shared memory issues); 0: constant-bank multiplier/addend. */
int hoir_ffma_peak(int iters, int reg_form, float* ms, double* flops, void* stream);
/* tcgen05.mma kind::tf32 micro-kernel (M = 128, N = 256, K = 8 per instruction, operands resident in shared memory,
 * accumulator in TMEM, one CTA per SM): the measured tensor-pipe denominator for the tf32 x 3 kernels. */
int hoir_tf32_peak(int iters, float* ms, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HOIR_H */
