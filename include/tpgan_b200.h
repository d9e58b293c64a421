/*
 * tpgan_b200 — C ABI of the B200 (sm_100a) TP-GAN training-step hot path.
 *
 * This is the drop-in boundary.  The reference (PandaKenWei/TP-GAN) has no FFI: its hot path is a set of
 * ATen calls made from ModificationLayer.py / D_and_G_model.py.  Each entry point below names the reference
 * call site(s) it replaces.  All pointers are raw device pointers owned by the caller (PyTorch's allocator
 * in the shipped host code); the library never allocates or frees device memory on the data path (it owns one fixed
 * scratch allocation made on first use, see tpgan_conv2d).
 * All tensors are fp32, channels-last ("NHWC") views with unit channel stride.  Tensor-core math is TF32
 * (tcgen05.mma kind::tf32, fp32 accumulation in TMEM).
 *
 * Every function returns 0 on success, a negative tpgan_status otherwise; tpgan_last_error() returns a
 * thread-local message.  `stream` is a cudaStream_t passed as void*.  No call synchronises the device.
 */
#ifndef TPGAN_B200_H_
#define TPGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPGAN_ABI_VERSION 3
#define TPGAN_API __attribute__((visibility("default")))

enum tpgan_status {
  TPGAN_OK = 0,
  TPGAN_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  TPGAN_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed (message has the CUDA error) */
  TPGAN_ERR_NO_DEVICE = -3,   /* no sm_100 device */
  TPGAN_ERR_KERNEL_ABORT = -4 /* a kernel's bounded barrier wait expired (see tpgan_kernel_status) */
};

/* NHWC view: element (n,y,x,c) is ptr[n*sn + y*sh + x*sw + c]; strides in elements. A view may be a channel
 * slice of a wider buffer (that is how every torch.cat of D_and_G_model.py:100-104,293-324 is eliminated). */
typedef struct tpgan_view {
  float* ptr;
  int64_t sn, sh, sw;
  int32_t n, h, w, c;
} tpgan_view;

enum tpgan_conv_kind {
  TPGAN_CONV_FWD = 0,     /* nn.Conv2d forward          ModificationLayer.py:101 */
  TPGAN_CONV_DGRAD = 1,   /* its input gradient         (aten::convolution_backward, input half) */
  TPGAN_DECONV_FWD = 2,   /* nn.ConvTranspose2d forward ModificationLayer.py:189 */
  TPGAN_DECONV_DGRAD = 3  /* its input gradient */
};

/* Operand type of the tensor-core kernels.  TF32: activations / packed weights are fp32 storage rounded to tf32
 * (tcgen05.mma kind::tf32).  BF16: the A operand view (`in`; `x` and `dy` of a weight gradient) and the packed weights are
 * bf16 storage - the view's ptr then addresses 2-byte elements and its strides count elements - multiplied with
 * tcgen05.mma kind::f16, fp32 accumulation in TMEM either way.  BASELINE.json configs[2] ("bf16"). */
enum tpgan_dtype {
  TPGAN_DTYPE_TF32 = 0,
  TPGAN_DTYPE_BF16 = 1
};

enum tpgan_epilogue {
  TPGAN_EPI_LINEAR = 0, /* y = v                                   (activation=None, ModificationLayer.py:154) */
  TPGAN_EPI_LEAKY = 1,  /* y = v > 0 ? v : slope*v                 (LeakyReLU(0.01) / ReLU with slope 0)      */
  TPGAN_EPI_MASK = 2    /* y = v * (mask > 0 ? 1 : slope[c])       (activation backward fused into dgrad)     */
};

/* One convolution-like problem, executed as a multi-tap implicit GEMM on tcgen05 tensor cores:
 *   v[pixel, co] = sum_taps sum_ci A_tap[pixel, ci] * Wp[tap][co][ci]   (+ bias[co] + add1 + add2), then epilogue.
 * `in` is the A operand (x for *_FWD, dY for *_DGRAD); `w_packed` must have been produced by
 * tpgan_pack_weights for the same kind.  Supported: kh==kw in {1..8}, stride 1 or 2 for convs (H, W even when
 * stride 2), stride 2 or 4 for deconvs with out = stride*in, width <= 128 of the tile space. */
typedef struct tpgan_conv_args {
  int32_t kind;
  int32_t kh, kw, stride, pad;
  tpgan_view in;
  tpgan_view out;
  const float* w_packed; /* [kh*kw + 1][w_rows_pad][w_k_pad], last tap all zero */
  int32_t w_rows_pad, w_k_pad;
  const float* bias;      /* [out.c] or NULL */
  tpgan_view add1, add2;  /* optional addends with out's geometry; ptr NULL = unused */
  tpgan_view mask;        /* TPGAN_EPI_MASK: tensor whose sign selects 1 or slope */
  const float* slopes;    /* optional per-out-channel negative slopes (overrides `slope`) */
  float slope;
  int32_t epilogue;
  int32_t round_tf32;     /* round the stored result to tf32 (round-to-nearest) */
  int32_t dtype;          /* tpgan_dtype of `in` and `w_packed` ([kh*kw + 1][w_rows_pad][w_k_pad] bf16, w_k_pad % 64 == 0) */
  tpgan_view out16;       /* BF16 only, optional (ptr NULL = unused): a bf16 copy of the result written by the same epilogue
                             (operand storage for the tensor-core consumers of this tensor); ptr addresses 2-byte elements,
                             8-byte aligned, strides in elements.  With out16 given, out.ptr may be NULL (no fp32 copy); out's
                             n/h/w/c must describe the output either way.  bias / add1 / add2 / mask stay fp32. */
  /* fp32-accurate "3xTF32" product in ONE launch (ABI v3; TF32 only, kh*kw*3 <= 64): `in` / `w_packed` hold the tf32-rounded
   * high parts, in_lo / w_lo_packed the residuals (tpgan_split_tf32, tpgan_pack_weights with round_tf32 = 2); the launch
   * accumulates in_hi*w_lo + in_lo*w_hi + in_hi*w_hi in the same TMEM accumulator.  in_lo.ptr NULL = plain product. */
  tpgan_view in_lo;
  const float* w_lo_packed;
} tpgan_conv_args;

/* Runs 1..4 independent problems in ONE persistent launch (the four local pathways of
 * D_and_G_model.py:390-393 are one grouped launch per layer).
 * Library-owned scratch: two kinds of launch use a small ring of device workspace slots that the library allocates once
 * (first call) - Linear-like launches whose reduction is split over the SMs (partial accumulators + per-tile arrival
 * counters; the result is deterministic), and, in tpgan_conv2d_wgrad, weight gradients of <= 4-channel inputs (a
 * zero-padded copy of the input).  A slot is written and read by launches that follow each other in `stream` and is
 * reused 4-8 such launches later; callers that issue these launches concurrently on SEVERAL streams of one process must
 * serialise them.  All launches are issued as programmatic dependents of the preceding kernel in `stream` and order
 * their global-memory accesses themselves (griddepcontrol.wait); no caller-visible change. */
TPGAN_API int tpgan_conv2d(const tpgan_conv_args* groups, int32_t ngroups, void* stream);

/* Weight gradient, accumulated (+=, fp32 atomics) into the forward-packed layout:
 *   conv   : dWp[tap][co][ci] += sum_{n,o} dy[n,o,co] * x[n, stride*o + tap - pad, ci]
 *   deconv : dWp[tap][co][ci] += sum_{n,i} x[n,i,ci]  * dy[n, stride*i + tap - pad, co]
 * (aten::convolution_backward, weight half). */
typedef struct tpgan_wgrad_args {
  int32_t kind; /* TPGAN_CONV_FWD or TPGAN_DECONV_FWD: which layer type the weights belong to */
  int32_t kh, kw, stride, pad;
  tpgan_view x;  /* layer input */
  tpgan_view dy; /* gradient w.r.t. the layer's pre-activation output */
  float* dw_packed;
  int32_t w_rows_pad, w_k_pad;
  int32_t accumulate; /* 1: dW += (atomics); 0: the caller guarantees this is the only launch writing dW since it was
                         cleared, so tiles whose reduction is not split are written with plain vector stores */
  int32_t dtype;      /* tpgan_dtype of x and dy (BF16: 2-byte elements, strides in elements); dw_packed is fp32 either way */
} tpgan_wgrad_args;
TPGAN_API int tpgan_conv2d_wgrad(const tpgan_wgrad_args* groups, int32_t ngroups, void* stream);

/* Weight (re)packing between the reference's parameter layout and the K-major tensor-core layout.
 * ref element for (row r, k index k, tap t) is ref[row_map[r]*ref_row_stride + k_map[k]*ref_k_stride + t]
 * (maps NULL = identity, entry -1 = zero padding).  packed is [taps+1][rows_pad][k_pad] with a zero last tap.
 *   conv  weight (Cout,Cin,kh,kw), FWD  : rows=Cout (stride Cin*taps), k=Cin (stride taps)
 *   conv  weight,                  DGRAD: rows=Cin  (stride taps),     k=Cout (stride Cin*taps)
 *   deconv weight (Cin,Cout,kh,kw), FWD : rows=Cout (stride taps),     k=Cin (stride Cout*taps)
 *   deconv weight,                 DGRAD: rows=Cin  (stride Cout*taps), k=Cout (stride taps)
 * round_tf32: 0 = keep fp32 bits, 1 = round to tf32 (rna), 2 = store the residual w - tf32(w). */
TPGAN_API int tpgan_pack_weights(const float* ref, float* packed, int32_t taps, int32_t rows, int32_t k, int32_t rows_pad,
                       int32_t k_pad, int64_t ref_row_stride, int64_t ref_k_stride, const int32_t* row_map,
                       const int32_t* k_map, int32_t round_tf32, void* stream);
/* Inverse scatter of a packed gradient into the reference layout: ref (+)= packed.  accumulate=0 overwrites. */
TPGAN_API int tpgan_unpack_weights(const float* packed, float* ref, int32_t taps, int32_t rows, int32_t k, int32_t rows_pad,
                         int32_t k_pad, int64_t ref_row_stride, int64_t ref_k_stride, const int32_t* row_map,
                         const int32_t* k_map, int32_t accumulate, void* stream);

/* Per-tap transpose between two packings of the same layer: dst[t][kk][r] = src[t][r][kk] for r < rows, kk < k
 * (forward packing [tap][Cout][Cin] <-> input-gradient packing [tap][Cin][Cout]); padding of dst is left untouched. */
TPGAN_API int tpgan_transpose_packed(const float* src, float* dst, int32_t taps, int32_t rows, int32_t k, int32_t rows_src_pad,
                           int32_t k_src_pad, int32_t rows_dst_pad, int32_t k_dst_pad, void* stream);

/* ---- multi-tensor variants: ONE launch over a device-resident job table (one job per layer) ------------------------
 * block_begin = first block of the job in the launch; jobs sorted by block_begin; total_blocks = sum over jobs. */
typedef struct tpgan_bias_job {   /* db[c] += sum over npix pixels of dy[pix*sw + c]; blocks = pix_blocks * cgroups */
  const float* dy;
  float* db;
  int64_t npix, sw;
  int32_t c, block_begin, pix_blocks, cgroups; /* cgroups = ceil(c / (4 * lanes)) */
  int32_t lanes;                  /* lanes of a warp that share one pixel (power of two <= 32, 4 channels each); the
                                     other 32 / lanes sub-groups of the warp take further pixels */
  int32_t pad_;
} tpgan_bias_job;
typedef struct tpgan_pack_job {   /* row-contiguous (un)pack, see tpgan_pack_weights; blocks = rows_pad (pack) / rows */
  const float* ref_c;             /* reference tensor (read when packing)  */
  float* ref;                     /* same tensor (written when unpacking)  */
  float* packed;
  const int32_t* row_map;
  const int32_t* k_map;
  int64_t rs;                     /* reference row stride = row_len */
  int32_t taps, rows, k, rows_pad, k_pad, row_len, flag /* pack: round mode 0/1/2; unpack: accumulate */, block_begin;
} tpgan_pack_job;
typedef struct tpgan_transpose_job { /* see tpgan_transpose_packed; blocks = taps * tiles_r * tiles_k */
  const float* src;
  float* dst;
  int32_t taps, rows, k, rows_src_pad, k_src_pad, rows_dst_pad, k_dst_pad, block_begin, tiles_k, tiles_r;
} tpgan_transpose_job;
typedef struct tpgan_cast_job {   /* packed fp32 [rows][k_pad] -> bf16 [rows][k_pad16]; blocks = ceil(rows * k_pad16 / 8192) */
  const float* src;
  uint16_t* dst;
  int64_t rows;                   /* (taps + 1) * rows_pad */
  int32_t k_pad, k_pad16;         /* k_pad16 % 64 == 0, >= k_pad; columns beyond k_pad are written as zero */
  int32_t block_begin, pad_;
} tpgan_cast_job;
/* bf16 operand copies of the packed weights of all layers in one launch (after every optimizer step, BF16 mode). */
TPGAN_API int tpgan_cast_packed_multi(const tpgan_cast_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);
/* dst16 (bf16 view of the same logical shape; ptr addresses 2-byte elements, strides in elements) = rne(src): the operand
 * copy of an activation that was not written by a tensor-core epilogue (inputs, pooled / stitched tensors, loss gradients). */
TPGAN_API int tpgan_cast_bf16(tpgan_view src, tpgan_view dst16, void* stream);
TPGAN_API int tpgan_bias_grad_multi(const tpgan_bias_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);
TPGAN_API int tpgan_pack_multi(const tpgan_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, int32_t max_row_len,
                     int32_t unpack, void* stream);
TPGAN_API int tpgan_transpose_multi(const tpgan_transpose_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);

/* ---- HBM-bound kernels of the path ------------------------------------------------------------------- */

/* NCHW (reference tensor layout) <-> NHWC view conversion, optional tf32 rounding on the way in. */
TPGAN_API int tpgan_nchw_to_nhwc(const float* src, tpgan_view dst, int32_t round_tf32, void* stream);
TPGAN_API int tpgan_nhwc_to_nchw(tpgan_view src, float* dst, void* stream);

/* dst = src * (mask > 0 ? 1 : slope[c])  (standalone activation backward; in place allowed). */
TPGAN_API int tpgan_act_backward(tpgan_view src, tpgan_view mask, tpgan_view dst, const float* slopes, float slope,
                       void* stream);
/* dst (+)= src over a view (accumulate != 0 adds). */
TPGAN_API int tpgan_view_copy(tpgan_view src, tpgan_view dst, int32_t accumulate, void* stream);
/* bias gradient: db[c] (+)= sum over pixels of dy[...,c]. */
TPGAN_API int tpgan_bias_grad(tpgan_view dy, float* db, int32_t accumulate, void* stream);

/* nn.ReflectionPad2d((left,right,top,bottom)) forward / backward (ModificationLayer.py:93). */
TPGAN_API int tpgan_reflect_pad(tpgan_view src, tpgan_view dst, int32_t left, int32_t top, void* stream);
TPGAN_API int tpgan_reflect_pad_backward(tpgan_view dpad, tpgan_view dsrc, int32_t left, int32_t top, int32_t accumulate,
                               void* stream);

/* Landmark-centred patch crop, DataAndDataset.py:10-56 (process()).  img: (N,128,128,C) view, landmarks:
 * device float[N][5][2] (x,y) as in process(); writes the four patches and the int32 crop boxes
 * boxes[N][4][4] = (left, upper, right, lower) exactly as passed to PIL.Image.crop.  Out-of-image pixels take
 * `fill` (PIL zero-fill is -1 after the dataset's *2-1 normalisation). */
TPGAN_API int tpgan_patch_crop(tpgan_view img, const float* landmarks, tpgan_view left_eye, tpgan_view right_eye,
                     tpgan_view nose, tpgan_view mouth, int32_t* boxes, float fill, void* stream);

/* LocalFuser, D_and_G_model.py:132-159: out = max over the four zero-padded patches at the fixed offsets.
 * argmax (optional, uint8 [N][128][128][C]) records the winning source 0..3 (first index wins ties, as
 * torch.max over the stacked tensor, so 0 where only zero padding covers the pixel).  Backward routes the
 * gradient to it. */
TPGAN_API int tpgan_local_fuse(tpgan_view left_eye, tpgan_view right_eye, tpgan_view nose, tpgan_view mouth, tpgan_view out,
                     uint8_t* argmax, void* stream);
TPGAN_API int tpgan_local_fuse_backward(tpgan_view dout, const uint8_t* argmax, tpgan_view d_left_eye, tpgan_view d_right_eye,
                              tpgan_view d_nose, tpgan_view d_mouth, int32_t accumulate, void* stream);

/* Identity network (restated ResNet18-128, ResNet.py:28-55): nn.MaxPool2d(3,2,1) (:33) with the window position of the
 * first maximum recorded in argmax (uint8 [N][Ho][Wo][C]) and its backward; nn.AdaptiveAvgPool2d((1,1)) (:45). */
TPGAN_API int tpgan_maxpool3s2(tpgan_view x, tpgan_view y, uint8_t* argmax, void* stream);
TPGAN_API int tpgan_maxpool3s2_backward(tpgan_view dy, const uint8_t* argmax, tpgan_view dx, int32_t accumulate, void* stream);
TPGAN_API int tpgan_avgpool(tpgan_view x, tpgan_view y, void* stream);
TPGAN_API int tpgan_avgpool_backward(tpgan_view dy, tpgan_view dx, int32_t accumulate, void* stream);

/* Fused image losses of the oracle step (config.py:71-82 weights): in one pass over fake and the
 * 128/64/32 targets (TrainDataset keys img_frontal/img64_frontal/img32_frontal, DataAndDataset.py:206-226)
 * computes pixel-L1 at 128/64/32 (fake average-pooled), symmetry-L1 at the same three scales and total variation, and
 * writes d(loss)/d(fake).  sums[8] (device) receives the un-normalised partial sums
 *   {l1_128, l1_64, l1_32, sym_128, sym_64, sym_32, tv_y, tv_x}.
 * w[8] are the final coefficients applied to each term's gradient (weight / element count, host-computed). */
TPGAN_API int tpgan_image_losses(tpgan_view fake, tpgan_view target128, tpgan_view target64, tpgan_view target32,
                       tpgan_view dfake, const float* w, float* sums, void* stream);
/* Patch L1 (local pixel loss): sums[0] += sum|a-b| ; da = coeff*sign(a-b). */
TPGAN_API int tpgan_l1_loss(tpgan_view a, tpgan_view b, tpgan_view da, float coeff, float* sum, void* stream);

/* maxout of nn.MaxPool1d(2,2) over adjacent feature pairs (D_and_G_model.py:214,290) and its backward. */
TPGAN_API int tpgan_maxout2(const float* x, float* y, int32_t rows, int32_t cols_out, void* stream);
TPGAN_API int tpgan_maxout2_backward(const float* x, const float* dy, float* dx, int32_t rows, int32_t cols_out, void* stream);

/* Fused Adam over a flat fp32 parameter bucket (torch.optim.Adam semantics incl. L2 weight_decay). */
TPGAN_API int tpgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, float grad_scale, void* stream);

/* Same update with the step count kept in device memory (*step_dev is incremented first, then used for the bias
 * corrections), so the launch can be captured in a CUDA graph and replayed. */
TPGAN_API int tpgan_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int32_t* step_dev, float grad_scale, void* stream);
/* The same update over a SLICE of the flat buffers; increment != 0 advances the device-resident step count first (exactly one
 * call per optimizer step must do so).  Data-parallel runs update bucket after bucket, each as soon as its all-reduce has
 * finished, so the NCCL transfer of bucket i+1 overlaps the HBM-bound update + re-pack of bucket i. */
TPGAN_API int tpgan_adam_slice_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                   float eps, float weight_decay, int32_t* step_dev, float grad_scale, int32_t increment,
                                   void* stream);

/* Per-sample gradient-penalty helpers: norms[n] = ||g[n]||_2 ; u = coeff[n] * g. */
TPGAN_API int tpgan_sample_sqnorm(tpgan_view g, float* sqnorm, void* stream);
TPGAN_API int tpgan_sample_scale(tpgan_view g, const float* coeff, tpgan_view u, void* stream);
/* coeff[n] = scale * (sqrt(sqnorm[n]) - 1) / sqrt(sqnorm[n]) ; *gp_sum += sum_n (sqrt(sqnorm[n]) - 1)^2 (gp_sum may be NULL). */
TPGAN_API int tpgan_gp_coeff(const float* sqnorm, float* coeff, int32_t n, float scale, float* gp_sum, void* stream);
/* out[n] = alpha[n]*a[n] + (1-alpha[n])*b[n]: the WGAN-GP interpolate x_hat (oracle step, config.py:72). */
TPGAN_API int tpgan_lerp(tpgan_view a, tpgan_view b, const float* alpha, tpgan_view out, void* stream);
/* out = a * b elementwise (nn.Dropout mask of FeaturePredict, D_and_G_model.py:344-346, and its backward). */
TPGAN_API int tpgan_mul(tpgan_view a, tpgan_view b, tpgan_view out, void* stream);
/* hi = tf32(src) (round to nearest), lo = src - hi.  With tpgan_pack_weights(round_tf32 = 2) (which stores w - tf32(w))
 * this gives the fp32-exact verification mode: conv(x, w) = conv(hi, wh) + conv(hi, wl) + conv(lo, wh) up to 2^-21. */
TPGAN_API int tpgan_split_tf32(tpgan_view src, tpgan_view hi, tpgan_view lo, void* stream);
/* out[...] = value. */
TPGAN_API int tpgan_fill(tpgan_view out, float value, void* stream);
/* Cross-entropy of FeaturePredict logits (config.py:82 weight_cross_entropy): *loss_sum += sum_n CE(logits[n], labels[n]);
 * dlogits[n] = coeff * (softmax(logits[n]) - onehot(labels[n])) (dlogits may be NULL).  labels are int64. */
TPGAN_API int tpgan_softmax_ce(const float* logits, int64_t row_stride, const int64_t* labels, float* dlogits,
                               int64_t drow_stride, int32_t rows, int32_t cols, float coeff, float* loss_sum, void* stream);


/* ---- Pretrain path: MobileNetV2 + SSDHead + MultiTaskLoss, MobileNetV2.py:10-534 --------------------- */

/* nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=False) - the depthwise layer of InvertedResidual (MobileNetV2.py:110) -
 * forward, input gradient and weight gradient.  w / dw are the reference tensors (C,1,3,3) themselves (no packing);
 * dw is accumulated with atomics (clear it per step).  C % 4 == 0, stride 1 or 2. */
TPGAN_API int tpgan_dwconv3x3(tpgan_view x, tpgan_view y, const float* w, int32_t stride, void* stream);
TPGAN_API int tpgan_dwconv3x3_dgrad(tpgan_view dy, tpgan_view dx, const float* w, int32_t stride, int32_t accumulate,
                                    void* stream);
TPGAN_API int tpgan_dwconv3x3_wgrad(tpgan_view x, tpgan_view dy, float* dw, int32_t stride, void* stream);

/* nn.BatchNorm2d (+ nn.ReLU6, + the residual add of InvertedResidual.forward, MobileNetV2.py:107-120,150-170) over a
 * pixel-dense NHWC view.  training != 0: batch statistics; sums = caller-owned scratch of 2*C + 1 doubles, ZEROED ONCE by
 * the caller and left zeroed by every call (the last word is a block ticket: the last block of the statistics kernel
 * finalises and clears), running_mean / running_var updated in place with `momentum` and the unbiased variance;
 * training == 0: running statistics.  coef = caller-owned float[4*C] (scale, shift, mean, invstd) consumed by the backward.
 * y = x*scale + shift (+ res) -> act -> tf32 rounding if round_tf32 (operand of a tensor-core conv).  act: 0 none, 1 ReLU6
 * (MobileNetV2; not combinable with res), 2 (Leaky)ReLU with `slope` (ResNet.py:30-31,104-112: BatchNorm -> ReLU, the block's
 * final ReLU after the shortcut add; slope 0 = ReLU). */
TPGAN_API int tpgan_bn_forward(tpgan_view x, tpgan_view res, tpgan_view y, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, int32_t training,
                               int32_t act, float slope, int32_t round_tf32, double* sums, float* coef, void* stream);
/* dz = dy * [0 < y < 6] (if relu6; for act 2 the caller masks dy by the sign of y first - tpgan_act_backward or the conv
 * dgrad epilogue); training: dx (+)= scale*(dz - mean(dz) - xhat*mean(dz*xhat)), dgamma = sum dz*xhat,
 * dbeta = sum dz (overwritten; may be NULL); eval: dx (+)= scale*dz, dgamma / dbeta as in training mode when requested.  dsums = scratch of 2*C + 1 doubles, zeroed once by the
 * caller and left zeroed by every call. */
TPGAN_API int tpgan_bn_backward(tpgan_view dy, tpgan_view x, tpgan_view dx, const float* coef, int32_t training, int32_t relu6,
                                int32_t accumulate, int32_t round_tf32, double* dsums, float* dgamma, float* dbeta,
                                void* stream);

/* SSDHead.forward (MobileNetV2.py:62-76): the NHWC head output (N,h,w,A*K) already is permute(0,2,3,1); view(N,-1,K) and
 * torch.cat(dim=1) are a copy of each image's h*w*A*K floats to `offset` of its concatenated row (row_stride floats).
 * reverse != 0 copies the other way (backward of the cat). */
TPGAN_API int tpgan_rows_gather(tpgan_view v, float* flat, int64_t row_stride, int64_t offset, int32_t reverse, void* stream);

/* MultiTaskLoss (MobileNetV2.py:342-534), batched: per sample, distances of the n predicted points to the 4 ground-truth
 * points, per-label threshold = k_near-th smallest distance, positives = dist <= threshold, label = nearest such ground
 * truth (first wins), location MSE on clamp(./[W,H],0,1) per label, cross-entropy per label + background (class 4) with
 * at most ratio_non_background * positives background points (the ones with the smallest keys u - the explicit form of
 * the reference's torch.multinomial draw); sample loss = alpha*loc + beta*cls; sums[0..2] += coeff*(total, loc, cls);
 * loc_stride / cls_stride = floats between consecutive samples of loc, dloc / cls, dcls (>= 2n / 5n);
 * dloc / dcls = coeff * d(sample loss)/d(.) ; labels[B][n] = assignment (-1 background) for the bit-exact check. */
TPGAN_API int tpgan_multitask_loss(const float* loc, const float* cls, const float* truth, const float* u, int32_t batch,
                                   int32_t n, int64_t loc_stride, int64_t cls_stride, int32_t num_classes, int32_t k_near, float img_w, float img_h, float alpha,
                                   float beta, float ratio_non_background, float coeff, float* dloc, float* dcls,
                                   int32_t* labels, float* sums, void* stream);

/* MultiTaskDecoder.forward (MobileNetV2.py:536-649) for the whole batch: per class, candidates = softmax score >
 * confidence_threshold; greedy distance-NMS in descending score order (a kept point removes the remaining ones within
 * nms_distance); the first top_k kept points are the detections.  Fixed-shape outputs: count[B][K], score[B][K][top_k],
 * point[B][K][top_k][2] (unused slots zero).  accuracy[B] (optional, with truth[B][8]) = _calculate_accuracy
 * (Pretrain.py:17-64) on the top-1 point of the four landmark classes; a class without detection contributes 0. */
TPGAN_API int tpgan_ssd_decode(const float* loc, const float* cls, int32_t batch, int32_t n, int64_t loc_stride,
                               int64_t cls_stride, int32_t num_classes, int32_t top_k, float confidence_threshold,
                               float nms_distance, int32_t* count, float* score, float* point, const float* truth,
                               float* accuracy, void* stream);

/* torch.optim.SGD(momentum, weight_decay, nesterov) over a flat fp32 bucket (getOptimizer 'SGD', UtilityMethods.py:30,
 * config.py:31-35); buf must start zeroed; the learning rate is read from device memory (MultiStepLR, Pretrain.py:117-121,
 * without re-capturing a CUDA graph). */
TPGAN_API int tpgan_sgd_step(float* p, const float* g, float* buf, int64_t n, const float* lr_dev, float momentum,
                             float weight_decay, int32_t nesterov, float grad_scale, void* stream);

/* ---- device-side input pipeline (DataAndDataset.py:179-256, UtilityMethods.py:146-164) ---------------- */

/* transforms.ToTensor() then *2.0 - 1.0 (DataAndDataset.py:214-220,251-255): uint8 HWC bytes [N][H][W][C] -> fp32 NHWC view
 * in [-1, 1], bit-identical to torch's three fp32 operations ((b / 255) * 2 - 1); optional tf32 rounding. */
TPGAN_API int tpgan_u8_to_nhwc(const uint8_t* src, tpgan_view dst, int32_t round_tf32, void* stream);
/* get_5_landmarks_pixal_position (UtilityMethods.py:146-164): out[n][j] = mean of points[n][lo_j..hi_j] (inclusive index
 * ranges, ranges_dev = device int32[nranges][2]; a range outside the list gives NaN, as np.mean of an empty slice does),
 * x scaled by scale_x and y by scale_y (the 128/width, 128/height rescale of TestDataset, DataAndDataset.py:242-245). */
TPGAN_API int tpgan_landmarks_reduce(const float* points, int32_t batch, int32_t npoints, const int32_t* ranges_dev,
                                     int32_t nranges, float scale_x, float scale_y, float* out, void* stream);
/* half = 2x2 and quarter = 4x4 average pool of src in one pass (the 64x64 / 32x32 targets of the oracle step). */
TPGAN_API int tpgan_pyramid(tpgan_view src, tpgan_view half, tpgan_view quarter, void* stream);

/* ---- diagnostics --------------------------------------------------------------------------------------- */
TPGAN_API const char* tpgan_last_error(void);
TPGAN_API int tpgan_abi_version(void);
/* Non-zero after a kernel aborted a barrier wait (code | cta<<8); reading clears it. */
TPGAN_API int tpgan_kernel_status(void);
/* Number of kernels launched by this library since load (monotonic; bench.py reports the delta). */
TPGAN_API int64_t tpgan_launch_count(void);
/* Which tensor-core kernel the calling thread's most recent tpgan_conv2d launched (profiling / bench attribution):
 * 0 = tapgemm_kernel, 1 = rowconv_kernel, 2 = rowstack_kernel. */
TPGAN_API int tpgan_last_conv_kernel(void);
/* 1 when that launch ran tapgemm_kernel over CTA pairs (clusters of 2 CTAs, tcgen05 cta_group::2: one M = 256 MMA per two
 * consecutive 128-pixel tiles, each CTA staging half of the weight tile); 0 otherwise.  TPGAN_PAIR=0 in the environment turns
 * pairing off (A/B measurements). */
TPGAN_API int tpgan_last_conv_pair(void);
/* Deterministic mode (process-wide; returns the previous setting).  By default the split reductions of
 * tpgan_conv2d_wgrad and the multi-block bias sums combine partial results with fp32 atomics, so two runs of the same step
 * agree only to summation order (torch.use_deterministic_algorithms is the reference-side analogue).  With on != 0 every
 * weight-gradient CTA owns whole output tiles and tpgan_bias_grad uses one block per channel group: results are
 * bit-identical from run to run (and between eager launches and CUDA-graph replays), at reduced speed. */
TPGAN_API int tpgan_set_deterministic(int32_t on);
TPGAN_API int tpgan_get_deterministic(void);
/* SMs the persistent tensor-core kernels (one CTA per SM, static tile partition) leave free for concurrent work on other
 * streams - the NCCL all-reduce overlapped with backward (0..64, process-wide; returns the previous value).  Applies to
 * launches issued (or captured into a CUDA graph) after the call. */
TPGAN_API int tpgan_set_sm_reserve(int32_t sms);

#ifdef __cplusplus
}
#endif
#endif /* TPGAN_B200_H_ */
