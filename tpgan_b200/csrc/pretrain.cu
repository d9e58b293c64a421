// Kernels of the Pretrain path (MobileNetV2 + SSDHead + MultiTaskLoss, MobileNetV2.py:10-534).
// The dense 1x1 / 3x3 convolutions of that network run on the tcgen05 kernels (tapgemm.cu, wgrad.cu); what is here is the
// HBM-bound rest: depthwise 3x3 convolutions (fwd / dgrad / wgrad, CUDA cores - 13.5 MFLOP per image, pure streaming),
// training-mode BatchNorm (+ReLU6, +residual add) forward / backward, the SSD head gather, the batched MultiTaskLoss
// (assignment + loss + gradients, one CTA per sample) and the Nesterov-SGD update.
// All activations are fp32 NHWC "pixel-dense" views: element (pixel p, channel c) at ptr[p * ld + c], C % 4 == 0.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdint.h>

#include "../../include/tpgan_b200.h"
#include "common.cuh"
#include "host_common.h"

namespace tpg {

static inline int grid_cap(long long want, int per_sm) {
  long long cap = (long long)std::max(1, device_sm_count() ? device_sm_count() : 148) * per_sm;
  return (int)std::max(1ll, std::min(want, cap));
}

struct PV {  // device copy of tpgan_view
  float* p;
  long long sn, sh, sw;
  int n, h, w, c;
};
static inline PV pv(const tpgan_view& v) { return PV{v.ptr, v.sn, v.sh, v.sw, v.n, v.h, v.w, v.c}; }
__device__ __forceinline__ long long poff(const PV& v, int n, int y, int x) {
  return (long long)n * v.sn + (long long)y * v.sh + (long long)x * v.sw;
}
static bool dense(const tpgan_view& v) {
  return v.ptr && v.c > 0 && v.c % 4 == 0 && v.sw % 4 == 0 && v.sh == (int64_t)v.w * v.sw && v.sn == (int64_t)v.h * v.sh &&
         (((uintptr_t)v.ptr) & 15) == 0;
}
static bool vec_view(const tpgan_view& v) {
  return v.ptr && v.c > 0 && v.c % 4 == 0 && v.sw % 4 == 0 && v.sh % 4 == 0 && v.sn % 4 == 0 && (((uintptr_t)v.ptr) & 15) == 0;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 rnd4(float4 v) {
  return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}

// ------------------------------------------------------------------------------------------- depthwise 3x3 convolution
// nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=False) (MobileNetV2.py:110).  w is the reference tensor (C,1,3,3) = [C][9];
// it is staged transposed ([9][C]) in shared memory so that a thread's four channels are one 16-byte read per tap.
__global__ void __launch_bounds__(256) dw3x3_fwd_kernel(PV x, PV y, const float* __restrict__ w, int stride) {
  extern __shared__ float ws[];
  const int C = x.c;
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) ws[(i % 9) * C + i / 9] = w[i];
  __syncthreads();
  const int cqn = C >> 2;
  const long long total = (long long)y.n * y.h * y.w * cqn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cqn);
    long long r = i / cqn;
    const int ox = (int)(r % y.w);
    r /= y.w;
    const int oy = (int)(r % y.h);
    const int n = (int)(r / y.h);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int iy = oy * stride - 1 + kr;
      if (iy < 0 || iy >= x.h) continue;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int ix = ox * stride - 1 + ks;
        if (ix < 0 || ix >= x.w) continue;
        const float4 v = ld4(x.p + poff(x, n, iy, ix) + q * 4);
        const float4 k = ld4(ws + (kr * 3 + ks) * C + q * 4);
        acc.x = fmaf(v.x, k.x, acc.x);
        acc.y = fmaf(v.y, k.y, acc.y);
        acc.z = fmaf(v.z, k.z, acc.z);
        acc.w = fmaf(v.w, k.w, acc.w);
      }
    }
    st4(y.p + poff(y, n, oy, ox) + q * 4, acc);
  }
}

// dx[n,iy,ix,c] (+)= sum_{kr,ks : (iy+1-kr) % s == 0, (ix+1-ks) % s == 0} dy[n,(iy+1-kr)/s,(ix+1-ks)/s,c] * w[c,kr,ks]
__global__ void __launch_bounds__(256) dw3x3_dgrad_kernel(PV dy, PV dx, const float* __restrict__ w, int stride, int accumulate) {
  extern __shared__ float ws[];
  const int C = dx.c;
  for (int i = threadIdx.x; i < C * 9; i += blockDim.x) ws[(i % 9) * C + i / 9] = w[i];
  __syncthreads();
  const int cqn = C >> 2;
  const long long total = (long long)dx.n * dx.h * dx.w * cqn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cqn);
    long long r = i / cqn;
    const int ix = (int)(r % dx.w);
    r /= dx.w;
    const int iy = (int)(r % dx.h);
    const int n = (int)(r / dx.h);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int ty = iy + 1 - kr;
      if (ty < 0 || (ty % stride) != 0) continue;
      const int oy = ty / stride;
      if (oy >= dy.h) continue;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int tx = ix + 1 - ks;
        if (tx < 0 || (tx % stride) != 0) continue;
        const int ox = tx / stride;
        if (ox >= dy.w) continue;
        const float4 g = ld4(dy.p + poff(dy, n, oy, ox) + q * 4);
        const float4 k = ld4(ws + (kr * 3 + ks) * C + q * 4);
        acc.x = fmaf(g.x, k.x, acc.x);
        acc.y = fmaf(g.y, k.y, acc.y);
        acc.z = fmaf(g.z, k.z, acc.z);
        acc.w = fmaf(g.w, k.w, acc.w);
      }
    }
    float* d = dx.p + poff(dx, n, iy, ix) + q * 4;
    if (accumulate) {
      const float4 o = ld4(d);
      acc.x += o.x, acc.y += o.y, acc.z += o.z, acc.w += o.w;
    }
    st4(d, acc);
  }
}

// dw[c][kr*3+ks] += sum_{n,oy,ox} dy[n,oy,ox,c] * x[n, oy*s-1+kr, ox*s-1+ks, c]   (atomics into the reference layout).
// block (32, 8): lane = channel quad (a warp reads 512 contiguous bytes of a pixel), ty = pixel sub-group.
__global__ void __launch_bounds__(256) dw3x3_wgrad_kernel(PV x, PV dy, float* __restrict__ dw, int stride) {
  __shared__ float sacc[36][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty * 32 + tx; i < 36 * 33; i += 256) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  const int cqn = x.c >> 2;
  const int q = blockIdx.y * 32 + tx;
  if (q < cqn) {
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long npix = (long long)dy.n * dy.h * dy.w;
    for (long long p = (long long)blockIdx.x * 8 + ty; p < npix; p += (long long)gridDim.x * 8) {
      const int ox = (int)(p % dy.w);
      long long r = p / dy.w;
      const int oy = (int)(r % dy.h);
      const int n = (int)(r / dy.h);
      const float4 g = ld4(dy.p + poff(dy, n, oy, ox) + q * 4);
#pragma unroll
      for (int kr = 0; kr < 3; ++kr) {
        const int iy = oy * stride - 1 + kr;
        if (iy < 0 || iy >= x.h) continue;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const int ix = ox * stride - 1 + ks;
          if (ix < 0 || ix >= x.w) continue;
          const float4 v = ld4(x.p + poff(x, n, iy, ix) + q * 4);
          float4& a = acc[kr * 3 + ks];
          a.x = fmaf(g.x, v.x, a.x);
          a.y = fmaf(g.y, v.y, a.y);
          a.z = fmaf(g.z, v.z, a.z);
          a.w = fmaf(g.w, v.w, a.w);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      atomicAdd(&sacc[t * 4 + 0][tx], acc[t].x);
      atomicAdd(&sacc[t * 4 + 1][tx], acc[t].y);
      atomicAdd(&sacc[t * 4 + 2][tx], acc[t].z);
      atomicAdd(&sacc[t * 4 + 3][tx], acc[t].w);
    }
  }
  __syncthreads();
  for (int i = ty * 32 + tx; i < 36 * 32; i += 256) {
    const int lane = i & 31, e = i >> 5;       // e = tap*4 + channel-in-quad
    const int t = e >> 2, j = e & 3;
    const int c = (blockIdx.y * 32 + lane) * 4 + j;
    if (c < x.c) atomicAdd(dw + (long long)c * 9 + t, sacc[e][lane]);
  }
}

// ------------------------------------------------------------------------------------------- BatchNorm (training mode)
// nn.BatchNorm2d (MobileNetV2.py:107,111,115,152,168) over a pixel-dense [M][ld] matrix.
//   stats    : sums[0][c] += sum_p x, sums[1][c] += sum_p x^2   (fp32 per-thread partials of <= ~32 elements, then fp64)
//   finalize : coef[0..3][c] = scale, shift, mean, invstd ; running statistics updated (momentum, unbiased variance)
//   apply    : y = x*scale + shift (+ residual) -> optional ReLU6 -> optional tf32 rounding
//   backward : dz = dy * [0 < y < 6] ; dsums = (sum dz, sum dz*xhat) ; dx = scale * (dz - mean(dz) - xhat * mean(dz*xhat))
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long M, long long ld, int C,
                                                      double* __restrict__ sums) {
  __shared__ double red[8][32][8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int q = blockIdx.y * 32 + tx;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  if (q * 4 < C) {
    for (long long r = (long long)blockIdx.x * 8 + ty; r < M; r += (long long)gridDim.x * 8) {
      const float4 v = ld4(x + r * ld + q * 4);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      ss.x = fmaf(v.x, v.x, ss.x), ss.y = fmaf(v.y, v.y, ss.y), ss.z = fmaf(v.z, v.z, ss.z), ss.w = fmaf(v.w, v.w, ss.w);
    }
  }
  double* d = red[ty][tx];
  d[0] = s.x, d[1] = s.y, d[2] = s.z, d[3] = s.w, d[4] = ss.x, d[5] = ss.y, d[6] = ss.z, d[7] = ss.w;
  __syncthreads();
  // 256 threads finish the 32 x 8 outputs of the block: thread -> (lane = tid & 31, component = tid >> 5)
  const int tid = ty * 32 + tx, lane = tid & 31, e = tid >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4 + (e & 3);
  if (c < C) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += red[k][lane][e];
    atomicAdd(sums + (e >> 2) * C + c, a);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, int training,
                                   float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (training) {
    mean = sums[c] / (double)M;
    var = sums[C + c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) {
      const double unb = (M > 1) ? var * (double)M / (double)(M - 1) : var;
      running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
      running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unb);
    }
  } else {
    mean = (double)running_mean[c];
    var = (double)running_var[c];
  }
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float scale = gamma[c] * invstd;
  coef[c] = scale;
  coef[C + c] = beta[c] - (float)mean * scale;
  coef[2 * C + c] = (float)mean;
  coef[3 * C + c] = invstd;
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, long long M, long long ldx, int C,
                                                      const float* __restrict__ coef, const float* __restrict__ res,
                                                      long long ldr, float* __restrict__ y, long long ldy, int relu6,
                                                      int round) {
  extern __shared__ float sc[];  // scale[C], shift[C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sc[i] = coef[i];
  __syncthreads();
  const int cqn = C >> 2;
  const long long total = M * cqn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cqn);
    const long long r = i / cqn;
    const float4 v = ld4(x + r * ldx + q * 4);
    const float4 a = ld4(sc + q * 4), b = ld4(sc + C + q * 4);
    float4 o = make_float4(fmaf(v.x, a.x, b.x), fmaf(v.y, a.y, b.y), fmaf(v.z, a.z, b.z), fmaf(v.w, a.w, b.w));
    if (res) {
      const float4 e = ld4(res + r * ldr + q * 4);
      o.x += e.x, o.y += e.y, o.z += e.z, o.w += e.w;
    }
    if (relu6) {
      o.x = fminf(fmaxf(o.x, 0.f), 6.f), o.y = fminf(fmaxf(o.y, 0.f), 6.f);
      o.z = fminf(fmaxf(o.z, 0.f), 6.f), o.w = fminf(fmaxf(o.w, 0.f), 6.f);
    }
    if (round) o = rnd4(o);
    st4(y + r * ldy + q * 4, o);
  }
}

__device__ __forceinline__ float relu6_gate(float x, float scale, float shift, float g, int relu6) {
  if (!relu6) return g;
  const float y = fmaf(x, scale, shift);   // the very expression of bn_apply_kernel: identical mask bit for bit
  return (y > 0.f && y < 6.f) ? g : 0.f;
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, long long ldd,
                                                           const float* __restrict__ x, long long ldx, long long M, int C,
                                                           const float* __restrict__ coef, int relu6,
                                                           double* __restrict__ dsums) {
  __shared__ double red[8][32][8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int q = blockIdx.y * 32 + tx;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (q * 4 < C) {
    const float4 a = ld4(coef + q * 4), b = ld4(coef + C + q * 4), mu = ld4(coef + 2 * C + q * 4),
                 is = ld4(coef + 3 * C + q * 4);
    for (long long r = (long long)blockIdx.x * 8 + ty; r < M; r += (long long)gridDim.x * 8) {
      const float4 v = ld4(x + r * ldx + q * 4);
      const float4 g = ld4(dy + r * ldd + q * 4);
      const float gx = relu6_gate(v.x, a.x, b.x, g.x, relu6), gy = relu6_gate(v.y, a.y, b.y, g.y, relu6),
                  gz = relu6_gate(v.z, a.z, b.z, g.z, relu6), gw = relu6_gate(v.w, a.w, b.w, g.w, relu6);
      s1.x += gx, s1.y += gy, s1.z += gz, s1.w += gw;
      s2.x = fmaf(gx, (v.x - mu.x) * is.x, s2.x), s2.y = fmaf(gy, (v.y - mu.y) * is.y, s2.y);
      s2.z = fmaf(gz, (v.z - mu.z) * is.z, s2.z), s2.w = fmaf(gw, (v.w - mu.w) * is.w, s2.w);
    }
  }
  double* d = red[ty][tx];
  d[0] = s1.x, d[1] = s1.y, d[2] = s1.z, d[3] = s1.w, d[4] = s2.x, d[5] = s2.y, d[6] = s2.z, d[7] = s2.w;
  __syncthreads();
  const int tid = ty * 32 + tx, lane = tid & 31, e = tid >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4 + (e & 3);
  if (c < C) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += red[k][lane][e];
    atomicAdd(dsums + (e >> 2) * C + c, acc);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, long long ldd,
                                                          const float* __restrict__ x, long long ldx, long long M, int C,
                                                          const float* __restrict__ coef, const double* __restrict__ dsums,
                                                          int relu6, int training, float* __restrict__ dx, long long ldo,
                                                          int accumulate, int round, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta) {
  extern __shared__ float sm[];  // c1[C] = mean(dz), c2[C] = mean(dz*xhat)
  float* c1 = sm;
  float* c2 = sm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double a = training ? dsums[c] : 0.0, b = training ? dsums[C + c] : 0.0;
    c1[c] = (float)(a / (double)M);
    c2[c] = (float)(b / (double)M);
    if (training && blockIdx.x == 0) {
      if (dbeta) dbeta[c] = (float)a;
      if (dgamma) dgamma[c] = (float)b;
    }
  }
  __syncthreads();
  const int cqn = C >> 2;
  const long long total = M * cqn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cqn);
    const long long r = i / cqn;
    const float4 a = ld4(coef + q * 4), b = ld4(coef + C + q * 4), mu = ld4(coef + 2 * C + q * 4),
                 is = ld4(coef + 3 * C + q * 4);
    const float4 v = ld4(x + r * ldx + q * 4);
    const float4 g = ld4(dy + r * ldd + q * 4);
    const float4 m1 = ld4(c1 + q * 4), m2 = ld4(c2 + q * 4);
    float4 o;
    o.x = a.x * (relu6_gate(v.x, a.x, b.x, g.x, relu6) - m1.x - (v.x - mu.x) * is.x * m2.x);
    o.y = a.y * (relu6_gate(v.y, a.y, b.y, g.y, relu6) - m1.y - (v.y - mu.y) * is.y * m2.y);
    o.z = a.z * (relu6_gate(v.z, a.z, b.z, g.z, relu6) - m1.z - (v.z - mu.z) * is.z * m2.z);
    o.w = a.w * (relu6_gate(v.w, a.w, b.w, g.w, relu6) - m1.w - (v.w - mu.w) * is.w * m2.w);
    float* d = dx + r * ldo + q * 4;
    if (accumulate) {
      const float4 p = ld4(d);
      o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
    }
    if (round) o = rnd4(o);
    st4(d, o);
  }
}

// ------------------------------------------------------------------------------------------------- SSD head gather
// SSDHead.forward (MobileNetV2.py:62-76): conv output (N,h,w,A*K) NHWC *is* permute(0,2,3,1); view(N,-1,K) + cat(dim 1)
// = copy of each image's h*w*A*K floats to offset `off` of the concatenated row.  reverse: flat row -> view (backward).
__global__ void rows_gather_kernel(PV v, float* __restrict__ flat, long long row_stride, long long off, int reverse) {
  const long long per = (long long)v.h * v.w * v.c;
  const long long total = per * v.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    const long long j = i - (long long)n * per;
    const int c = (int)(j % v.c);
    long long r = j / v.c;
    const int x = (int)(r % v.w);
    const int y = (int)(r / v.w);
    float* a = v.p + poff(v, n, y, x) + c;
    float* b = flat + n * row_stride + off + j;
    if (reverse)
      *a = *b;
    else
      *b = *a;
  }
}

// ------------------------------------------------------------------------------------------------- MultiTaskLoss
// MobileNetV2.py:360-534, one CTA per sample (oracle/pretrain_port.py states the batched semantics).  Distances use
// explicitly un-contracted fp32 operations so that the assignment is bit-identical to the CPU restatement.
struct MtlArgs {
  const float* loc;    // [B][n][2]
  const float* cls;    // [B][n][K]
  const float* truth;  // [B][8]
  const float* u;      // [B][n] background sub-sampling keys, may be NULL (then no sub-sampling may be needed)
  float* dloc;         // [B][n][2] or NULL
  float* dcls;         // [B][n][K] or NULL
  int* labels;         // [B][n] or NULL
  float* sums;         // [3]: total, location, classification (each already multiplied by coeff)
  long long loc_stride, cls_stride;   // floats between consecutive samples of loc/dloc and cls/dcls
  int n, K, k_near;
  float img_w, img_h, alpha, beta, ratio_nb, coeff;
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += scratch[i];   // same order in every thread: deterministic
  return t;
}

__global__ void __launch_bounds__(256) multitask_loss_kernel(MtlArgs A) {
  extern __shared__ float smf[];
  const int n = A.n, K = A.K, b = blockIdx.x, tid = threadIdx.x;
  float* lx = smf;
  float* ly = lx + n;
  float* dist = ly + n;            // [4][n]
  float* key = dist + 4 * n;       // [n]
  int* lab = reinterpret_cast<int*>(key + n);
  int* sel = lab + n;
  __shared__ float thr[4], tx[4], ty[4], scratch[8];
  __shared__ int cnt[5];           // per-label positives, [4] = background points
  if (tid < 4) {
    tx[tid] = A.truth[b * 8 + tid * 2];
    ty[tid] = A.truth[b * 8 + tid * 2 + 1];
  }
  if (tid < 5) cnt[tid] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const float px = A.loc[b * A.loc_stride + i * 2], py = A.loc[b * A.loc_stride + i * 2 + 1];
    lx[i] = px, ly[i] = py;
    key[i] = A.u ? A.u[(long long)b * n + i] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dx = __fsub_rn(px, tx[j]), dy = __fsub_rn(py, ty[j]);
      dist[j * n + i] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
  }
  __syncthreads();
  // threshold of label j = k-th smallest distance (topk(k, largest=False)[0].max(), MobileNetV2.py:399-401)
  for (int j = 0; j < 4; ++j) {
    const float* d = dist + j * n;
    for (int i = tid; i < n; i += blockDim.x) {
      const float v = d[i];
      int lt = 0, le = 0;
      for (int m = 0; m < n; ++m) {
        const float o = d[m];
        lt += (o < v);
        le += (o <= v);
      }
      if (lt <= A.k_near - 1 && A.k_near - 1 < le) thr[j] = v;   // every writer writes the same value
    }
  }
  __syncthreads();
  // label = nearest label among those the point is positive for; strict '<', first label wins (:420-430)
  for (int i = tid; i < n; i += blockDim.x) {
    float best = INFINITY;
    int l = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = dist[j * n + i];
      if (v <= thr[j] && v < best) best = v, l = j;
    }
    lab[i] = l;
    if (A.labels) A.labels[(long long)b * n + i] = l;
    atomicAdd(&cnt[l < 0 ? 4 : l], 1);
  }
  __syncthreads();
  const int n_pos = cnt[0] + cnt[1] + cnt[2] + cnt[3], n_bg = cnt[4];
  const int max_bg = (int)((double)n_pos * (double)A.ratio_nb);   // int(count * ratio), :499
  const bool sub = n_bg > max_bg;
  const int n_sel = sub ? max_bg : n_bg;
  // background sub-sampling: the max_bg background points with the smallest keys (ties: lower index)
  for (int i = tid; i < n; i += blockDim.x) {
    int s = 0;
    if (lab[i] < 0) {
      if (!sub) {
        s = 1;
      } else {
        const float v = key[i];
        int rank = 0;
        for (int m = 0; m < n; ++m) rank += (lab[m] < 0) && (key[m] < v || (key[m] == v && m < i));
        s = rank < max_bg;
      }
    }
    sel[i] = s;
  }
  __syncthreads();
  float loc_part = 0.f, cls_part = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    const int l = lab[i];
    const long long li = b * A.loc_stride + (long long)i * 2, ci = b * A.cls_stride + (long long)i * K;
    float gx = 0.f, gy = 0.f;
    if (l >= 0) {
      // clamp(pred / [W, H], 0, 1) vs clamp(true / [W, H], 0, 1), nn.MSELoss mean over (count, 2) (:457-481)
      const float qx = lx[i] / A.img_w, qy = ly[i] / A.img_h;
      const float cx = fminf(fmaxf(qx, 0.f), 1.f), cy = fminf(fmaxf(qy, 0.f), 1.f);
      const float ex = cx - fminf(fmaxf(tx[l] / A.img_w, 0.f), 1.f), ey = cy - fminf(fmaxf(ty[l] / A.img_h, 0.f), 1.f);
      const float inv = 1.f / (2.f * (float)cnt[l]);
      loc_part += (ex * ex + ey * ey) * inv;
      const float gsc = A.coeff * A.alpha * 2.f * inv;
      gx = (qx >= 0.f && qx <= 1.f) ? gsc * ex / A.img_w : 0.f;
      gy = (qy >= 0.f && qy <= 1.f) ? gsc * ey / A.img_h : 0.f;
    }
    if (A.dloc) A.dloc[li] = gx, A.dloc[li + 1] = gy;
    const int target = l >= 0 ? l : (sel[i] ? 4 : -1);
    const float* z = A.cls + ci;
    if (target >= 0) {
      const float wgt = 1.f / (float)(l >= 0 ? cnt[l] : n_sel);
      float mx = z[0];
      for (int k = 1; k < K; ++k) mx = fmaxf(mx, z[k]);
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += expf(z[k] - mx);
      const float lse = mx + logf(s);
      cls_part += (lse - z[target]) * wgt;
      if (A.dcls) {
        const float gsc = A.coeff * A.beta * wgt;
        for (int k = 0; k < K; ++k) A.dcls[ci + k] = gsc * (expf(z[k] - lse) - (k == target ? 1.f : 0.f));
      }
    } else if (A.dcls) {
      for (int k = 0; k < K; ++k) A.dcls[ci + k] = 0.f;
    }
  }
  const float loc_loss = block_sum(loc_part, scratch);
  const float cls_loss = block_sum(cls_part, scratch);
  if (tid == 0) {
    atomicAdd(A.sums + 0, A.coeff * (A.alpha * loc_loss + A.beta * cls_loss));
    atomicAdd(A.sums + 1, A.coeff * loc_loss);
    atomicAdd(A.sums + 2, A.coeff * cls_loss);
  }
}

// ------------------------------------------------------------------------------------------------- SGD (Nesterov)
// torch.optim.SGD(lr, momentum, weight_decay, nesterov=True), UtilityMethods.py:30 / config.py:31-35, dampening 0:
//   g = grad_scale*g + wd*p ; buf = momentum*buf + g (a zero-initialised buf reproduces torch's first step buf = g) ;
//   p -= lr * (g + momentum*buf)   [nesterov]   or   p -= lr * buf.   lr is read from device memory (graph-replay safe).
__global__ void sgd_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ buf, long long n4,
                           const float* __restrict__ lr_dev, float momentum, float wd, int nesterov, float grad_scale) {
  const float lr = *lr_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 w = p[i], m = buf[i];
    const float4 gr = g[i];
    float* wp = &w.x;
    float* mp = &m.x;
    const float* gp = &gr.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = fmaf(wd, wp[j], grad_scale * gp[j]);
      mp[j] = fmaf(momentum, mp[j], gg);
      const float step = nesterov ? fmaf(momentum, mp[j], gg) : mp[j];
      wp[j] = wp[j] - lr * step;
    }
    p[i] = w;
    buf[i] = m;
  }
}

}  // namespace tpg

using namespace tpg;
#define ST ((cudaStream_t)stream)

extern "C" {

int tpgan_dwconv3x3(tpgan_view x, tpgan_view y, const float* w, int32_t stride, void* stream) {
  if (!vec_view(x) || !vec_view(y) || !w || x.c != y.c || x.n != y.n || (stride != 1 && stride != 2) ||
      y.h != (x.h + 2 - 3) / stride + 1 || y.w != (x.w + 2 - 3) / stride + 1 || x.c * 36 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3: bad geometry (C %% 4 == 0, C <= 1365, stride 1|2, pad 1)");
  const long long total = (long long)y.n * y.h * y.w * (y.c / 4);
  dw3x3_fwd_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)x.c * 36, ST>>>(pv(x), pv(y), w, stride);
  TPG_CHECK_LAUNCH("dwconv3x3");
  return 0;
}

int tpgan_dwconv3x3_dgrad(tpgan_view dy, tpgan_view dx, const float* w, int32_t stride, int32_t accumulate, void* stream) {
  if (!vec_view(dx) || !vec_view(dy) || !w || dx.c != dy.c || dx.n != dy.n || (stride != 1 && stride != 2) ||
      dy.h != (dx.h + 2 - 3) / stride + 1 || dy.w != (dx.w + 2 - 3) / stride + 1 || dx.c * 36 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3_dgrad: bad geometry");
  const long long total = (long long)dx.n * dx.h * dx.w * (dx.c / 4);
  dw3x3_dgrad_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)dx.c * 36, ST>>>(pv(dy), pv(dx), w, stride,
                                                                                     accumulate);
  TPG_CHECK_LAUNCH("dwconv3x3_dgrad");
  return 0;
}

int tpgan_dwconv3x3_wgrad(tpgan_view x, tpgan_view dy, float* dw, int32_t stride, void* stream) {
  if (!vec_view(x) || !vec_view(dy) || !dw || x.c != dy.c || x.n != dy.n || (stride != 1 && stride != 2) ||
      dy.h != (x.h + 2 - 3) / stride + 1 || dy.w != (x.w + 2 - 3) / stride + 1)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3_wgrad: bad geometry");
  const long long npix = (long long)dy.n * dy.h * dy.w;
  dim3 grid((unsigned)grid_cap((npix + 63) / 64, 2), (unsigned)((x.c / 4 + 31) / 32));
  dw3x3_wgrad_kernel<<<grid, dim3(32, 8), 0, ST>>>(pv(x), pv(dy), dw, stride);
  TPG_CHECK_LAUNCH("dwconv3x3_wgrad");
  return 0;
}

int tpgan_bn_forward(tpgan_view x, tpgan_view res, tpgan_view y, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int32_t training, int32_t relu6, int32_t round_tf32,
                     double* sums, float* coef, void* stream) {
  if (!dense(x) || !dense(y) || x.c != y.c || x.n != y.n || x.h != y.h || x.w != y.w || !gamma || !beta || !coef ||
      (training && !sums) || (!training && (!running_mean || !running_var)) || x.c * 8 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "bn_forward: bad arguments (pixel-dense views, C %% 4 == 0)");
  if (res.ptr && (!dense(res) || res.c != x.c || res.n != x.n || res.h != x.h || res.w != x.w || relu6))
    return set_error(TPGAN_ERR_INVALID, "bn_forward: residual must match x and excludes ReLU6");
  const long long M = (long long)x.n * x.h * x.w;
  const int C = x.c;
  if (training) {
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, ST);
    if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)grid_cap((M + 255) / 256, 4), (unsigned)((C / 4 + 31) / 32));
    bn_stats_kernel<<<grid, dim3(32, 8), 0, ST>>>(x.ptr, M, x.sw, C, sums);
    TPG_CHECK_LAUNCH("bn_stats");
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(sums, M, C, gamma, beta, running_mean, running_var, momentum, eps,
                                                     training, coef);
  TPG_CHECK_LAUNCH("bn_finalize");
  const long long total = M * (C / 4);
  bn_apply_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)C * 8, ST>>>(x.ptr, M, x.sw, C, coef, res.ptr, res.sw,
                                                                               y.ptr, y.sw, relu6, round_tf32);
  TPG_CHECK_LAUNCH("bn_apply");
  return 0;
}

int tpgan_bn_backward(tpgan_view dy, tpgan_view x, tpgan_view dx, const float* coef, int32_t training, int32_t relu6,
                      int32_t accumulate, int32_t round_tf32, double* dsums, float* dgamma, float* dbeta, void* stream) {
  if (!dense(x) || !dense(dy) || !dense(dx) || x.c != dy.c || x.c != dx.c || x.n != dy.n || x.h != dy.h || x.w != dy.w ||
      x.n != dx.n || x.h != dx.h || x.w != dx.w || !coef || (training && !dsums) || x.c * 8 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "bn_backward: bad arguments");
  const long long M = (long long)x.n * x.h * x.w;
  const int C = x.c;
  if (training) {
    cudaError_t e = cudaMemsetAsync(dsums, 0, sizeof(double) * 2 * C, ST);
    if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    dim3 grid((unsigned)grid_cap((M + 255) / 256, 4), (unsigned)((C / 4 + 31) / 32));
    bn_bwd_reduce_kernel<<<grid, dim3(32, 8), 0, ST>>>(dy.ptr, dy.sw, x.ptr, x.sw, M, C, coef, relu6, dsums);
    TPG_CHECK_LAUNCH("bn_bwd_reduce");
  }
  const long long total = M * (C / 4);
  bn_bwd_apply_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)C * 8, ST>>>(
      dy.ptr, dy.sw, x.ptr, x.sw, M, C, coef, dsums, relu6, training, dx.ptr, dx.sw, accumulate, round_tf32, dgamma, dbeta);
  TPG_CHECK_LAUNCH("bn_bwd_apply");
  return 0;
}

int tpgan_rows_gather(tpgan_view v, float* flat, int64_t row_stride, int64_t offset, int32_t reverse, void* stream) {
  if (!v.ptr || !flat || v.n < 1 || offset < 0 || offset + (int64_t)v.h * v.w * v.c > row_stride)
    return set_error(TPGAN_ERR_INVALID, "rows_gather: bad arguments");
  const long long total = (long long)v.n * v.h * v.w * v.c;
  rows_gather_kernel<<<grid_cap((total + 255) / 256, 8), 256, 0, ST>>>(pv(v), flat, row_stride, offset, reverse);
  TPG_CHECK_LAUNCH("rows_gather");
  return 0;
}

int tpgan_multitask_loss(const float* loc, const float* cls, const float* truth, const float* u, int32_t batch, int32_t n,
                         int64_t loc_stride, int64_t cls_stride, int32_t num_classes, int32_t k_near, float img_w, float img_h, float alpha, float beta,
                         float ratio_non_background, float coeff, float* dloc, float* dcls, int32_t* labels, float* sums,
                         void* stream) {
  if (!loc || !cls || !truth || !sums || batch < 1 || n < 1 || n > 1024 || num_classes != 5 || k_near < 1 || k_near > n ||
      loc_stride < 2ll * n || cls_stride < (long long)num_classes * n)
    return set_error(TPGAN_ERR_INVALID, "multitask_loss: bad arguments (1 <= n <= 1024, 5 classes, 1 <= k <= n)");
  MtlArgs A{loc, cls, truth, u, dloc, dcls, labels, sums, loc_stride, cls_stride, n, num_classes, k_near, img_w, img_h, alpha, beta,
            ratio_non_background, coeff};
  multitask_loss_kernel<<<batch, 256, (size_t)n * 9 * 4, ST>>>(A);
  TPG_CHECK_LAUNCH("multitask_loss");
  return 0;
}

int tpgan_sgd_step(float* p, const float* g, float* buf, int64_t n, const float* lr_dev, float momentum, float weight_decay,
                   int32_t nesterov, float grad_scale, void* stream) {
  if (!p || !g || !buf || !lr_dev || n <= 0 || (n % 4) || ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf) & 15) != 0))
    return set_error(TPGAN_ERR_INVALID, "sgd_step: flat buffers must be 16-byte aligned with n %% 4 == 0");
  sgd_kernel<<<grid_cap((n / 4 + 255) / 256, 16), 256, 0, ST>>>(reinterpret_cast<float4*>(p),
                                                              reinterpret_cast<const float4*>(g),
                                                              reinterpret_cast<float4*>(buf), n / 4, lr_dev, momentum,
                                                              weight_decay, nesterov, grad_scale);
  TPG_CHECK_LAUNCH("sgd_step");
  return 0;
}

}  // extern "C"
