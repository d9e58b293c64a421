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
#include <stdlib.h>

#include "../../include/tpgan_b200.h"
#include "common.cuh"
#include "host_common.h"

namespace tpg {

static inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
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
// nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=False) (MobileNetV2.py:110).  w is the reference tensor (C,1,3,3) = [C][9].
// Thread mapping shared by all Pretrain streaming kernels: a block of 256 threads is QB channel quads x RG pixel groups
// (QB = all C/4 quads when they fit, else an even split over gridDim.y); a thread keeps ONE quad for its whole life, so
// its 36 weights (4 channels x 9 taps = 144 contiguous bytes of w) sit in registers, and walks pixels p = blockIdx.x*RG
// + rg, += gridDim.x*RG.  Consecutive threads read consecutive 16-byte quads of a pixel and then the next pixel: fully
// coalesced for pixel-dense buffers whatever C is (96 channels = 24 quads no longer idle a quarter of each warp).
struct QMap {
  int qb, rg, ny;   // quads per block, pixel groups per block, blocks along the channel axis
};
static inline QMap qmap(int C) {
  const int cqn = C / 4;
  QMap m;
  m.ny = (cqn + 255) / 256;
  m.qb = (cqn + m.ny - 1) / m.ny;
  m.rg = 256 / m.qb;
  return m;
}

__device__ __forceinline__ void load_w36(const float* __restrict__ w, int q, float (&k)[36]) {
  const float4* src = reinterpret_cast<const float4*>(w + (long long)q * 36);   // channels 4q..4q+3, 9 taps each
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float4 v = __ldg(src + i);
    k[4 * i] = v.x, k[4 * i + 1] = v.y, k[4 * i + 2] = v.z, k[4 * i + 3] = v.w;
  }
}
// weight of channel j (0..3 within the quad), tap t: k[j*9 + t]
#define DWK(j, t) k[(j) * 9 + (t)]

template <int STRIDE>
__global__ void __launch_bounds__(256) dw3x3_fwd_kernel(PV x, PV y, const float* __restrict__ w, int qb, int rg) {
  const int ql = threadIdx.x % qb, g = threadIdx.x / qb;
  const int q = blockIdx.y * qb + ql;
  if (g >= rg || q * 4 >= x.c) return;
  float k[36];
  load_w36(w, q, k);
  const long long npix = (long long)y.n * y.h * y.w;
  for (long long p = (long long)blockIdx.x * rg + g; p < npix; p += (long long)gridDim.x * rg) {
    const int ox = (int)(p % y.w);
    const long long r = p / y.w;
    const int oy = (int)(r % y.h);
    const int n = (int)(r / y.h);
    // all nine 16-byte loads are issued unconditionally from clamped coordinates and zero-selected afterwards, so they
    // are in flight together (a branch per tap serialises them)
    float4 v[9];
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int iy = oy * STRIDE - 1 + kr;
      const int cy = min(max(iy, 0), x.h - 1);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int ix = ox * STRIDE - 1 + ks;
        const int cx = min(max(ix, 0), x.w - 1);
        v[kr * 3 + ks] = ld4(x.p + poff(x, n, cy, cx) + q * 4);
        if (iy != cy || ix != cx) v[kr * 3 + ks] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc.x = fmaf(v[t].x, DWK(0, t), acc.x);
      acc.y = fmaf(v[t].y, DWK(1, t), acc.y);
      acc.z = fmaf(v[t].z, DWK(2, t), acc.z);
      acc.w = fmaf(v[t].w, DWK(3, t), acc.w);
    }
    st4(y.p + poff(y, n, oy, ox) + q * 4, acc);
  }
}

// dx[n,iy,ix,c] (+)= sum_{kr,ks : (iy+1-kr) % s == 0, (ix+1-ks) % s == 0} dy[n,(iy+1-kr)/s,(ix+1-ks)/s,c] * w[c,kr,ks]
template <int STRIDE>
__global__ void __launch_bounds__(256) dw3x3_dgrad_kernel(PV dy, PV dx, const float* __restrict__ w, int accumulate, int qb,
                                                         int rg) {
  const int ql = threadIdx.x % qb, g = threadIdx.x / qb;
  const int q = blockIdx.y * qb + ql;
  if (g >= rg || q * 4 >= dx.c) return;
  float k[36];
  load_w36(w, q, k);
  const long long npix = (long long)dx.n * dx.h * dx.w;
  for (long long p = (long long)blockIdx.x * rg + g; p < npix; p += (long long)gridDim.x * rg) {
    const int ix = (int)(p % dx.w);
    const long long r = p / dx.w;
    const int iy = (int)(r % dx.h);
    const int n = (int)(r / dx.h);
    float4 v[9];
#pragma unroll
    for (int kr = 0; kr < 3; ++kr) {
      const int ty = iy + 1 - kr;
      const bool oky = ty >= 0 && (ty % STRIDE) == 0 && ty / STRIDE < dy.h;
      const int oy = oky ? ty / STRIDE : 0;
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        const int tx = ix + 1 - ks;
        const bool okx = tx >= 0 && (tx % STRIDE) == 0 && tx / STRIDE < dy.w;
        const int ox = okx ? tx / STRIDE : 0;
        v[kr * 3 + ks] = ld4(dy.p + poff(dy, n, oy, ox) + q * 4);
        if (!(oky && okx)) v[kr * 3 + ks] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc.x = fmaf(v[t].x, DWK(0, t), acc.x);
      acc.y = fmaf(v[t].y, DWK(1, t), acc.y);
      acc.z = fmaf(v[t].z, DWK(2, t), acc.z);
      acc.w = fmaf(v[t].w, DWK(3, t), acc.w);
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
// 36 accumulators per thread; the RG pixel groups of a block are folded through shared memory before ONE global atomic
// per (channel, tap) and block.
template <int STRIDE>
__global__ void __launch_bounds__(256) dw3x3_wgrad_kernel(PV x, PV dy, float* __restrict__ dw, int qb, int rg) {
  extern __shared__ float sacc[];   // [qb][36]
  for (int i = threadIdx.x; i < qb * 36; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int ql = threadIdx.x % qb, g = threadIdx.x / qb;
  const int q = blockIdx.y * qb + ql;
  if (g < rg && q * 4 < x.c) {
    float a[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) a[i] = 0.f;
    const long long npix = (long long)dy.n * dy.h * dy.w;
    for (long long p = (long long)blockIdx.x * rg + g; p < npix; p += (long long)gridDim.x * rg) {
      const int ox = (int)(p % dy.w);
      const long long r = p / dy.w;
      const int oy = (int)(r % dy.h);
      const int n = (int)(r / dy.h);
      const float4 gv = ld4(dy.p + poff(dy, n, oy, ox) + q * 4);
      float4 v[9];
#pragma unroll
      for (int kr = 0; kr < 3; ++kr) {
        const int iy = oy * STRIDE - 1 + kr;
        const int cy = min(max(iy, 0), x.h - 1);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const int ix = ox * STRIDE - 1 + ks;
          const int cx = min(max(ix, 0), x.w - 1);
          v[kr * 3 + ks] = ld4(x.p + poff(x, n, cy, cx) + q * 4);
          if (iy != cy || ix != cx) v[kr * 3 + ks] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        a[t] = fmaf(gv.x, v[t].x, a[t]);
        a[9 + t] = fmaf(gv.y, v[t].y, a[9 + t]);
        a[18 + t] = fmaf(gv.z, v[t].z, a[18 + t]);
        a[27 + t] = fmaf(gv.w, v[t].w, a[27 + t]);
      }
    }
#pragma unroll
    for (int i = 0; i < 36; ++i) atomicAdd(&sacc[ql * 36 + i], a[i]);   // a[j*9 + t] -> dw[(4q + j)*9 + t]: same order
  }
  __syncthreads();
  const int qn = min(qb, x.c / 4 - blockIdx.y * qb);
  float* dst = dw + (long long)blockIdx.y * qb * 36;
  for (int i = threadIdx.x; i < qn * 36; i += 256) atomicAdd(dst + i, sacc[i]);
}


// ---- column-strip variants (the production kernels for stride-1 forward / input gradient and all weight gradients).
// A block is QB channel quads x RG adjacent output columns of ONE image and walks down a strip of output rows keeping the
// 3x3 input window in registers: per output pixel a thread issues 3 (stride 1) or 6 (stride 2) new 16-byte loads instead
// of 9, its horizontal neighbours' loads hit L1, and every input row is fetched from L2 once per block (the per-pixel
// kernels above re-fetch each input ~3.7x from L2; ncu: DRAM traffic was already algorithmic, L1/L2 -> SM was the bound).
struct Row3 {
  float4 a, b, c;   // columns x-1, x, x+1 (stride 1) or 2x-1, 2x, 2x+1 (stride 2)
};
template <int STRIDE>
__device__ __forceinline__ Row3 load_row3(const PV& x, int n, int iy, int ox, int q) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  Row3 r{z, z, z};
  if (iy < 0 || iy >= x.h) return r;
  const int xc = ox * STRIDE;
  const float* base = x.p + poff(x, n, iy, 0) + q * 4;
  const int xl = xc - 1, xr = xc + 1;
  r.b = ld4(base + (long long)xc * x.sw);
  r.a = ld4(base + (long long)max(xl, 0) * x.sw);
  r.c = ld4(base + (long long)min(xr, x.w - 1) * x.sw);
  if (xl < 0) r.a = z;
  if (xr >= x.w) r.c = z;
  return r;
}
__device__ __forceinline__ void fma_row(float4& acc, const Row3& r, const float (&k)[36], int t0) {
  acc.x = fmaf(r.a.x, k[t0], acc.x), acc.x = fmaf(r.b.x, k[t0 + 1], acc.x), acc.x = fmaf(r.c.x, k[t0 + 2], acc.x);
  acc.y = fmaf(r.a.y, k[9 + t0], acc.y), acc.y = fmaf(r.b.y, k[9 + t0 + 1], acc.y), acc.y = fmaf(r.c.y, k[9 + t0 + 2], acc.y);
  acc.z = fmaf(r.a.z, k[18 + t0], acc.z), acc.z = fmaf(r.b.z, k[18 + t0 + 1], acc.z), acc.z = fmaf(r.c.z, k[18 + t0 + 2], acc.z);
  acc.w = fmaf(r.a.w, k[27 + t0], acc.w), acc.w = fmaf(r.b.w, k[27 + t0 + 1], acc.w), acc.w = fmaf(r.c.w, k[27 + t0 + 2], acc.w);
}

struct StripGeom {
  int qb, rg, ncg, nst, rows;   // column groups per row, strips per image, output rows per strip
};

// y[n,oy,ox,c] (+)= sum_{kr,ks} x[n, oy*S-1+kr, ox*S-1+ks, c] * w[c, kr, ks]   (FLIP: w[c, 2-kr, 2-ks] - the stride-1
// input gradient is the same correlation with the flipped kernel)
template <int STRIDE, bool FLIP>
__global__ void __launch_bounds__(256) dw3x3_strip_kernel(PV x, PV y, const float* __restrict__ w, StripGeom G,
                                                         int accumulate) {
  const int ql = threadIdx.x % G.qb, g = threadIdx.x / G.qb;
  const int q = blockIdx.y * G.qb + ql;
  int b = blockIdx.x;
  const int cg = b % G.ncg;
  b /= G.ncg;
  const int st = b % G.nst, n = b / G.nst;
  const int ox = cg * G.rg + g;
  if (g >= G.rg || ox >= y.w || q * 4 >= x.c) return;
  float k[36];
  load_w36(w, q, k);
  if (FLIP) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float tmp = k[j * 9 + t];
        k[j * 9 + t] = k[j * 9 + 8 - t];
        k[j * 9 + 8 - t] = tmp;
      }
  }
  const int oy0 = st * G.rows, oy1 = min(oy0 + G.rows, y.h);
  // the input rows of the NEXT output row are requested before the current one is computed (per-thread memory-level
  // parallelism: this kernel runs at ~25 % occupancy, 96+ registers)
  Row3 r0 = load_row3<STRIDE>(x, n, oy0 * STRIDE - 1, ox, q), r1, r2, p1, p2;
  if (STRIDE == 1) {
    r1 = load_row3<STRIDE>(x, n, oy0, ox, q);
    r2 = load_row3<STRIDE>(x, n, oy0 + 1, ox, q);
  } else {
    r1 = load_row3<STRIDE>(x, n, oy0 * 2, ox, q);
    r2 = load_row3<STRIDE>(x, n, oy0 * 2 + 1, ox, q);
  }
  for (int oy = oy0; oy < oy1; ++oy) {
    const bool more = oy + 1 < oy1;
    if (STRIDE == 1) {
      if (more) p2 = load_row3<STRIDE>(x, n, oy + 2, ox, q);
    } else if (more) {
      p1 = load_row3<STRIDE>(x, n, oy * 2 + 2, ox, q);
      p2 = load_row3<STRIDE>(x, n, oy * 2 + 3, ox, q);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    fma_row(acc, r0, k, 0);
    fma_row(acc, r1, k, 3);
    fma_row(acc, r2, k, 6);
    float* d = y.p + poff(y, n, oy, ox) + q * 4;
    if (accumulate) {
      const float4 o = ld4(d);
      acc.x += o.x, acc.y += o.y, acc.z += o.z, acc.w += o.w;
    }
    st4(d, acc);
    if (STRIDE == 1) {
      r0 = r1;
      r1 = r2;
      r2 = p2;
    } else {
      r0 = r2;
      r1 = p1;
      r2 = p2;
    }
  }
}

__device__ __forceinline__ void wg_row(float (&a)[36], const float4& g, const Row3& r, int t0) {
  a[t0] = fmaf(g.x, r.a.x, a[t0]), a[t0 + 1] = fmaf(g.x, r.b.x, a[t0 + 1]), a[t0 + 2] = fmaf(g.x, r.c.x, a[t0 + 2]);
  a[9 + t0] = fmaf(g.y, r.a.y, a[9 + t0]), a[9 + t0 + 1] = fmaf(g.y, r.b.y, a[9 + t0 + 1]), a[9 + t0 + 2] = fmaf(g.y, r.c.y, a[9 + t0 + 2]);
  a[18 + t0] = fmaf(g.z, r.a.z, a[18 + t0]), a[18 + t0 + 1] = fmaf(g.z, r.b.z, a[18 + t0 + 1]), a[18 + t0 + 2] = fmaf(g.z, r.c.z, a[18 + t0 + 2]);
  a[27 + t0] = fmaf(g.w, r.a.w, a[27 + t0]), a[27 + t0 + 1] = fmaf(g.w, r.b.w, a[27 + t0 + 1]), a[27 + t0 + 2] = fmaf(g.w, r.c.w, a[27 + t0 + 2]);
}

// dw[c][kr*3+ks] += sum over the block's strip of dy[n,oy,ox,c] * x[n, oy*S-1+kr, ox*S-1+ks, c]
template <int STRIDE>
__global__ void __launch_bounds__(256) dw3x3_strip_wgrad_kernel(PV x, PV dy, float* __restrict__ dw, StripGeom G) {
  extern __shared__ float sacc[];   // [qb][36]
  for (int i = threadIdx.x; i < G.qb * 36; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int ql = threadIdx.x % G.qb, g = threadIdx.x / G.qb;
  const int q = blockIdx.y * G.qb + ql;
  int b = blockIdx.x;
  const int cg = b % G.ncg;
  b /= G.ncg;
  const int st = b % G.nst, n = b / G.nst;
  const int ox = cg * G.rg + g;
  if (g < G.rg && ox < dy.w && q * 4 < x.c) {
    float a[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) a[i] = 0.f;
    const int oy0 = st * G.rows, oy1 = min(oy0 + G.rows, dy.h);
    Row3 r0 = load_row3<STRIDE>(x, n, oy0 * STRIDE - 1, ox, q), r1, r2, p1, p2;
    if (STRIDE == 1) {
      r1 = load_row3<STRIDE>(x, n, oy0, ox, q);
      r2 = load_row3<STRIDE>(x, n, oy0 + 1, ox, q);
    } else {
      r1 = load_row3<STRIDE>(x, n, oy0 * 2, ox, q);
      r2 = load_row3<STRIDE>(x, n, oy0 * 2 + 1, ox, q);
    }
    float4 gv = ld4(dy.p + poff(dy, n, oy0, ox) + q * 4), gn = gv;
    for (int oy = oy0; oy < oy1; ++oy) {
      const bool more = oy + 1 < oy1;
      if (more) {
        gn = ld4(dy.p + poff(dy, n, oy + 1, ox) + q * 4);
        if (STRIDE == 1) {
          p2 = load_row3<STRIDE>(x, n, oy + 2, ox, q);
        } else {
          p1 = load_row3<STRIDE>(x, n, oy * 2 + 2, ox, q);
          p2 = load_row3<STRIDE>(x, n, oy * 2 + 3, ox, q);
        }
      }
      wg_row(a, gv, r0, 0);
      wg_row(a, gv, r1, 3);
      wg_row(a, gv, r2, 6);
      gv = gn;
      if (STRIDE == 1) {
        r0 = r1;
        r1 = r2;
        r2 = p2;
      } else {
        r0 = r2;
        r1 = p1;
        r2 = p2;
      }
    }
#pragma unroll
    for (int i = 0; i < 36; ++i) atomicAdd(&sacc[ql * 36 + i], a[i]);
  }
  __syncthreads();
  const int qn = min(G.qb, x.c / 4 - blockIdx.y * G.qb);
  float* dst = dw + (long long)blockIdx.y * G.qb * 36;
  for (int i = threadIdx.x; i < qn * 36; i += 256) atomicAdd(dst + i, sacc[i]);
}

// ------------------------------------------------------------------------------------------- BatchNorm (training mode)
// nn.BatchNorm2d (MobileNetV2.py:107,111,115,152,168) over a pixel-dense [M][ld] matrix.
//   stats    : sums[0][c] += sum_p x, sums[1][c] += sum_p x^2   (fp32 per-thread partials of <= ~32 elements, then fp64);
//              the last block to finish writes coef[0..3][c] = scale, shift, mean, invstd, updates the running statistics
//              (momentum, unbiased variance) and re-zeroes the scratch: no separate finalize / memset launches
//   apply    : y = x*scale + shift (+ residual) -> optional ReLU6 -> optional tf32 rounding
//   backward : dz = dy * [0 < y < 6] ; dsums = (sum dz, sum dz*xhat) ; dx = scale * (dz - mean(dz) - xhat * mean(dz*xhat))
__device__ __forceinline__ void bn_finalize_channel(int c, double sum, double sumsq, long long M, int C,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float* __restrict__ running_mean, float* __restrict__ running_var,
                                                    float momentum, float eps, int training, float* __restrict__ coef) {
  double mean, var;
  if (training) {
    mean = sum / (double)M;
    var = sumsq / (double)M - mean * mean;
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

// Shared tail of the two reduction kernels: fold the RG pixel groups of the block (fp64), one atomic per (channel,
// component) and block.  part[8] = the thread's 2 x 4 partial sums (component-major: s0.xyzw, s1.xyzw).
__device__ __forceinline__ void bn_block_reduce(const float (&part)[8], int ql, int g, int qb, int rg, bool active, int C,
                                                int q0, double* __restrict__ sums, double* red /* [rg][qb][8] */) {
  if (active) {
    double* d = red + ((long long)g * qb + ql) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = (double)part[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < qb * 8; i += 256) {
    const int lq = i >> 3, e = i & 7;
    const int c = (q0 + lq) * 4 + (e & 3);
    if (c >= C) continue;
    double a = 0.0;
    for (int k = 0; k < rg; ++k) a += red[((long long)k * qb + lq) * 8 + e];
    atomicAdd(sums + (e >> 2) * C + c, a);
  }
}

__global__ void __launch_bounds__(256, 6) bn_stats_kernel(const float* __restrict__ x, long long M, long long ld, int C,
                                                      double* __restrict__ sums, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* __restrict__ running_mean,
                                                      float* __restrict__ running_var, float momentum, float eps,
                                                      float* __restrict__ coef, int qb, int rg) {
  extern __shared__ double red[];
  __shared__ int is_last;
  const int ql = threadIdx.x % qb, g = threadIdx.x / qb;
  const int q = blockIdx.y * qb + ql;
  const bool active = g < rg && q * 4 < C;
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    // a block streams ONE contiguous chunk of rows (sequential DRAM pages), its RG row groups interleaved inside it
    const long long step = rg;
    const long long chunk = (M + gridDim.x - 1) / gridDim.x;
    const long long Mend = min(M, (long long)(blockIdx.x + 1) * chunk);
    const float* px = x + q * 4;
    long long r = (long long)blockIdx.x * chunk + g;
    for (; r + 3 * step < Mend; r += 4 * step) {      // four independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld4(px + (r + u * step) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        part[0] += v[u].x, part[1] += v[u].y, part[2] += v[u].z, part[3] += v[u].w;
        part[4] = fmaf(v[u].x, v[u].x, part[4]), part[5] = fmaf(v[u].y, v[u].y, part[5]);
        part[6] = fmaf(v[u].z, v[u].z, part[6]), part[7] = fmaf(v[u].w, v[u].w, part[7]);
      }
    }
    for (; r < Mend; r += step) {
      const float4 v = ld4(px + r * ld);
      part[0] += v.x, part[1] += v.y, part[2] += v.z, part[3] += v.w;
      part[4] = fmaf(v.x, v.x, part[4]), part[5] = fmaf(v.y, v.y, part[5]);
      part[6] = fmaf(v.z, v.z, part[6]), part[7] = fmaf(v.w, v.w, part[7]);
    }
  }
  bn_block_reduce(part, ql, g, qb, rg, active, C, blockIdx.y * qb, sums, red);
  // the last block to arrive folds the statistics into (scale, shift, mean, invstd), updates the running statistics and
  // leaves the scratch zeroed for the next launch (threadfence-reduction pattern; the ticket lives behind the sums)
  unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + 2 * C);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < C; ch += 256) {
    const double su = __ldcg(sums + ch), sq = __ldcg(sums + C + ch);
    bn_finalize_channel(ch, su, sq, M, C, gamma, beta, running_mean, running_var, momentum, eps, 1, coef);
    sums[ch] = 0.0;
    sums[C + ch] = 0.0;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// eval mode only (running statistics); in training mode the last block of bn_stats_kernel finalises
__global__ void bn_finalize_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float eps,
                                   float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_finalize_channel(c, 0.0, 0.0, 1, C, gamma, beta, running_mean, running_var, 0.f, eps, 0, coef);
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, long long M, long long ldx, int C,
                                                      const float* __restrict__ coef, const float* __restrict__ res,
                                                      long long ldr, float* __restrict__ y, long long ldy, int act,
                                                      float slope, int round) {
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
    if (act == 1) {
      o.x = fminf(fmaxf(o.x, 0.f), 6.f), o.y = fminf(fmaxf(o.y, 0.f), 6.f);
      o.z = fminf(fmaxf(o.z, 0.f), 6.f), o.w = fminf(fmaxf(o.w, 0.f), 6.f);
    } else if (act == 2) {
      o.x = o.x > 0.f ? o.x : o.x * slope, o.y = o.y > 0.f ? o.y : o.y * slope;
      o.z = o.z > 0.f ? o.z : o.z * slope, o.w = o.w > 0.f ? o.w : o.w * slope;
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
                                                           double* __restrict__ dsums, int qb, int rg) {
  extern __shared__ double red[];
  const int ql = threadIdx.x % qb, g = threadIdx.x / qb;
  const int q = blockIdx.y * qb + ql;
  const bool active = g < rg && q * 4 < C;
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float4 a = ld4(coef + q * 4), b = ld4(coef + C + q * 4), mu = ld4(coef + 2 * C + q * 4),
                 is = ld4(coef + 3 * C + q * 4);
    const long long step = rg;
    const long long chunk = (M + gridDim.x - 1) / gridDim.x;
    const long long Mend = min(M, (long long)(blockIdx.x + 1) * chunk);
    const float* px = x + q * 4;
    const float* pg = dy + q * 4;
    auto add = [&](const float4& v, const float4& gr) {
      const float gx = relu6_gate(v.x, a.x, b.x, gr.x, relu6), gy = relu6_gate(v.y, a.y, b.y, gr.y, relu6),
                  gz = relu6_gate(v.z, a.z, b.z, gr.z, relu6), gw = relu6_gate(v.w, a.w, b.w, gr.w, relu6);
      part[0] += gx, part[1] += gy, part[2] += gz, part[3] += gw;
      part[4] = fmaf(gx, (v.x - mu.x) * is.x, part[4]), part[5] = fmaf(gy, (v.y - mu.y) * is.y, part[5]);
      part[6] = fmaf(gz, (v.z - mu.z) * is.z, part[6]), part[7] = fmaf(gw, (v.w - mu.w) * is.w, part[7]);
    };
    long long r = (long long)blockIdx.x * chunk + g;
    for (; r + 3 * step < Mend; r += 4 * step) {      // eight independent 16-byte loads in flight per thread
      float4 v[4], gr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = ld4(px + (r + u * step) * ldx);
        gr[u] = ld4(pg + (r + u * step) * ldd);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) add(v[u], gr[u]);
    }
    for (; r < Mend; r += step) add(ld4(px + r * ldx), ld4(pg + r * ldd));
  }
  bn_block_reduce(part, ql, g, qb, rg, active, C, blockIdx.y * qb, dsums, red);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, long long ldd,
                                                          const float* __restrict__ x, long long ldx, long long M, int C,
                                                          const float* __restrict__ coef, const double* __restrict__ dsums,
                                                          int relu6, int training, int have_sums, float* __restrict__ dx,
                                                          long long ldo, int accumulate, int round,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm[];  // c1[C] = mean(dz), c2[C] = mean(dz*xhat)
  __shared__ int is_last;
  float* c1 = sm;
  float* c2 = sm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double a = have_sums ? __ldcg(dsums + c) : 0.0, b = have_sums ? __ldcg(dsums + C + c) : 0.0;
    c1[c] = training ? (float)(a / (double)M) : 0.f;     // eval mode: statistics are constants, dx = scale * dz
    c2[c] = training ? (float)(b / (double)M) : 0.f;
    if (have_sums && blockIdx.x == 0) {
      if (dbeta) dbeta[c] = (float)a;
      if (dgamma) dgamma[c] = (float)b;
    }
  }
  __syncthreads();
  if (have_sums) {   // every block has consumed dsums once it takes a ticket: the last one leaves the scratch zeroed
    unsigned int* ticket = reinterpret_cast<unsigned int*>(const_cast<double*>(dsums) + 2 * C);
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      double* z = const_cast<double*>(dsums);
      for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) z[c] = 0.0;
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }
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


// ------------------------------------------------------------------------------------------------- MultiTaskDecoder
// MobileNetV2.py:536-649 + _calculate_accuracy (Pretrain.py:17-64), one CTA per sample.  Per class: candidates = points
// whose softmax score exceeds the confidence threshold; greedy distance-NMS in descending score order keeps a point and
// drops every remaining one within nms_dist of it; the first top_k kept points are the detections.  Equivalent
// formulation without a sort: repeat top_k times { arg-max score among live candidates (lowest index on ties); record;
// kill candidates with distance <= nms_dist }.  Outputs are fixed-shape: count[B][K], score[B][K][top_k],
// point[B][K][top_k][2] (unused slots zero).  accuracy[B] (optional, needs truth[B][8]): mean over the four landmark
// classes of the distance-bucket weight of the class's top-1 point (no detection = 0).
__global__ void __launch_bounds__(256) ssd_decode_kernel(const float* __restrict__ loc, const float* __restrict__ cls,
                                                        long long loc_stride, long long cls_stride, int n, int K, int top_k,
                                                        float conf_thr, float nms_dist, int* __restrict__ count,
                                                        float* __restrict__ score, float* __restrict__ point,
                                                        const float* __restrict__ truth, float* __restrict__ accuracy) {
  extern __shared__ float smd[];
  float* px = smd;          // [n]
  float* py = px + n;       // [n]
  float* sc = py + n;       // [n] score of the current class, -1 = not a candidate / suppressed
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ float best_x, best_y, acc_sum;
  __shared__ int best_i;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) acc_sum = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    px[i] = loc[b * loc_stride + 2 * i];
    py[i] = loc[b * loc_stride + 2 * i + 1];
  }
  __syncthreads();
  for (int c = 0; c < K; ++c) {
    for (int i = tid; i < n; i += blockDim.x) {
      const float* z = cls + b * cls_stride + (long long)i * K;
      float mx = z[0];
      for (int k = 1; k < K; ++k) mx = fmaxf(mx, z[k]);
      float sum = 0.f;
      for (int k = 0; k < K; ++k) sum += expf(z[k] - mx);
      const float p = expf(z[c] - mx) / sum;
      sc[i] = p > conf_thr ? p : -1.f;
    }
    __syncthreads();
    int found = 0;
    for (int t = 0; t < top_k; ++t) {
      float bv = -1.f;
      int bi = 0x7fffffff;
      for (int i = tid; i < n; i += blockDim.x)
        if (sc[i] > bv) bv = sc[i], bi = i;          // ascending i: the lowest index wins ties within a thread
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
      }
      if ((tid & 31) == 0) red_v[tid >> 5] = bv, red_i[tid >> 5] = bi;
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < 8; ++w)
          if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) bv = red_v[w], bi = red_i[w];
        best_i = bv > 0.f ? bi : -1;
        if (bv > 0.f) {
          best_x = px[bi], best_y = py[bi];
          const long long o = ((long long)b * K + c) * top_k + t;
          score[o] = bv;
          point[2 * o] = best_x, point[2 * o + 1] = best_y;
        }
      }
      __syncthreads();
      if (best_i < 0) break;
      ++found;
      if (c < 4 && t == 0 && truth && tid == 0) {     // Pretrain.py:33-58: bucket weights of the top-1 distance
        const float dx = best_x - truth[b * 8 + 2 * c], dy = best_y - truth[b * 8 + 2 * c + 1];
        const float d = sqrtf(dx * dx + dy * dy);
        const float w = (d > 0.f && d <= 5.f) ? 1.0f : (d <= 10.f && d > 5.f) ? 0.9f : (d <= 18.f && d > 10.f) ? 0.65f
                        : (d <= 30.f && d > 18.f) ? 0.35f : (d <= 45.f && d > 30.f) ? 0.1f : 0.f;
        acc_sum += w;
      }
      for (int i = tid; i < n; i += blockDim.x) {
        const float dx = px[i] - best_x, dy = py[i] - best_y;
        if (i == best_i || !(sqrtf(dx * dx + dy * dy) > nms_dist)) sc[i] = -1.f;
      }
      __syncthreads();
    }
    if (tid == 0) count[b * K + c] = found;
    __syncthreads();
  }
  if (tid == 0 && accuracy) accuracy[b] = acc_sum * 0.25f;
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

// strips: enough blocks to fill the machine (~4 per SM) but at least 4 output rows each so the window reuse pays
static StripGeom strip_geom(int C, int n, int out_h, int out_w, unsigned* blocks_x, unsigned* blocks_y) {
  const QMap m = qmap(C);
  StripGeom G;
  G.qb = m.qb;
  G.rg = std::min(m.rg, out_w);
  G.ncg = (out_w + G.rg - 1) / G.rg;
  const long long base = (long long)G.ncg * n * m.ny;
  const int want = (int)std::max(1ll, (592 + base - 1) / base);
  const int max_strips = std::max(1, out_h / 4);
  const int strips = std::min(want, max_strips);
  G.rows = (out_h + strips - 1) / strips;
  G.nst = (out_h + G.rows - 1) / G.rows;
  *blocks_x = (unsigned)(G.ncg * G.nst * n);
  *blocks_y = (unsigned)m.ny;
  return G;
}

extern "C" {

int tpgan_dwconv3x3(tpgan_view x, tpgan_view y, const float* w, int32_t stride, void* stream) {
  if (!vec_view(x) || !vec_view(y) || !w || x.c != y.c || x.n != y.n || (stride != 1 && stride != 2) ||
      y.h != (x.h + 2 - 3) / stride + 1 || y.w != (x.w + 2 - 3) / stride + 1)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3: bad geometry (C %% 4 == 0, stride 1|2, pad 1)");
  static const bool per_pixel = getenv("TPGAN_DW_PER_PIXEL") != nullptr;   // the previous per-pixel kernels (A/B timing)
  if (per_pixel) {
    const QMap m = qmap(x.c);
    const long long npix = (long long)y.n * y.h * y.w;
    dim3 grid((unsigned)grid_cap((npix + m.rg * 2 - 1) / (m.rg * 2), 8), (unsigned)m.ny);
    if (stride == 1)
      dw3x3_fwd_kernel<1><<<grid, 256, 0, ST>>>(pv(x), pv(y), w, m.qb, m.rg);
    else
      dw3x3_fwd_kernel<2><<<grid, 256, 0, ST>>>(pv(x), pv(y), w, m.qb, m.rg);
  } else {
    unsigned bx, by;
    const StripGeom G = strip_geom(x.c, y.n, y.h, y.w, &bx, &by);
    if (stride == 1)
      dw3x3_strip_kernel<1, false><<<dim3(bx, by), 256, 0, ST>>>(pv(x), pv(y), w, G, 0);
    else
      dw3x3_strip_kernel<2, false><<<dim3(bx, by), 256, 0, ST>>>(pv(x), pv(y), w, G, 0);
  }
  TPG_CHECK_LAUNCH("dwconv3x3");
  return 0;
}

int tpgan_dwconv3x3_dgrad(tpgan_view dy, tpgan_view dx, const float* w, int32_t stride, int32_t accumulate, void* stream) {
  if (!vec_view(dx) || !vec_view(dy) || !w || dx.c != dy.c || dx.n != dy.n || (stride != 1 && stride != 2) ||
      dy.h != (dx.h + 2 - 3) / stride + 1 || dy.w != (dx.w + 2 - 3) / stride + 1)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3_dgrad: bad geometry");
  static const bool per_pixel = getenv("TPGAN_DW_PER_PIXEL") != nullptr;
  if (stride == 1 && !per_pixel) {   // stride-1 input gradient = the same correlation with the flipped kernel
    unsigned bx, by;
    const StripGeom G = strip_geom(dx.c, dx.n, dx.h, dx.w, &bx, &by);
    dw3x3_strip_kernel<1, true><<<dim3(bx, by), 256, 0, ST>>>(pv(dy), pv(dx), w, G, accumulate);
  } else {
    const QMap m = qmap(dx.c);
    const long long npix = (long long)dx.n * dx.h * dx.w;
    dim3 grid((unsigned)grid_cap((npix + m.rg * 2 - 1) / (m.rg * 2), 8), (unsigned)m.ny);
    if (stride == 1)
      dw3x3_dgrad_kernel<1><<<grid, 256, 0, ST>>>(pv(dy), pv(dx), w, accumulate, m.qb, m.rg);
    else
      dw3x3_dgrad_kernel<2><<<grid, 256, 0, ST>>>(pv(dy), pv(dx), w, accumulate, m.qb, m.rg);
  }
  TPG_CHECK_LAUNCH("dwconv3x3_dgrad");
  return 0;
}

int tpgan_dwconv3x3_wgrad(tpgan_view x, tpgan_view dy, float* dw, int32_t stride, void* stream) {
  if (!vec_view(x) || !vec_view(dy) || !dw || x.c != dy.c || x.n != dy.n || (stride != 1 && stride != 2) ||
      dy.h != (x.h + 2 - 3) / stride + 1 || dy.w != (x.w + 2 - 3) / stride + 1)
    return set_error(TPGAN_ERR_INVALID, "dwconv3x3_wgrad: bad geometry");
  static const bool per_pixel = getenv("TPGAN_DW_PER_PIXEL") != nullptr;
  const QMap m = qmap(x.c);
  const size_t smem = (size_t)m.qb * 36 * 4;
  if (per_pixel) {
    const long long npix = (long long)dy.n * dy.h * dy.w;
    dim3 grid((unsigned)grid_cap((npix + m.rg * 8 - 1) / (m.rg * 8), 4), (unsigned)m.ny);   // >= 8 pixels per thread
    if (stride == 1)
      dw3x3_wgrad_kernel<1><<<grid, 256, smem, ST>>>(pv(x), pv(dy), dw, m.qb, m.rg);
    else
      dw3x3_wgrad_kernel<2><<<grid, 256, smem, ST>>>(pv(x), pv(dy), dw, m.qb, m.rg);
  } else {
    unsigned bx, by;
    const StripGeom G = strip_geom(x.c, dy.n, dy.h, dy.w, &bx, &by);
    if (stride == 1)
      dw3x3_strip_wgrad_kernel<1><<<dim3(bx, by), 256, smem, ST>>>(pv(x), pv(dy), dw, G);
    else
      dw3x3_strip_wgrad_kernel<2><<<dim3(bx, by), 256, smem, ST>>>(pv(x), pv(dy), dw, G);
  }
  TPG_CHECK_LAUNCH("dwconv3x3_wgrad");
  return 0;
}

int tpgan_bn_forward(tpgan_view x, tpgan_view res, tpgan_view y, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int32_t training, int32_t act, float slope,
                     int32_t round_tf32, double* sums, float* coef, void* stream) {
  const int relu6 = act == 1;
  if (act < 0 || act > 2) return set_error(TPGAN_ERR_INVALID, "bn_forward: act must be 0 (none), 1 (ReLU6) or 2 (leaky)");
  if (!dense(x) || !dense(y) || x.c != y.c || x.n != y.n || x.h != y.h || x.w != y.w || !gamma || !beta || !coef ||
      (training && !sums) || (!training && (!running_mean || !running_var)) || x.c * 8 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "bn_forward: bad arguments (pixel-dense views, C %% 4 == 0)");
  if (res.ptr && (!dense(res) || res.c != x.c || res.n != x.n || res.h != x.h || res.w != x.w || relu6))
    return set_error(TPGAN_ERR_INVALID, "bn_forward: residual must match x and excludes ReLU6");
  const long long M = (long long)x.n * x.h * x.w;
  const int C = x.c;
  if (training) {
    const QMap m = qmap(C);
    static const int per_sm = env_int("TPGAN_BN_PER_SM", 6);
    dim3 grid((unsigned)grid_cap((M + m.rg * 8 - 1) / (m.rg * 8), per_sm), (unsigned)m.ny);   // >= 8 rows per thread
    bn_stats_kernel<<<grid, 256, (size_t)m.rg * m.qb * 64, ST>>>(x.ptr, M, x.sw, C, sums, gamma, beta, running_mean,
                                                                  running_var, momentum, eps, coef, m.qb, m.rg);
    TPG_CHECK_LAUNCH("bn_stats");
  } else {
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST>>>(C, gamma, beta, running_mean, running_var, eps, coef);
    TPG_CHECK_LAUNCH("bn_finalize");
  }
  const long long total = M * (C / 4);
  bn_apply_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)C * 8, ST>>>(x.ptr, M, x.sw, C, coef, res.ptr, res.sw,
                                                                               y.ptr, y.sw, act, slope, round_tf32);
  TPG_CHECK_LAUNCH("bn_apply");
  return 0;
}

int tpgan_bn_backward(tpgan_view dy, tpgan_view x, tpgan_view dx, const float* coef, int32_t training, int32_t relu6,
                      int32_t accumulate, int32_t round_tf32, double* dsums, float* dgamma, float* dbeta, void* stream) {
  if (!dense(x) || !dense(dy) || !dense(dx) || x.c != dy.c || x.c != dx.c || x.n != dy.n || x.h != dy.h || x.w != dy.w ||
      x.n != dx.n || x.h != dx.h || x.w != dx.w || !coef || ((training || dgamma || dbeta) && !dsums) || x.c * 8 > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "bn_backward: bad arguments");
  const long long M = (long long)x.n * x.h * x.w;
  const int C = x.c;
  // the reduction feeds dx in training mode and the affine gradients in both modes (eval: BatchNorm frozen statistics with
  // trainable gamma / beta, as nn.BatchNorm2d.eval() behaves under autograd)
  const int have_sums = (training || dgamma || dbeta) ? 1 : 0;
  if (have_sums) {
    const QMap m = qmap(C);
    static const int per_sm = env_int("TPGAN_BN_PER_SM", 5);
    dim3 grid((unsigned)grid_cap((M + m.rg * 8 - 1) / (m.rg * 8), per_sm), (unsigned)m.ny);
    bn_bwd_reduce_kernel<<<grid, 256, (size_t)m.rg * m.qb * 64, ST>>>(dy.ptr, dy.sw, x.ptr, x.sw, M, C, coef, relu6, dsums,
                                                                       m.qb, m.rg);
    TPG_CHECK_LAUNCH("bn_bwd_reduce");
  }
  const long long total = M * (C / 4);
  bn_bwd_apply_kernel<<<grid_cap((total + 255) / 256, 8), 256, (size_t)C * 8, ST>>>(
      dy.ptr, dy.sw, x.ptr, x.sw, M, C, coef, dsums, relu6, training, have_sums, dx.ptr, dx.sw, accumulate, round_tf32, dgamma,
      dbeta);
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

int tpgan_ssd_decode(const float* loc, const float* cls, int32_t batch, int32_t n, int64_t loc_stride, int64_t cls_stride,
                     int32_t num_classes, int32_t top_k, float confidence_threshold, float nms_distance, int32_t* count,
                     float* score, float* point, const float* truth, float* accuracy, void* stream) {
  if (!loc || !cls || !count || !score || !point || batch < 1 || n < 1 || n > 4000 || num_classes < 1 || num_classes > 8 ||
      top_k < 1 || top_k > 64 || loc_stride < 2ll * n || cls_stride < (long long)num_classes * n || (accuracy && !truth))
    return set_error(TPGAN_ERR_INVALID, "ssd_decode: bad arguments (n <= 4000, classes <= 8, top_k <= 64)");
  const size_t bytes = (size_t)batch * num_classes * top_k * sizeof(float);
  cudaError_t e = cudaMemsetAsync(score, 0, bytes, ST);
  if (e == cudaSuccess) e = cudaMemsetAsync(point, 0, 2 * bytes, ST);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  ssd_decode_kernel<<<batch, 256, (size_t)n * 12, ST>>>(loc, cls, loc_stride, cls_stride, n, num_classes, top_k,
                                                       confidence_threshold, nms_distance, count, score, point, truth,
                                                       accuracy);
  TPG_CHECK_LAUNCH("ssd_decode");
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
