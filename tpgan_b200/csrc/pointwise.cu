// HBM-bound kernels of the TP-GAN hot path: weight (un)packing, layout conversion, activation backward, bias
// gradient, reflection padding, landmark patch crop, LocalFuser max-stitch, fused losses, maxout, Adam.
// All tensors are fp32 NHWC views (unit channel stride).  Grid-stride loops, grids sized from the SM count.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdint.h>

#include "../../include/tpgan_b200.h"
#include "common.cuh"
#include "host_common.h"

namespace tpg {

static inline int grid_for(long long n, int block, int per_sm = 8) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)std::max(1, device_sm_count()) * per_sm;
  if (device_sm_count() == 0) cap = 148 * per_sm;
  return (int)std::max(1ll, std::min(want, cap));
}

struct V {  // device copy of tpgan_view
  float* p;
  long long sn, sh, sw;
  int n, h, w, c;
};
static inline V dv(const tpgan_view& v) { return V{v.ptr, v.sn, v.sh, v.sw, v.n, v.h, v.w, v.c}; }
__device__ __forceinline__ long long voff(const V& v, int n, int y, int x) {
  return (long long)n * v.sn + (long long)y * v.sh + (long long)x * v.sw;
}

// ------------------------------------------------------------------------------------------------ pack / unpack
__global__ void pack_kernel(const float* __restrict__ ref, float* __restrict__ packed, int taps, int rows, int k,
                            int rows_pad, int k_pad, long long rs, long long ks, const int* __restrict__ row_map,
                            const int* __restrict__ k_map, int round) {
  const long long total = (long long)(taps + 1) * rows_pad * k_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(i % k_pad);
    long long r2 = i / k_pad;
    int rr = (int)(r2 % rows_pad);
    int t = (int)(r2 / rows_pad);
    float v = 0.f;
    if (t < taps && rr < rows && kk < k) {
      int rref = row_map ? row_map[rr] : rr;
      int kref = k_map ? k_map[kk] : kk;
      if (rref >= 0 && kref >= 0) v = ref[rref * rs + kref * ks + t];
    }
    packed[i] = (round == 1) ? round_tf32(v) : ((round == 2) ? (v - round_tf32(v)) : v);
  }
}

__global__ void unpack_kernel(const float* __restrict__ packed, float* __restrict__ ref, int taps, int rows, int k,
                              int rows_pad, int k_pad, long long rs, long long ks, const int* __restrict__ row_map,
                              const int* __restrict__ k_map, int accumulate) {
  const long long total = (long long)taps * rows * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(i % k);
    long long r2 = i / k;
    int rr = (int)(r2 % rows);
    int t = (int)(r2 / rows);
    int rref = row_map ? row_map[rr] : rr;
    int kref = k_map ? k_map[kk] : kk;
    if (rref < 0 || kref < 0) continue;
    float v = packed[((long long)t * rows_pad + rr) * k_pad + kk];
    float* d = ref + rref * rs + kref * ks + t;
    *d = accumulate ? (*d + v) : v;
  }
}


// ---- fast (un)packing for the row-contiguous case (ref_k_stride == taps): one block per packed row, the reference row
// (k_ref x taps contiguous floats) staged through shared memory so that both the global read and the global write are
// coalesced.  Used after every optimizer step (pack) and after every backward (unpack) for all 172 conv layers.
// Row staging in shared memory is indexed [k][tap]; the pack / unpack phases read it with consecutive k per thread, i.e.
// with a stride of `taps` floats - a 32-way bank conflict for taps = 64 (fc1's 8x8 window), 4-way for taps = 4.  An odd
// stride (taps | 1) makes it conflict-free; the linear phase maps i -> (i / taps) * stride + i % taps.
__device__ __forceinline__ int srow_index(int i, int taps, int tstride) {
  if (tstride == taps) return i;
  const int q = i / taps;
  return q * tstride + (i - q * taps);
}

__global__ void __launch_bounds__(1024) pack_rows_kernel(const float* __restrict__ ref, float* __restrict__ packed, int taps,
                                                         int rows, int k, int rows_pad, int k_pad, long long rs, int row_len,
                                                         const int* __restrict__ row_map, const int* __restrict__ k_map,
                                                         int round) {
  extern __shared__ float srow[];
  const int r = blockIdx.x;
  const int tstride = taps | 1;
  const int rref = (r < rows) ? (row_map ? row_map[r] : r) : -1;
  if (rref >= 0) {
    const float* src = ref + (long long)rref * rs;
    if ((row_len & 3) == 0 && (rs & 3) == 0 && (((uintptr_t)ref) & 15) == 0) {   // 16-byte loads of the reference row
      for (int i = threadIdx.x * 4; i < row_len; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        srow[srow_index(i, taps, tstride)] = v.x;
        srow[srow_index(i + 1, taps, tstride)] = v.y;
        srow[srow_index(i + 2, taps, tstride)] = v.z;
        srow[srow_index(i + 3, taps, tstride)] = v.w;
      }
    } else {
      for (int i = threadIdx.x; i < row_len; i += blockDim.x) srow[srow_index(i, taps, tstride)] = src[i];
    }
  }
  __syncthreads();
  auto cvt = [&](float v) { return (round == 1) ? round_tf32(v) : ((round == 2) ? (v - round_tf32(v)) : v); };
  // (tap, 4 consecutive k) per thread iteration: one 16-byte store
  const int kq = k_pad >> 2;
  for (int idx = threadIdx.x; idx < (taps + 1) * kq; idx += blockDim.x) {
    const int t = idx / kq, kk = (idx - t * kq) * 4;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < taps && rref >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kk + j < k) {
          const int kref = k_map ? k_map[kk + j] : kk + j;
          if (kref >= 0) o[j] = srow[kref * tstride + t];
        }
      }
    }
    *reinterpret_cast<float4*>(packed + ((long long)t * rows_pad + r) * k_pad + kk) = make_float4(cvt(o[0]), cvt(o[1]), cvt(o[2]), cvt(o[3]));
  }
}
__global__ void __launch_bounds__(1024) unpack_rows_kernel(const float* __restrict__ packed, float* __restrict__ ref, int taps,
                                                           int rows, int k, int rows_pad, int k_pad, long long rs, int row_len,
                                                           const int* __restrict__ row_map, const int* __restrict__ k_map,
                                                           int accumulate) {
  extern __shared__ float srow[];
  const int r = blockIdx.x;
  const int tstride = taps | 1;
  const int rref = row_map ? row_map[r] : r;
  if (rref < 0) return;
  const int kq = k_pad >> 2;
  for (int idx = threadIdx.x; idx < taps * kq; idx += blockDim.x) {
    const int t = idx / kq, kk = (idx - t * kq) * 4;
    const float4 v = *reinterpret_cast<const float4*>(packed + ((long long)t * rows_pad + r) * k_pad + kk);
    const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (kk + j < k) {
        const int kref = k_map ? k_map[kk + j] : kk + j;
        if (kref >= 0) srow[kref * tstride + t] = o[j];
      }
    }
  }
  __syncthreads();
  float* dst = ref + (long long)rref * rs;
  for (int i = threadIdx.x; i < row_len; i += blockDim.x) {
    const float v = srow[srow_index(i, taps, tstride)];
    dst[i] = accumulate ? (dst[i] + v) : v;
  }
}
// dst[t][kk][r] = src[t][r][kk] for r < rows, kk < k (per-tap transpose between the forward and the dgrad packing)
__global__ void transpose_packed_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int k,
                                        int rows_src_pad, int k_src_pad, int rows_dst_pad, int k_dst_pad) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const float* s = src + (long long)t * rows_src_pad * k_src_pad;
  float* d = dst + (long long)t * rows_dst_pad * k_dst_pad;
  const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, kk = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && kk < k) ? s[(long long)r * k_src_pad + kk] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int kk = k0 + i, r = r0 + threadIdx.x;
    if (kk < k && r < rows) d[(long long)kk * k_dst_pad + r] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, V dst, int round) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dst.c);
    long long r = i / dst.c;
    int x = (int)(r % dst.w);
    r /= dst.w;
    int y = (int)(r % dst.h);
    int n = (int)(r / dst.h);
    float v = src[(((long long)n * dst.c + c) * dst.h + y) * dst.w + x];
    dst.p[voff(dst, n, y, x) + c] = round ? round_tf32(v) : v;
  }
}
__global__ void nhwc_to_nchw_kernel(V src, float* __restrict__ dst) {
  const long long total = (long long)src.n * src.h * src.w * src.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % src.w);
    long long r = i / src.w;
    int y = (int)(r % src.h);
    r /= src.h;
    int c = (int)(r % src.c);
    int n = (int)(r / src.c);
    dst[i] = src.p[voff(src, n, y, x) + c];
  }
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void act_backward_kernel(V src, V mask, V dst, const float* __restrict__ slopes, float slope) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dst.c);
    long long r = i / dst.c;
    int x = (int)(r % dst.w);
    r /= dst.w;
    int y = (int)(r % dst.h);
    int n = (int)(r / dst.h);
    float g = src.p[voff(src, n, y, x) + c];
    float m = mask.p[voff(mask, n, y, x) + c];
    float s = slopes ? slopes[c] : slope;
    dst.p[voff(dst, n, y, x) + c] = m > 0.f ? g : g * s;
  }
}
__global__ void view_copy_kernel(V src, V dst, int accumulate) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dst.c);
    long long r = i / dst.c;
    int x = (int)(r % dst.w);
    r /= dst.w;
    int y = (int)(r % dst.h);
    int n = (int)(r / dst.h);
    float v = src.p[voff(src, n, y, x) + c];
    float* d = dst.p + voff(dst, n, y, x) + c;
    *d = accumulate ? (*d + v) : v;
  }
}

// db[c] (+)= sum_pixels dy[pix, c].  Pixel-dense views (sh == w*sw, sn == h*sh: every slice of an NHWC buffer) take the
// vectorised path: a warp reads 32 x float4 = 128 consecutive channels of a pixel, 4 pixels in flight per thread.
__global__ void bias_grad_kernel(V dy, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  const long long npix = (long long)dy.n * dy.h * dy.w;
  float acc = 0.f;
  if (c < dy.c) {
    for (long long pix = (long long)blockIdx.x * 8 + pl; pix < npix; pix += (long long)gridDim.x * 8) {
      int x = (int)(pix % dy.w);
      long long r = pix / dy.w;
      int y = (int)(r % dy.h);
      int n = (int)(r / dy.h);
      acc += dy.p[voff(dy, n, y, x) + c];
    }
  }
  red[pl][cl] = acc;
  __syncthreads();
  if (pl == 0 && c < dy.c) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][cl];
    atomicAdd(db + c, s);
  }
}
__global__ void __launch_bounds__(256) bias_grad_dense_kernel(const float* __restrict__ p, long long npix, long long sw, int C,
                                                              float* __restrict__ db) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + lane) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    const float* base = p + c;
    const long long stride = (long long)gridDim.x * 8;
    long long pix = (long long)blockIdx.x * 8 + wp;
    for (; pix + 3 * stride < npix; pix += 4 * stride) {
      const float4 v0 = *reinterpret_cast<const float4*>(base + pix * sw);
      const float4 v1 = *reinterpret_cast<const float4*>(base + (pix + stride) * sw);
      const float4 v2 = *reinterpret_cast<const float4*>(base + (pix + 2 * stride) * sw);
      const float4 v3 = *reinterpret_cast<const float4*>(base + (pix + 3 * stride) * sw);
      a.x += (v0.x + v1.x) + (v2.x + v3.x);
      a.y += (v0.y + v1.y) + (v2.y + v3.y);
      a.z += (v0.z + v1.z) + (v2.z + v3.z);
      a.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; pix < npix; pix += stride) {
      const float4 v = *reinterpret_cast<const float4*>(base + pix * sw);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  red[wp][lane] = a;
  __syncthreads();
  if (wp == 0 && c < C) {
    float4 s = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += red[i][lane].x; s.y += red[i][lane].y; s.z += red[i][lane].z; s.w += red[i][lane].w; }
    atomicAdd(db + c, s.x);
    if (c + 1 < C) atomicAdd(db + c + 1, s.y);
    if (c + 2 < C) atomicAdd(db + c + 2, s.z);
    if (c + 3 < C) atomicAdd(db + c + 3, s.w);
  }
}

// ------------------------------------------------------------------------------------------------ reflection pad
__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__global__ void reflect_pad_kernel(V src, V dst, int left, int top) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dst.c);
    long long r = i / dst.c;
    int x = (int)(r % dst.w);
    r /= dst.w;
    int y = (int)(r % dst.h);
    int n = (int)(r / dst.h);
    int sy = reflect(y - top, src.h), sx = reflect(x - left, src.w);
    dst.p[voff(dst, n, y, x) + c] = src.p[voff(src, n, sy, sx) + c];
  }
}
// gather form of the backward: each source pixel sums the padded pixels that mirror onto it
__global__ void reflect_pad_backward_kernel(V dpad, V dsrc, int left, int top, int accumulate) {
  const long long total = (long long)dsrc.n * dsrc.h * dsrc.w * dsrc.c;
  const int right = dpad.w - dsrc.w - left, bottom = dpad.h - dsrc.h - top;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dsrc.c);
    long long r = i / dsrc.c;
    int x = (int)(r % dsrc.w);
    r /= dsrc.w;
    int y = (int)(r % dsrc.h);
    int n = (int)(r / dsrc.h);
    // candidate padded coordinates mapping to (y, x): direct, top/left mirror, bottom/right mirror
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + top;
    if (y >= 1 && y <= top) ys[ny++] = top - y;
    if (y <= dsrc.h - 2 && y >= dsrc.h - 1 - bottom) ys[ny++] = top + 2 * (dsrc.h - 1) - y;
    xs[nx++] = x + left;
    if (x >= 1 && x <= left) xs[nx++] = left - x;
    if (x <= dsrc.w - 2 && x >= dsrc.w - 1 - right) xs[nx++] = left + 2 * (dsrc.w - 1) - x;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) s += dpad.p[voff(dpad, n, ys[a], xs[b]) + c];
    float* d = dsrc.p + voff(dsrc, n, y, x) + c;
    *d = accumulate ? (*d + s) : s;
  }
}

// ------------------------------------------------------------------------------------------------ patch crop
struct CropOut {
  V v[4];
};
// one warp per (image, part, patch row): lanes sweep (x, c) contiguously
// vec: every view stores a pixel as one aligned float4 (c <= 4, channel stride 4): a lane moves a whole pixel per access
__global__ void patch_crop_kernel(V img, const float* __restrict__ lm, CropOut out, int* __restrict__ boxes, float fill, int vec) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pw[4] = {40, 40, 40, 48}, phh[4] = {40, 40, 32, 32};
  int rows_total = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) rows_total += phh[i];  // 144 rows per image
  const long long total_rows = (long long)img.n * rows_total;
  for (long long wr = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wr < total_rows;
       wr += (long long)gridDim.x * warps_per_block) {
    int n = (int)(wr / rows_total);
    int rr = (int)(wr % rows_total);
    int part = 0;
    while (rr >= phh[part]) { rr -= phh[part]; ++part; }
    const float* l = lm + (long long)n * 10;
    float lx, ly;
    if (part < 3) {
      lx = l[part * 2];
      ly = l[part * 2 + 1];
    } else {  // mouth centre = mean of the two mouth corners (DataAndDataset.py:42-43), float32 arithmetic
      lx = __fdiv_rn(__fadd_rn(l[6], l[8]), 2.0f);
      ly = __fdiv_rn(__fadd_rn(l[7], l[9]), 2.0f);
    }
    const int cx = (int)floorf(lx), cy = (int)floorf(ly);
    const int left = cx - pw[part] / 2 + 1, upper = cy - phh[part] / 2 + 1;
    if (rr == 0 && lane < 4 && boxes) {
      int vals[4] = {left, upper, cx + pw[part] / 2 + 1, cy + phh[part] / 2 + 1};
      boxes[((long long)n * 4 + part) * 4 + lane] = vals[lane];
    }
    const V& o = out.v[part];
    const int sy = upper + rr;
    const int C = o.c;
    if (vec) {
      const bool row_in = sy >= 0 && sy < img.h;
      for (int x = lane; x < pw[part]; x += 32) {
        const int sx = left + x;
        float4 v = make_float4(fill, fill, fill, C > 3 ? fill : 0.f);
        if (row_in && sx >= 0 && sx < img.w) v = *reinterpret_cast<const float4*>(img.p + voff(img, n, sy, sx));
        *reinterpret_cast<float4*>(o.p + voff(o, n, rr, x)) = v;
      }
      continue;
    }
    for (int e = lane; e < pw[part] * C; e += 32) {
      int x = e / C, c = e % C;
      int sx = left + x;
      float v = fill;
      if (sy >= 0 && sy < img.h && sx >= 0 && sx < img.w) v = img.p[voff(img, n, sy, sx) + c];
      o.p[voff(o, n, rr, x) + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ LocalFuser
// Fixed paste rectangles, D_and_G_model.py:148-157: (left, top, width, height) of LE, RE, nose, mouth.
__constant__ int kFuseRect[4][4] = {{18, 19, 40, 40}, {65, 18, 40, 40}, {43, 47, 40, 32}, {40, 72, 48, 32}};

struct FuseIn {
  V v[4];
};
__global__ void local_fuse_kernel(FuseIn in, V out, uint8_t* __restrict__ argmax) {
  const long long total = (long long)out.n * out.h * out.w * out.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % out.c);
    long long r = i / out.c;
    int x = (int)(r % out.w);
    r /= out.w;
    int y = (int)(r % out.h);
    int n = (int)(r / out.h);
    // torch.max over the stack [LE, RE, N, M] of zero-padded maps; first maximal index wins
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int px = x - kFuseRect[k][0], py = y - kFuseRect[k][1];
      float v = 0.f;
      if (px >= 0 && px < kFuseRect[k][2] && py >= 0 && py < kFuseRect[k][3]) v = in.v[k].p[voff(in.v[k], n, py, px) + c];
      if (k == 0 || v > best) {
        best = v;
        arg = k;
      }
    }
    out.p[voff(out, n, y, x) + c] = best;
    if (argmax) argmax[i] = (uint8_t)arg;
  }
}
// Vectorised variants (C % 4 == 0, 16-byte aligned views - the 64-channel feature stitch): one thread = one pixel x four
// channels; float4 loads / stores, uchar4 arg-max, index math by shifts (the output is always 128 x 128).
__global__ void local_fuse_vec4_kernel(FuseIn in, V out, uint8_t* __restrict__ argmax) {
  const int c4n = out.c >> 2;
  const long long total = (long long)out.n * 16384 * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long pix = i / c4n;
    const int x = (int)(pix & 127), y = (int)((pix >> 7) & 127), n = (int)(pix >> 14);
    float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 arg = make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = x - kFuseRect[k][0], py = y - kFuseRect[k][1];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (px >= 0 && px < kFuseRect[k][2] && py >= 0 && py < kFuseRect[k][3])
        v = *reinterpret_cast<const float4*>(in.v[k].p + voff(in.v[k], n, py, px) + c);
      if (k == 0) {
        best = v;
      } else {   // strict > : the first maximal index of the stack [LE, RE, N, M] wins, as torch.max
        if (v.x > best.x) { best.x = v.x; arg.x = (unsigned char)k; }
        if (v.y > best.y) { best.y = v.y; arg.y = (unsigned char)k; }
        if (v.z > best.z) { best.z = v.z; arg.z = (unsigned char)k; }
        if (v.w > best.w) { best.w = v.w; arg.w = (unsigned char)k; }
      }
    }
    *reinterpret_cast<float4*>(out.p + voff(out, n, y, x) + c) = best;
    if (argmax) *reinterpret_cast<uchar4*>(argmax + pix * out.c + c) = arg;
  }
}
__global__ void local_fuse_backward_vec4_kernel(V dout, const uint8_t* __restrict__ argmax, FuseIn din, int accumulate) {
  const int c4n = dout.c >> 2;
  long long sizes[4], total = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sizes[k] = (long long)din.v[k].n * din.v[k].h * din.v[k].w * c4n;
    total += sizes[k];
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    int k = 0;
    while (j >= sizes[k]) { j -= sizes[k]; ++k; }
    const V& d = din.v[k];
    const int c = (int)(j % c4n) * 4;
    long long r = j / c4n;
    const int px = (int)(r % d.w);
    r /= d.w;
    const int py = (int)(r % d.h);
    const int n = (int)(r / d.h);
    const int x = px + kFuseRect[k][0], y = py + kFuseRect[k][1];
    const long long oi = ((((long long)n << 14) + (y << 7) + x)) * dout.c + c;
    const uchar4 a = *reinterpret_cast<const uchar4*>(argmax + oi);
    const float4 go = *reinterpret_cast<const float4*>(dout.p + voff(dout, n, y, x) + c);
    float4 g = make_float4(a.x == k ? go.x : 0.f, a.y == k ? go.y : 0.f, a.z == k ? go.z : 0.f, a.w == k ? go.w : 0.f);
    float4* dst = reinterpret_cast<float4*>(d.p + voff(d, n, py, px) + c);
    if (accumulate) {
      const float4 o = *dst;
      g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
    }
    *dst = g;
  }
}

__global__ void local_fuse_backward_kernel(V dout, const uint8_t* __restrict__ argmax, FuseIn din, int accumulate) {
  long long sizes[4], total = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sizes[k] = (long long)din.v[k].n * din.v[k].h * din.v[k].w * din.v[k].c;
    total += sizes[k];
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    int k = 0;
    while (j >= sizes[k]) { j -= sizes[k]; ++k; }
    const V& d = din.v[k];
    int c = (int)(j % d.c);
    long long r = j / d.c;
    int px = (int)(r % d.w);
    r /= d.w;
    int py = (int)(r % d.h);
    int n = (int)(r / d.h);
    int x = px + kFuseRect[k][0], y = py + kFuseRect[k][1];
    long long oi = (((long long)n * dout.h + y) * dout.w + x) * dout.c + c;
    float g = (argmax[oi] == k) ? dout.p[voff(dout, n, y, x) + c] : 0.f;
    float* dst = d.p + voff(d, n, py, px) + c;
    *dst = accumulate ? (*dst + g) : g;
  }
}

// ------------------------------------------------------------------------------------------------ losses
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ float block_mean(const V& f, int n, int y0, int x0, int s, int c) {
  float a = 0.f;
  for (int dy = 0; dy < s; ++dy)
    for (int dx = 0; dx < s; ++dx) a += f.p[voff(f, n, y0 + dy, x0 + dx) + c];
  return a / (float)(s * s);
}

// w: {l1_128, l1_64, l1_32, sym_128, sym_64, sym_32, tv_y, tv_x} coefficients (already divided by element counts)
struct LossW {
  float w[8];
};
__global__ void image_losses_kernel(V f, V g128, V g64, V g32, V df, LossW W, float* __restrict__ sums) {
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int H = f.h, Wd = f.w;
  const long long total = (long long)f.n * H * Wd * f.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % f.c);
    long long r = i / f.c;
    int x = (int)(r % Wd);
    r /= Wd;
    int y = (int)(r % H);
    int n = (int)(r / H);
    const int mx = Wd - 1 - x;
    float grad = 0.f;
    const float v = f.p[voff(f, n, y, x) + c];
    // pixel L1 @128
    float d = v - g128.p[voff(g128, n, y, x) + c];
    part[0] += fabsf(d);
    grad += W.w[0] * sgn(d);
    // symmetry @128
    float ds = v - f.p[voff(f, n, y, mx) + c];
    part[3] += fabsf(ds);
    grad += 2.f * W.w[3] * sgn(ds);
    // scales 64 and 32
#pragma unroll
    for (int lv = 1; lv <= 2; ++lv) {
      const int s = 1 << lv;
      const V& gt = (lv == 1) ? g64 : g32;
      const int by = y / s, bx = x / s;
      const int mbx = (Wd / s) - 1 - bx;
      const float m = block_mean(f, n, by * s, bx * s, s, c);
      const float mm = block_mean(f, n, by * s, mbx * s, s, c);
      const float dd = m - gt.p[voff(gt, n, by, bx) + c];
      const float dsm = m - mm;
      const bool owner = ((y % s) == 0) && ((x % s) == 0);
      if (owner) {
        part[lv] += fabsf(dd);
        part[3 + lv] += fabsf(dsm);
      }
      const float inv = 1.f / (float)(s * s);
      grad += W.w[lv] * inv * sgn(dd) + 2.f * W.w[3 + lv] * inv * sgn(dsm);
    }
    // total variation
    if (y + 1 < H) {
      float t = f.p[voff(f, n, y + 1, x) + c] - v;
      part[6] += fabsf(t);
      grad -= W.w[6] * sgn(t);
    }
    if (y > 0) grad += W.w[6] * sgn(v - f.p[voff(f, n, y - 1, x) + c]);
    if (x + 1 < Wd) {
      float t = f.p[voff(f, n, y, x + 1) + c] - v;
      part[7] += fabsf(t);
      grad -= W.w[7] * sgn(t);
    }
    if (x > 0) grad += W.w[7] * sgn(v - f.p[voff(f, n, y, x - 1) + c]);
    df.p[voff(df, n, y, x) + c] = grad;
  }
  // block reduction of the 8 partial sums
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = part[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[k][wp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, s);
  }
}

// Tiled variant for the layout the step uses (c <= 4 with a channel stride of 4 floats, W <= 256, W % 4 == 0): one block per
// (image, band of 4 rows).  The band and one halo row above / below are staged in shared memory with one coalesced 16-byte
// load per pixel, the 2x2 / 4x4 block means of the band (a band is aligned to both grids) are computed once, and every
// term of the pixel - mirror column, vertical / horizontal neighbours, block means and their mirrors - is read from shared
// memory: each global byte is read once (the halo rows twice) and the gradient leaves as one 16-byte store per pixel.
// The per-element arithmetic and its order are those of image_losses_kernel (identical gradients, sums to fp32 addition order).
__global__ void __launch_bounds__(128) image_losses_tiled_kernel(V f, V g128, V g64, V g32, V df, LossW W, float* __restrict__ sums) {
  extern __shared__ float4 tl_smem[];
  const int Wd = f.w, H = f.h;
  float4* tile = tl_smem;                 // [6][Wd]   rows y0-1 .. y0+4
  float4* m2 = tile + 6 * Wd;             // [2][Wd/2] 2x2 means of the band
  float4* m4 = m2 + Wd;                   // [Wd/4]    4x4 means of the band
  const int bands = H / 4;
  const int n = blockIdx.x / bands, y0 = (blockIdx.x % bands) * 4;
  const int tid = threadIdx.x;
  for (int i = tid; i < 6 * Wd; i += blockDim.x) {
    const int row = i / Wd, x = i - row * Wd, y = y0 - 1 + row;
    tile[i] = (y >= 0 && y < H) ? *reinterpret_cast<const float4*>(f.p + voff(f, n, y, x)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  auto ld = [&](const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); };
  for (int i = tid; i < Wd; i += blockDim.x) {          // 2 x (Wd/2) block means, (dy, dx) summation order of block_mean
    const int by = i / (Wd / 2), bx = i - by * (Wd / 2);
    float a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t = 0.f;
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) t += ld(tile[(1 + by * 2 + dy) * Wd + bx * 2 + dx], c);
      a[c] = t / 4.f;
    }
    m2[i] = make_float4(a[0], a[1], a[2], a[3]);
  }
  for (int i = tid; i < Wd / 4; i += blockDim.x) {
    float a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t = 0.f;
      for (int dy = 0; dy < 4; ++dy)
        for (int dx = 0; dx < 4; ++dx) t += ld(tile[(1 + dy) * Wd + i * 4 + dx], c);
      a[c] = t / 16.f;
    }
    m4[i] = make_float4(a[0], a[1], a[2], a[3]);
  }
  __syncthreads();
  float part[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = tid; i < 4 * Wd; i += blockDim.x) {
    const int ry = i / Wd, x = i - ry * Wd, y = y0 + ry, mx = Wd - 1 - x;
    const float4 v4 = tile[(1 + ry) * Wd + x], mir = tile[(1 + ry) * Wd + mx];
    const float4 up = tile[ry * Wd + x], dn = tile[(2 + ry) * Wd + x];
    const float4 lf = tile[(1 + ry) * Wd + (x > 0 ? x - 1 : x)], rt = tile[(1 + ry) * Wd + (x + 1 < Wd ? x + 1 : x)];
    const float4 t128 = *reinterpret_cast<const float4*>(g128.p + voff(g128, n, y, x));
    const float4 t64 = *reinterpret_cast<const float4*>(g64.p + voff(g64, n, y >> 1, x >> 1));
    const float4 t32 = *reinterpret_cast<const float4*>(g32.p + voff(g32, n, y >> 2, x >> 2));
    const float4 a2 = m2[(ry >> 1) * (Wd / 2) + (x >> 1)], a2m = m2[(ry >> 1) * (Wd / 2) + (Wd / 2 - 1 - (x >> 1))];
    const float4 a4 = m4[x >> 2], a4m = m4[Wd / 4 - 1 - (x >> 2)];
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= f.c) continue;
      const float v = ld(v4, c);
      float grad = 0.f;
      float d = v - ld(t128, c);
      part[0] += fabsf(d);
      grad += W.w[0] * sgn(d);
      float ds = v - ld(mir, c);
      part[3] += fabsf(ds);
      grad += 2.f * W.w[3] * sgn(ds);
      {
        const float dd = ld(a2, c) - ld(t64, c), dsm = ld(a2, c) - ld(a2m, c);
        if (((y & 1) == 0) && ((x & 1) == 0)) { part[1] += fabsf(dd); part[4] += fabsf(dsm); }
        const float inv = 1.f / 4.f;
        grad += W.w[1] * inv * sgn(dd) + 2.f * W.w[4] * inv * sgn(dsm);
      }
      {
        const float dd = ld(a4, c) - ld(t32, c), dsm = ld(a4, c) - ld(a4m, c);
        if (((y & 3) == 0) && ((x & 3) == 0)) { part[2] += fabsf(dd); part[5] += fabsf(dsm); }
        const float inv = 1.f / 16.f;
        grad += W.w[2] * inv * sgn(dd) + 2.f * W.w[5] * inv * sgn(dsm);
      }
      if (y + 1 < H) {
        const float t = ld(dn, c) - v;
        part[6] += fabsf(t);
        grad -= W.w[6] * sgn(t);
      }
      if (y > 0) grad += W.w[6] * sgn(v - ld(up, c));
      if (x + 1 < Wd) {
        const float t = ld(rt, c) - v;
        part[7] += fabsf(t);
        grad -= W.w[7] * sgn(t);
      }
      if (x > 0) grad += W.w[7] * sgn(v - ld(lf, c));
      g[c] = grad;
    }
    float* o = df.p + voff(df, n, y, x);
    if (df.c >= 4 || f.c >= 4) *reinterpret_cast<float4*>(o) = make_float4(g[0], g[1], g[2], g[3]);
    else {   // 3-channel views: the fourth lane is the layout's zero padding
      *reinterpret_cast<float4*>(o) = make_float4(g[0], g[1], g[2], 0.f);
    }
  }
  __shared__ float red[8][4];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float sacc = part[k];
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    if (lane == 0) red[k][wp] = sacc;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float sacc = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sacc += red[threadIdx.x][i];
    atomicAdd(sums + threadIdx.x, sacc);
  }
}

__global__ void l1_loss_kernel(V a, V b, V da, float coeff, float* __restrict__ sum) {
  float part = 0.f;
  const long long total = (long long)a.n * a.h * a.w * a.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % a.c);
    long long r = i / a.c;
    int x = (int)(r % a.w);
    r /= a.w;
    int y = (int)(r % a.h);
    int n = (int)(r / a.h);
    float d = a.p[voff(a, n, y, x) + c] - b.p[voff(b, n, y, x) + c];
    part += fabsf(d);
    if (da.p) da.p[voff(da, n, y, x) + c] = coeff * sgn(d);
  }
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[wp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(sum, s);
  }
}

// ------------------------------------------------------------------------------------------------ maxout / Adam / GP
__global__ void maxout2_kernel(const float* __restrict__ x, float* __restrict__ y, long long n_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    float a = x[2 * i], b = x[2 * i + 1];
    y[i] = a >= b ? a : b;  // MaxPool1d keeps the first element on ties
  }
}
__global__ void maxout2_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                        long long n_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    float a = x[2 * i], b = x[2 * i + 1];
    float g = dy[i];
    dx[2 * i] = a >= b ? g : 0.f;
    dx[2 * i + 1] = a >= b ? 0.f : g;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                            float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i];
    if (wd != 0.f) gi += wd * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// Graph-capturable Adam: the step count lives on the device (incremented by counter_inc_kernel before the launch).
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, float lr, float b1, float b2, float eps, float wd, const int* __restrict__ step,
                                float gscale) {
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float lr1 = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float pi = p[i];
    if (wd != 0.f) gi += wd * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - lr1 * (mi / denom);
  }
}
// float4 variant (all four arrays 16-byte aligned, n % 4 == 0: the flat parameter buffers): 112 B per thread iteration
__global__ void adam_dev_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                     float4* __restrict__ v, long long n4, float lr, float b1, float b2, float eps, float wd,
                                     const int* __restrict__ step, float gscale) {
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float lr1 = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = g[i];
    float4 p4 = p[i], m4 = m[i], v4 = v[i];
    float* pp = reinterpret_cast<float*>(&p4);
    float* mm = reinterpret_cast<float*>(&m4);
    float* vv = reinterpret_cast<float*>(&v4);
    const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gi = gg[j] * gscale;
      if (wd != 0.f) gi += wd * pp[j];
      const float mi = b1 * mm[j] + (1.f - b1) * gi;
      const float vi = b2 * vv[j] + (1.f - b2) * gi * gi;
      mm[j] = mi;
      vv[j] = vi;
      pp[j] = pp[j] - lr1 * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
    p[i] = p4;
    m[i] = m4;
    v[i] = v4;
  }
}
__global__ void counter_inc_kernel(int* c) { *c += 1; }

// one block per sample
__global__ void sample_sqnorm_kernel(V g, float* __restrict__ out) {
  const int n = blockIdx.x;
  const long long per = (long long)g.h * g.w * g.c;
  float part = 0.f;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    int c = (int)(i % g.c);
    long long r = i / g.c;
    int x = (int)(r % g.w);
    int y = (int)(r / g.w);
    float v = g.p[voff(g, n, y, x) + c];
    part += v * v;
  }
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[wp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    out[n] = s;
  }
}
__global__ void sample_scale_kernel(V g, const float* __restrict__ coeff, V u) {
  const long long total = (long long)g.n * g.h * g.w * g.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % g.c);
    long long r = i / g.c;
    int x = (int)(r % g.w);
    r /= g.w;
    int y = (int)(r % g.h);
    int n = (int)(r / g.h);
    u.p[voff(u, n, y, x) + c] = coeff[n] * g.p[voff(g, n, y, x) + c];
  }
}


// ------------------------------------------------------------------------------------------------ small elementwise ops
// out = a * b (dropout mask application and its backward)
__global__ void mul_kernel(V a, V b, V o) {
  const long long total = (long long)a.n * a.h * a.w * a.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % a.c);
    long long r = i / a.c;
    int x = (int)(r % a.w);
    r /= a.w;
    int y = (int)(r % a.h);
    int n = (int)(r / a.h);
    o.p[voff(o, n, y, x) + c] = a.p[voff(a, n, y, x) + c] * b.p[voff(b, n, y, x) + c];
  }
}
// out[n] = alpha[n] * a[n] + (1 - alpha[n]) * b[n]   (WGAN-GP interpolate)
__global__ void lerp_kernel(V a, V b, const float* __restrict__ alpha, V o) {
  const long long total = (long long)a.n * a.h * a.w * a.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % a.c);
    long long r = i / a.c;
    int x = (int)(r % a.w);
    r /= a.w;
    int y = (int)(r % a.h);
    int n = (int)(r / a.h);
    const float t = alpha[n];
    o.p[voff(o, n, y, x) + c] = t * a.p[voff(a, n, y, x) + c] + (1.f - t) * b.p[voff(b, n, y, x) + c];
  }
}
// gradient-penalty scalars: coeff[n] = scale * (||g_n|| - 1) / ||g_n|| ; gp_sum += sum_n (||g_n|| - 1)^2
__global__ void gp_coeff_kernel(const float* __restrict__ sqnorm, float* __restrict__ coeff, int n, float scale,
                                float* __restrict__ gp_sum) {
  float part = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float nr = sqrtf(sqnorm[i]);
    const float d = nr - 1.f;
    coeff[i] = nr > 0.f ? scale * d / nr : 0.f;
    part += d * d;
  }
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[wp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    if (gp_sum) atomicAdd(gp_sum, s);
  }
}
// fill every element of a view with a constant
__global__ void fill_kernel(V o, float v) {
  const long long total = (long long)o.n * o.h * o.w * o.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % o.c);
    long long r = i / o.c;
    int x = (int)(r % o.w);
    r /= o.w;
    int y = (int)(r % o.h);
    int n = (int)(r / o.h);
    o.p[voff(o, n, y, x) + c] = v;
  }
}
// softmax cross-entropy over rows of logits (B, C): loss_sum += sum_n -log softmax(x_n)[label_n];
// dlogits = coeff * (softmax - onehot).  One warp per row.
__global__ void softmax_ce_kernel(const float* __restrict__ x, long long row_stride, const long long* __restrict__ label,
                                  float* __restrict__ dx, long long drow_stride, int rows, int cols, float coeff,
                                  float* __restrict__ loss_sum, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * row_stride;
  float m = -INFINITY;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, xr[j]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s += expf(xr[j] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  // a label outside [0, cols) would be an out-of-bounds read: clamp it and raise the host-visible status word
  // (tpgan_kernel_status() != 0 -> the Python wrappers raise), code 0x7E
  const long long lab64 = label[row];
  const int lab = lab64 < 0 ? 0 : (lab64 >= cols ? cols - 1 : (int)lab64);
  if (lab64 != lab && lane == 0 && status) atomicCAS(status, 0, 0x7E | (row << 8));
  const float lse = m + logf(s);
  if (lane == 0 && loss_sum) atomicAdd(loss_sum, lse - xr[lab]);
  if (dx) {
    float* dr = dx + row * drow_stride;
    for (int j = lane; j < cols; j += 32) dr[j] = coeff * (expf(xr[j] - lse) - (j == lab ? 1.f : 0.f));
  }
}


// hi = rna_tf32(src), lo = src - hi  (operand split of the fp32-exact verification mode: a*w ~ ah*wh + ah*wl + al*wh)
__global__ void split_tf32_kernel(V s, V hi, V lo) {
  const long long total = (long long)s.n * s.h * s.w * s.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % s.c);
    long long r = i / s.c;
    int x = (int)(r % s.w);
    r /= s.w;
    int y = (int)(r % s.h);
    int n = (int)(r / s.h);
    const float v = s.p[voff(s, n, y, x) + c];
    const float h = round_tf32(v);
    hi.p[voff(hi, n, y, x) + c] = h;
    lo.p[voff(lo, n, y, x) + c] = v - h;
  }
}


// ------------------------------------------------------------------------------------------------ multi-tensor launches
// One launch over a device-resident job table instead of one tiny launch per layer (172 conv layers): bias gradients,
// weight packing (+ per-tap transposes) after the optimizer step, gradient unpacking after backward.
template <class Job>
__device__ __forceinline__ int find_job(const Job* jobs, int njobs, int block) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_begin <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------ bf16 operand copies
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// dst16 (bf16 view, same logical shape) = round-to-nearest-even(src).  One thread per (pixel, 4 channels): a 16-byte load
// and an 8-byte store when both views allow it, element-wise otherwise.  Padding lanes inside the last quad (channels
// >= c of a buffer whose channel stride is a multiple of 4) are copied too - they are zero in the source.
__global__ void cast_bf16_kernel(V s, uint16_t* __restrict__ d, long long dsn, long long dsh, long long dsw, int vec) {
  const int cq = (s.c + 3) >> 2;
  const long long total = (long long)s.n * s.h * s.w * cq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    long long r = i / cq;
    const int x = (int)(r % s.w);
    r /= s.w;
    const int y = (int)(r % s.h);
    const int n = (int)(r / s.h);
    const float* sp = s.p + voff(s, n, y, x) + 4 * q;
    uint16_t* dp = d + (long long)n * dsn + (long long)y * dsh + (long long)x * dsw + 4 * q;
    if (vec) {
      const float4 v = *reinterpret_cast<const float4*>(sp);
      *reinterpret_cast<uint2*>(dp) = make_uint2(cvt_bf16x2(v.x, v.y), cvt_bf16x2(v.z, v.w));
    } else {
      for (int j = 0; j < 4 && 4 * q + j < s.c; ++j) {
        uint16_t h;
        asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(sp[j]));
        dp[j] = h;
      }
    }
  }
}
// Packed fp32 weights [rows][k_pad] -> bf16 [rows][k_pad16] (k_pad16 >= k_pad, zero beyond), all layers in one launch.
// A block converts 2048 quads (4 consecutive k of one row) of one job.
__global__ void __launch_bounds__(256) cast_packed_multi_kernel(const tpgan_cast_job* __restrict__ jobs, int njobs) {
  const int ji = find_job(jobs, njobs, (int)blockIdx.x);
  const tpgan_cast_job J = jobs[ji];
  const int kq = J.k_pad16 >> 2;
  const long long total = J.rows * kq;
  long long i = (long long)((int)blockIdx.x - J.block_begin) * 2048 + threadIdx.x;
  for (int it = 0; it < 8 && i < total; ++it, i += 256) {
    const long long r = i / kq;
    const int kk = (int)(i - r * kq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk < J.k_pad) v = *reinterpret_cast<const float4*>(J.src + r * J.k_pad + kk);
    *reinterpret_cast<uint2*>(J.dst + r * J.k_pad16 + kk) = make_uint2(cvt_bf16x2(v.x, v.y), cvt_bf16x2(v.z, v.w));
  }
}

__global__ void __launch_bounds__(256) bias_grad_multi_kernel(const tpgan_bias_job* __restrict__ jobs, int njobs) {
  __shared__ float4 red[8][32];
  const int ji = find_job(jobs, njobs, (int)blockIdx.x);
  const tpgan_bias_job J = jobs[ji];
  const int local = (int)blockIdx.x - J.block_begin;
  const int cg = local % J.cgroups, pb = local / J.cgroups;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  // `lanes` lanes cover the (up to 4 * lanes) channels of one pixel; the 32 / lanes sub-groups of a warp take different
  // pixels, so narrow tensors (64, 3 channels) keep every lane loading 16 B
  const int lpp = J.lanes, ppw = 32 / lpp;
  const int sub = lane / lpp, cl = lane - sub * lpp;
  const int c = (cg * lpp + cl) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < J.c) {
    const float* base = J.dy + c;
    const long long stride = (long long)J.pix_blocks * 8 * ppw;
    long long pix = ((long long)pb * 8 + wp) * ppw + sub;
    for (; pix + 7 * stride < J.npix; pix += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(base + (pix + i * stride) * J.sw);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a.x += v[i].x; a.y += v[i].y; a.z += v[i].z; a.w += v[i].w; }
    }
    for (; pix < J.npix; pix += stride) {
      const float4 v = *reinterpret_cast<const float4*>(base + pix * J.sw);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  for (int o = lpp; o < 32; o <<= 1) {   // fold the pixel sub-groups of the warp
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
    a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
  }
  red[wp][lane] = a;
  __syncthreads();
  if (wp == 0 && lane < lpp && c < J.c) {
    float4 s = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += red[i][lane].x; s.y += red[i][lane].y; s.z += red[i][lane].z; s.w += red[i][lane].w; }
    atomicAdd(J.db + c, s.x);
    if (c + 1 < J.c) atomicAdd(J.db + c + 1, s.y);
    if (c + 2 < J.c) atomicAdd(J.db + c + 2, s.z);
    if (c + 3 < J.c) atomicAdd(J.db + c + 3, s.w);
  }
}

// mode 0: pack rows (ref -> packed), mode 1: unpack rows (packed -> ref, accumulate per job flag)
constexpr int kMultiRun = 8;        // consecutive tiles per block of transpose_multi_kernel (below)
// A block = one row.  These launches are bound by the bytes in flight per SM (resident rows x ~20 KB; shared memory caps the
// resident blocks at 8 for the longest rows): 256-thread blocks keep twice as many rows in flight as 512-thread ones
// (0.86 -> 0.65 ms per step); several rows per block, serialised, measured slower (1.01 ms).
constexpr int kPackThreads = 256;
__global__ void __launch_bounds__(kPackThreads) pack_multi_kernel(const tpgan_pack_job* __restrict__ jobs, int njobs, int mode) {
  extern __shared__ float srow[];
  const int ji = find_job(jobs, njobs, (int)blockIdx.x);
  const tpgan_pack_job J = jobs[ji];
  const int r = (int)blockIdx.x - J.block_begin;
  const int taps = J.taps;
  const int tstride = taps | 1;
  const int kq = J.k_pad >> 2;          // k_pad is a multiple of 32: (tap, 4 consecutive k) per thread = 16-byte accesses
  if (mode == 0) {
    const int rref = (r < J.rows) ? (J.row_map ? J.row_map[r] : r) : -1;
    if (rref >= 0) {
      const float* src = J.ref_c + (long long)rref * J.rs;
      if ((J.row_len & 3) == 0 && (J.rs & 3) == 0 && (((uintptr_t)J.ref_c) & 15) == 0) {
        for (int i = threadIdx.x * 4; i < J.row_len; i += blockDim.x * 4) {
          const float4 v = *reinterpret_cast<const float4*>(src + i);
          srow[srow_index(i, taps, tstride)] = v.x;
          srow[srow_index(i + 1, taps, tstride)] = v.y;
          srow[srow_index(i + 2, taps, tstride)] = v.z;
          srow[srow_index(i + 3, taps, tstride)] = v.w;
        }
      } else {
        for (int i = threadIdx.x; i < J.row_len; i += blockDim.x) srow[srow_index(i, taps, tstride)] = src[i];
      }
    }
    __syncthreads();
    auto cvt = [&](float v) { return (J.flag == 1) ? round_tf32(v) : ((J.flag == 2) ? (v - round_tf32(v)) : v); };
    for (int idx = threadIdx.x; idx < (taps + 1) * kq; idx += blockDim.x) {
      const int t = idx / kq, kk = (idx - t * kq) * 4;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (t < taps && rref >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (kk + j < J.k) {
            const int kref = J.k_map ? J.k_map[kk + j] : kk + j;
            if (kref >= 0) o[j] = srow[kref * tstride + t];
          }
        }
      }
      *reinterpret_cast<float4*>(J.packed + ((long long)t * J.rows_pad + r) * J.k_pad + kk) =
          make_float4(cvt(o[0]), cvt(o[1]), cvt(o[2]), cvt(o[3]));
    }
  } else {
    if (r >= J.rows) return;
    const int rref = J.row_map ? J.row_map[r] : r;
    if (rref < 0) return;
    for (int idx = threadIdx.x; idx < taps * kq; idx += blockDim.x) {
      const int t = idx / kq, kk = (idx - t * kq) * 4;
      const float4 v = *reinterpret_cast<const float4*>(J.packed + ((long long)t * J.rows_pad + r) * J.k_pad + kk);
      const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kk + j < J.k) {
          const int kref = J.k_map ? J.k_map[kk + j] : kk + j;
          if (kref >= 0) srow[kref * tstride + t] = o[j];
        }
      }
    }
    __syncthreads();
    float* dst = J.ref + (long long)rref * J.rs;
    if ((J.row_len & 3) == 0 && (J.rs & 3) == 0 && (((uintptr_t)J.ref) & 15) == 0 && tstride == taps) {
      for (int i = threadIdx.x * 4; i < J.row_len; i += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(srow + i);
        if (J.flag) {
          const float4 d = *reinterpret_cast<const float4*>(dst + i);
          v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
        }
        *reinterpret_cast<float4*>(dst + i) = v;
      }
    } else {
      for (int i = threadIdx.x; i < J.row_len; i += blockDim.x) {
        const float v = srow[srow_index(i, taps, tstride)];
        dst[i] = J.flag ? (dst[i] + v) : v;
      }
    }
  }
}

// A block takes kMultiRun consecutive 32x32 tiles and looks its job up once per run of tiles of the same job (the binary
// search over the job table is ~8 dependent global loads, which used to be paid per 4 KB tile): 0.567 -> 0.423 ms per step.
// (Issuing the loads of four tiles before one barrier measured slower: 0.74 ms.)
__global__ void __launch_bounds__(256) transpose_multi_kernel(const tpgan_transpose_job* __restrict__ jobs, int njobs, int total) {
  __shared__ float tile[2][32][33];
  const int vb0 = (int)blockIdx.x * kMultiRun, vb1 = min(total, vb0 + kMultiRun);
  int ji = -1, job_end = 0;
  tpgan_transpose_job J;
  for (int vb = vb0; vb < vb1; ++vb) {
    if (ji < 0 || vb >= job_end) {
      ji = find_job(jobs, njobs, vb);
      J = jobs[ji];
      job_end = (ji + 1 < njobs) ? jobs[ji + 1].block_begin : total;
    }
    float (*tile_)[33] = tile[vb & 1];   // alternate buffers: one barrier per tile
    int local = vb - J.block_begin;
    const int tk = local % J.tiles_k;
    local /= J.tiles_k;
    const int tr = local % J.tiles_r;
    const int t = local / J.tiles_r;
    const float* s = J.src + (long long)t * J.rows_src_pad * J.k_src_pad;
    float* d = J.dst + (long long)t * J.rows_dst_pad * J.k_dst_pad;
    const int k0 = tk * 32, r0 = tr * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = r0 + i, kk = k0 + threadIdx.x;
      tile_[i][threadIdx.x] = (r < J.rows && kk < J.k) ? s[(long long)r * J.k_src_pad + kk] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int kk = k0 + i, r = r0 + threadIdx.x;
      if (kk < J.k && r < J.rows) d[(long long)kk * J.k_dst_pad + r] = tile_[threadIdx.x][i];
    }
  }
}


// ------------------------------------------------------------------------------------------------ pooling (identity net)
// nn.MaxPool2d(3, 2, 1) (ResNet.py:33): out[oy][ox] = max over the 3x3 window at (2*oy-1, 2*ox-1); the window position of
// the first maximum (row-major scan, strict >, as ATen) is kept in `arg` for the backward pass.
__global__ void maxpool3s2_kernel(V x, V y, uint8_t* __restrict__ arg) {
  const long long total = (long long)y.n * y.h * y.w * y.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % y.c);
    long long r = i / y.c;
    int ox = (int)(r % y.w);
    r /= y.w;
    int oy = (int)(r % y.h);
    int n = (int)(r / y.h);
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int iy = 2 * oy - 1 + dy, ix = 2 * ox - 1 + dx;
        if (iy >= 0 && iy < x.h && ix >= 0 && ix < x.w) {
          const float v = x.p[voff(x, n, iy, ix) + c];
          if (bi == 0 || v > best) { best = v; bi = dy * 3 + dx + 1; }
        }
      }
    y.p[voff(y, n, oy, ox) + c] = best;
    if (arg) arg[i] = (uint8_t)(bi - 1);
  }
}
// gather form: input pixel (iy, ix) collects dy of the (at most 2x2) windows whose recorded maximum is this pixel
__global__ void maxpool3s2_backward_kernel(V dy, const uint8_t* __restrict__ arg, V dx, int accumulate) {
  const long long total = (long long)dx.n * dx.h * dx.w * dx.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dx.c);
    long long r = i / dx.c;
    int ix = (int)(r % dx.w);
    r /= dx.w;
    int iy = (int)(r % dx.h);
    int n = (int)(r / dx.h);
    float g = 0.f;
    // windows covering iy: oy with 2*oy-1 <= iy <= 2*oy+1  <=>  oy in [ceil((iy-1)/2), floor((iy+1)/2)]
    for (int oy = iy / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy < 0 || oy >= dy.h) continue;
      const int wy = iy - (2 * oy - 1);
      if (wy < 0 || wy > 2) continue;
      for (int ox = ix / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox < 0 || ox >= dy.w) continue;
        const int wx = ix - (2 * ox - 1);
        if (wx < 0 || wx > 2) continue;
        const long long oi = (((long long)n * dy.h + oy) * dy.w + ox) * dy.c + c;
        if (arg[oi] == wy * 3 + wx) g += dy.p[voff(dy, n, oy, ox) + c];
      }
    }
    float* d = dx.p + voff(dx, n, iy, ix) + c;
    *d = accumulate ? (*d + g) : g;
  }
}
// nn.AdaptiveAvgPool2d((1,1)) (ResNet.py:45): one warp per (n, 32-channel group)
__global__ void avgpool_kernel(V x, V y) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= x.c) return;
  float s = 0.f;
  for (int iy = 0; iy < x.h; ++iy)
    for (int ix = 0; ix < x.w; ++ix) s += x.p[voff(x, n, iy, ix) + c];
  y.p[voff(y, n, 0, 0) + c] = s / (float)(x.h * x.w);
}
__global__ void avgpool_backward_kernel(V dy, V dx, int accumulate) {
  const long long total = (long long)dx.n * dx.h * dx.w * dx.c;
  const float inv = 1.f / (float)(dx.h * dx.w);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % dx.c);
    long long r = i / dx.c;
    int ix = (int)(r % dx.w);
    r /= dx.w;
    int iy = (int)(r % dx.h);
    int n = (int)(r / dx.h);
    const float g = dy.p[voff(dy, n, 0, 0) + c] * inv;
    float* d = dx.p + voff(dx, n, iy, ix) + c;
    *d = accumulate ? (*d + g) : g;
  }
}

static bool view_vec4_ok(const tpgan_view& v) {
  return (((uintptr_t)v.ptr & 15) == 0) && v.sn % 4 == 0 && v.sh % 4 == 0 && v.sw % 4 == 0;
}
static bool same_geom(const tpgan_view& a, const tpgan_view& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}

}  // namespace tpg

using namespace tpg;
#define ST ((cudaStream_t)stream)

extern "C" {

int tpgan_pack_weights(const float* ref, float* packed, int32_t taps, int32_t rows, int32_t k, int32_t rows_pad,
                       int32_t k_pad, int64_t ref_row_stride, int64_t ref_k_stride, const int32_t* row_map,
                       const int32_t* k_map, int32_t round_tf32, void* stream) {
  if (!ref || !packed || taps < 1 || rows > rows_pad || k > k_pad) return set_error(TPGAN_ERR_INVALID, "pack: bad args");
  if (ref_k_stride == taps && ref_row_stride % taps == 0 && (ref_row_stride / taps) * (taps | 1) * 4 <= 192 * 1024) {
    const int row_len = (int)ref_row_stride;   // k_ref * taps contiguous floats per reference row
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(pack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
      cudaFuncSetAttribute(unpack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
      attr_set = true;
    }
    pack_rows_kernel<<<rows_pad, 1024, (size_t)((row_len / taps) * (taps | 1) + row_len % taps) * 4, ST>>>(ref, packed, taps, rows, k, rows_pad, k_pad, ref_row_stride,
                                                                row_len, row_map, k_map, round_tf32);
    TPG_CHECK_LAUNCH("pack_weights");
    return 0;
  }
  long long total = (long long)(taps + 1) * rows_pad * k_pad;
  pack_kernel<<<grid_for(total, 256), 256, 0, ST>>>(ref, packed, taps, rows, k, rows_pad, k_pad, ref_row_stride,
                                                    ref_k_stride, row_map, k_map, round_tf32);
  TPG_CHECK_LAUNCH("pack_weights");
  return 0;
}
int tpgan_unpack_weights(const float* packed, float* ref, int32_t taps, int32_t rows, int32_t k, int32_t rows_pad,
                         int32_t k_pad, int64_t ref_row_stride, int64_t ref_k_stride, const int32_t* row_map,
                         const int32_t* k_map, int32_t accumulate, void* stream) {
  if (!ref || !packed || taps < 1 || rows > rows_pad || k > k_pad) return set_error(TPGAN_ERR_INVALID, "unpack: bad args");
  if (ref_k_stride == taps && ref_row_stride % taps == 0 && (ref_row_stride / taps) * (taps | 1) * 4 <= 192 * 1024) {
    const int row_len = (int)ref_row_stride;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(pack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
      cudaFuncSetAttribute(unpack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
      attr_set = true;
    }
    unpack_rows_kernel<<<rows, 1024, (size_t)((row_len / taps) * (taps | 1) + row_len % taps) * 4, ST>>>(packed, ref, taps, rows, k, rows_pad, k_pad, ref_row_stride,
                                                              row_len, row_map, k_map, accumulate);
    TPG_CHECK_LAUNCH("unpack_weights");
    return 0;
  }
  long long total = (long long)taps * rows * k;
  unpack_kernel<<<grid_for(total, 256), 256, 0, ST>>>(packed, ref, taps, rows, k, rows_pad, k_pad, ref_row_stride,
                                                      ref_k_stride, row_map, k_map, accumulate);
  TPG_CHECK_LAUNCH("unpack_weights");
  return 0;
}

int tpgan_nchw_to_nhwc(const float* src, tpgan_view dst, int32_t round_tf32, void* stream) {
  long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  if (!src || !dst.ptr || total <= 0) return set_error(TPGAN_ERR_INVALID, "nchw_to_nhwc: bad args");
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, ST>>>(src, dv(dst), round_tf32);
  TPG_CHECK_LAUNCH("nchw_to_nhwc");
  return 0;
}
int tpgan_nhwc_to_nchw(tpgan_view src, float* dst, void* stream) {
  long long total = (long long)src.n * src.h * src.w * src.c;
  if (!dst || !src.ptr || total <= 0) return set_error(TPGAN_ERR_INVALID, "nhwc_to_nchw: bad args");
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(src), dst);
  TPG_CHECK_LAUNCH("nhwc_to_nchw");
  return 0;
}

int tpgan_act_backward(tpgan_view src, tpgan_view mask, tpgan_view dst, const float* slopes, float slope, void* stream) {
  if (!same_geom(src, dst) || !same_geom(mask, dst)) return set_error(TPGAN_ERR_INVALID, "act_backward: geometry mismatch");
  long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  act_backward_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(src), dv(mask), dv(dst), slopes, slope);
  TPG_CHECK_LAUNCH("act_backward");
  return 0;
}
int tpgan_view_copy(tpgan_view src, tpgan_view dst, int32_t accumulate, void* stream) {
  if (!same_geom(src, dst)) return set_error(TPGAN_ERR_INVALID, "view_copy: geometry mismatch");
  long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  view_copy_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(src), dv(dst), accumulate);
  TPG_CHECK_LAUNCH("view_copy");
  return 0;
}
int tpgan_bias_grad(tpgan_view dy, float* db, int32_t accumulate, void* stream) {
  if (!dy.ptr || !db) return set_error(TPGAN_ERR_INVALID, "bias_grad: bad args");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * dy.c, ST);
    if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  long long npix = (long long)dy.n * dy.h * dy.w;
  const bool dense = dy.sh == (int64_t)dy.w * dy.sw && dy.sn == (int64_t)dy.h * dy.sh && (dy.sw % 4 == 0) &&
                     (((uintptr_t)dy.ptr & 15) == 0);
  if (dense) {
    // channels beyond dy.c inside the last float4 are padding lanes of the same buffer (always readable)
    dim3 grid((unsigned)std::max(1ll, std::min((npix + 31) / 32, 8ll * 148)), (unsigned)((dy.c + 127) / 128));
    if (tpg::g_deterministic.load(std::memory_order_relaxed)) grid.x = 1;   // one block per channel group: no atomics race
    bias_grad_dense_kernel<<<grid, 256, 0, ST>>>(dy.ptr, npix, dy.sw, dy.c, db);
    TPG_CHECK_LAUNCH("bias_grad");
    return 0;
  }
  dim3 grid((unsigned)std::max(1ll, std::min((npix + 63) / 64, 4ll * 148)), (unsigned)((dy.c + 31) / 32));
  if (tpg::g_deterministic.load(std::memory_order_relaxed)) grid.x = 1;
  bias_grad_kernel<<<grid, 256, 0, ST>>>(dv(dy), db);
  TPG_CHECK_LAUNCH("bias_grad");
  return 0;
}

int tpgan_reflect_pad(tpgan_view src, tpgan_view dst, int32_t left, int32_t top, void* stream) {
  if (src.n != dst.n || src.c != dst.c || dst.h < src.h + top || dst.w < src.w + left)
    return set_error(TPGAN_ERR_INVALID, "reflect_pad: geometry mismatch");
  long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  reflect_pad_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(src), dv(dst), left, top);
  TPG_CHECK_LAUNCH("reflect_pad");
  return 0;
}
int tpgan_reflect_pad_backward(tpgan_view dpad, tpgan_view dsrc, int32_t left, int32_t top, int32_t accumulate,
                               void* stream) {
  if (dsrc.n != dpad.n || dsrc.c != dpad.c || dpad.h < dsrc.h + top || dpad.w < dsrc.w + left)
    return set_error(TPGAN_ERR_INVALID, "reflect_pad_backward: geometry mismatch");
  long long total = (long long)dsrc.n * dsrc.h * dsrc.w * dsrc.c;
  reflect_pad_backward_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(dpad), dv(dsrc), left, top, accumulate);
  TPG_CHECK_LAUNCH("reflect_pad_backward");
  return 0;
}

int tpgan_patch_crop(tpgan_view img, const float* landmarks, tpgan_view left_eye, tpgan_view right_eye, tpgan_view nose,
                     tpgan_view mouth, int32_t* boxes, float fill, void* stream) {
  const tpgan_view* o[4] = {&left_eye, &right_eye, &nose, &mouth};
  const int pw[4] = {40, 40, 40, 48}, ph[4] = {40, 40, 32, 32};
  CropOut co;
  for (int i = 0; i < 4; ++i) {
    if (o[i]->w != pw[i] || o[i]->h != ph[i] || o[i]->n != img.n || o[i]->c != img.c)
      return set_error(TPGAN_ERR_INVALID, "patch_crop: patch %d must be (N,%d,%d,C)", i, ph[i], pw[i]);
    co.v[i] = dv(*o[i]);
  }
  if (!landmarks) return set_error(TPGAN_ERR_INVALID, "patch_crop: landmarks NULL");
  long long rows = (long long)img.n * 144;
  int grid = (int)std::max(1ll, std::min((rows + 7) / 8, 8ll * 148));
  auto px4 = [](const tpgan_view& v) {
    return v.sw == 4 && v.c <= 4 && (((uintptr_t)v.ptr) & 15) == 0 && v.sh % 4 == 0 && v.sn % 4 == 0;
  };
  int vec = px4(img) ? 1 : 0;
  for (int i = 0; i < 4; ++i) vec = vec && px4(*o[i]);
  patch_crop_kernel<<<grid, 256, 0, ST>>>(dv(img), landmarks, co, boxes, fill, vec);
  TPG_CHECK_LAUNCH("patch_crop");
  return 0;
}

int tpgan_local_fuse(tpgan_view left_eye, tpgan_view right_eye, tpgan_view nose, tpgan_view mouth, tpgan_view out,
                     uint8_t* argmax, void* stream) {
  const tpgan_view* in[4] = {&left_eye, &right_eye, &nose, &mouth};
  const int pw[4] = {40, 40, 40, 48}, ph[4] = {40, 40, 32, 32};
  FuseIn fi;
  for (int i = 0; i < 4; ++i) {
    if (in[i]->w != pw[i] || in[i]->h != ph[i] || in[i]->n != out.n || in[i]->c != out.c)
      return set_error(TPGAN_ERR_INVALID, "local_fuse: patch %d must be (N,%d,%d,C)", i, ph[i], pw[i]);
    fi.v[i] = dv(*in[i]);
  }
  if (out.h != 128 || out.w != 128) return set_error(TPGAN_ERR_INVALID, "local_fuse: output must be 128x128");
  long long total = (long long)out.n * out.h * out.w * out.c;
  bool vec = (out.c % 4 == 0) && view_vec4_ok(out) && (((uintptr_t)argmax & 3) == 0);
  for (int i = 0; i < 4; ++i) vec = vec && view_vec4_ok(*in[i]);
  if (vec)
    local_fuse_vec4_kernel<<<grid_for(total / 4, 256), 256, 0, ST>>>(fi, dv(out), argmax);
  else
    local_fuse_kernel<<<grid_for(total, 256), 256, 0, ST>>>(fi, dv(out), argmax);
  TPG_CHECK_LAUNCH("local_fuse");
  return 0;
}
int tpgan_local_fuse_backward(tpgan_view dout, const uint8_t* argmax, tpgan_view d_left_eye, tpgan_view d_right_eye,
                              tpgan_view d_nose, tpgan_view d_mouth, int32_t accumulate, void* stream) {
  const tpgan_view* in[4] = {&d_left_eye, &d_right_eye, &d_nose, &d_mouth};
  const int pw[4] = {40, 40, 40, 48}, ph[4] = {40, 40, 32, 32};
  FuseIn fi;
  long long total = 0;
  for (int i = 0; i < 4; ++i) {
    if (in[i]->w != pw[i] || in[i]->h != ph[i] || in[i]->n != dout.n || in[i]->c != dout.c)
      return set_error(TPGAN_ERR_INVALID, "local_fuse_backward: patch %d must be (N,%d,%d,C)", i, ph[i], pw[i]);
    fi.v[i] = dv(*in[i]);
    total += (long long)in[i]->n * in[i]->h * in[i]->w * in[i]->c;
  }
  if (!argmax) return set_error(TPGAN_ERR_INVALID, "local_fuse_backward: argmax NULL");
  bool vec = (dout.c % 4 == 0) && view_vec4_ok(dout) && (((uintptr_t)argmax & 3) == 0) && dout.h == 128 && dout.w == 128;
  for (int i = 0; i < 4; ++i) vec = vec && view_vec4_ok(*in[i]);
  if (vec)
    local_fuse_backward_vec4_kernel<<<grid_for(total / 4, 256), 256, 0, ST>>>(dv(dout), argmax, fi, accumulate);
  else
    local_fuse_backward_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(dout), argmax, fi, accumulate);
  TPG_CHECK_LAUNCH("local_fuse_backward");
  return 0;
}

int tpgan_image_losses(tpgan_view fake, tpgan_view t128, tpgan_view t64, tpgan_view t32, tpgan_view dfake, const float* w,
                       float* sums, void* stream) {
  if (!same_geom(fake, t128) || !same_geom(fake, dfake) || t64.h * 2 != fake.h || t32.h * 4 != fake.h || (fake.w % 4) ||
      (fake.h % 4))
    return set_error(TPGAN_ERR_INVALID, "image_losses: geometry mismatch");
  LossW W;
  for (int i = 0; i < 8; ++i) W.w[i] = w[i];
  long long total = (long long)fake.n * fake.h * fake.w * fake.c;
  auto vec4 = [](const tpgan_view& v) {   // a pixel is one aligned float4 whose lanes beyond c are layout padding
    return v.sw == 4 && v.c <= 4 && (((uintptr_t)v.ptr) & 15) == 0 && v.sh % 4 == 0 && v.sn % 4 == 0;
  };
  const char* gen_ev = getenv("TPGAN_IMAGE_LOSSES_GENERIC");   // per call: the parity test compares the two kernels
  if (!(gen_ev && atoi(gen_ev)) && vec4(fake) && vec4(t128) && vec4(t64) && vec4(t32) && vec4(dfake) && fake.w <= 256 && t64.w * 2 == fake.w &&
      t32.w * 4 == fake.w && t64.c == fake.c && t32.c == fake.c) {
    const int threads = 128;   // small blocks: ~8 bands in flight per SM hide the load -> sync -> load latency chain
    const size_t smem = (size_t)(6 * fake.w + fake.w + fake.w / 4) * sizeof(float4);
    image_losses_tiled_kernel<<<fake.n * (fake.h / 4), threads, smem, ST>>>(dv(fake), dv(t128), dv(t64), dv(t32), dv(dfake), W, sums);
    TPG_CHECK_LAUNCH("image_losses_tiled");
    return 0;
  }
  image_losses_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(fake), dv(t128), dv(t64), dv(t32), dv(dfake), W, sums);
  TPG_CHECK_LAUNCH("image_losses");
  return 0;
}
int tpgan_l1_loss(tpgan_view a, tpgan_view b, tpgan_view da, float coeff, float* sum, void* stream) {
  if (!same_geom(a, b)) return set_error(TPGAN_ERR_INVALID, "l1_loss: geometry mismatch");
  long long total = (long long)a.n * a.h * a.w * a.c;
  l1_loss_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(a), dv(b), dv(da), coeff, sum);
  TPG_CHECK_LAUNCH("l1_loss");
  return 0;
}

int tpgan_maxout2(const float* x, float* y, int32_t rows, int32_t cols_out, void* stream) {
  long long n = (long long)rows * cols_out;
  maxout2_kernel<<<grid_for(n, 256), 256, 0, ST>>>(x, y, n);
  TPG_CHECK_LAUNCH("maxout2");
  return 0;
}
int tpgan_maxout2_backward(const float* x, const float* dy, float* dx, int32_t rows, int32_t cols_out, void* stream) {
  long long n = (long long)rows * cols_out;
  maxout2_backward_kernel<<<grid_for(n, 256), 256, 0, ST>>>(x, dy, dx, n);
  TPG_CHECK_LAUNCH("maxout2_backward");
  return 0;
}

int tpgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) return set_error(TPGAN_ERR_INVALID, "adam_step: bad args");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n, 256, 16), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                   (float)sqrt(bc2), grad_scale);
  TPG_CHECK_LAUNCH("adam_step");
  return 0;
}

int tpgan_sample_sqnorm(tpgan_view g, float* sqnorm, void* stream) {
  sample_sqnorm_kernel<<<g.n, 1024, 0, ST>>>(dv(g), sqnorm);
  TPG_CHECK_LAUNCH("sample_sqnorm");
  return 0;
}
int tpgan_sample_scale(tpgan_view g, const float* coeff, tpgan_view u, void* stream) {
  if (!same_geom(g, u)) return set_error(TPGAN_ERR_INVALID, "sample_scale: geometry mismatch");
  long long total = (long long)g.n * g.h * g.w * g.c;
  sample_scale_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(g), coeff, dv(u));
  TPG_CHECK_LAUNCH("sample_scale");
  return 0;
}

int tpgan_mul(tpgan_view a, tpgan_view b, tpgan_view out, void* stream) {
  if (!same_geom(a, b) || !same_geom(a, out)) return set_error(TPGAN_ERR_INVALID, "mul: geometry mismatch");
  long long total = (long long)a.n * a.h * a.w * a.c;
  mul_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(a), dv(b), dv(out));
  TPG_CHECK_LAUNCH("mul");
  return 0;
}
int tpgan_lerp(tpgan_view a, tpgan_view b, const float* alpha, tpgan_view out, void* stream) {
  if (!same_geom(a, b) || !same_geom(a, out)) return set_error(TPGAN_ERR_INVALID, "lerp: geometry mismatch");
  long long total = (long long)a.n * a.h * a.w * a.c;
  lerp_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(a), dv(b), alpha, dv(out));
  TPG_CHECK_LAUNCH("lerp");
  return 0;
}
int tpgan_gp_coeff(const float* sqnorm, float* coeff, int32_t n, float scale, float* gp_sum, void* stream) {
  if (n < 1) return set_error(TPGAN_ERR_INVALID, "gp_coeff: n < 1");
  gp_coeff_kernel<<<1, 256, 0, ST>>>(sqnorm, coeff, n, scale, gp_sum);
  TPG_CHECK_LAUNCH("gp_coeff");
  return 0;
}
int tpgan_fill(tpgan_view out, float value, void* stream) {
  long long total = (long long)out.n * out.h * out.w * out.c;
  if (total <= 0) return 0;
  fill_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(out), value);
  TPG_CHECK_LAUNCH("fill");
  return 0;
}
int tpgan_softmax_ce(const float* logits, int64_t row_stride, const int64_t* labels, float* dlogits, int64_t drow_stride,
                     int32_t rows, int32_t cols, float coeff, float* loss_sum, void* stream) {
  if (rows < 1 || cols < 1) return set_error(TPGAN_ERR_INVALID, "softmax_ce: empty");
  const int wpb = 4;
  softmax_ce_kernel<<<(rows + wpb - 1) / wpb, wpb * 32, 0, ST>>>(logits, row_stride, (const long long*)labels, dlogits,
                                                                 drow_stride, rows, cols, coeff, loss_sum,
                                                                 tpg::device_status_word());
  TPG_CHECK_LAUNCH("softmax_ce");
  return 0;
}

int tpgan_split_tf32(tpgan_view src, tpgan_view hi, tpgan_view lo, void* stream) {
  if (!same_geom(src, hi) || !same_geom(src, lo)) return set_error(TPGAN_ERR_INVALID, "split_tf32: geometry mismatch");
  long long total = (long long)src.n * src.h * src.w * src.c;
  if (total <= 0) return 0;
  split_tf32_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(src), dv(hi), dv(lo));
  TPG_CHECK_LAUNCH("split_tf32");
  return 0;
}

static int adam_dev_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int32_t* step_dev, float grad_scale, int increment, void* stream);
int tpgan_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int32_t* step_dev, float grad_scale, void* stream) {
  return adam_dev_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale, 1, stream);
}
int tpgan_adam_slice_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int32_t* step_dev, float grad_scale, int32_t increment, void* stream) {
  return adam_dev_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale, increment, stream);
}
static int adam_dev_launch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int32_t* step_dev, float grad_scale, int increment, void* stream) {
  if (!p || !g || !m || !v || !step_dev || n <= 0) return set_error(TPGAN_ERR_INVALID, "adam_step_dev: bad args");
  if (increment) {
    counter_inc_kernel<<<1, 1, 0, ST>>>(step_dev);
    TPG_CHECK_LAUNCH("counter_inc");
  }
  const bool vec = (n % 4 == 0) && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  if (vec)
    adam_dev_vec4_kernel<<<grid_for(n / 4, 256, 16), 256, 0, ST>>>(
        reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale);
  else
    adam_dev_kernel<<<grid_for(n, 256, 16), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev,
                                                          grad_scale);
  TPG_CHECK_LAUNCH("adam_step_dev");
  return 0;
}

int tpgan_transpose_packed(const float* src, float* dst, int32_t taps, int32_t rows, int32_t k, int32_t rows_src_pad,
                           int32_t k_src_pad, int32_t rows_dst_pad, int32_t k_dst_pad, void* stream) {
  if (!src || !dst || taps < 1 || rows < 1 || k < 1 || rows > rows_src_pad || k > k_src_pad || k > rows_dst_pad ||
      rows > k_dst_pad)
    return set_error(TPGAN_ERR_INVALID, "transpose_packed: bad args");
  dim3 grid((unsigned)((k + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)taps), block(32, 8);
  transpose_packed_kernel<<<grid, block, 0, ST>>>(src, dst, rows, k, rows_src_pad, k_src_pad, rows_dst_pad, k_dst_pad);
  TPG_CHECK_LAUNCH("transpose_packed");
  return 0;
}

int tpgan_cast_bf16(tpgan_view src, tpgan_view dst16, void* stream) {
  if (!src.ptr || !dst16.ptr || !same_geom(src, dst16)) return set_error(TPGAN_ERR_INVALID, "cast_bf16: bad views");
  const long long total = (long long)src.n * src.h * src.w * ((src.c + 3) / 4);
  if (total <= 0) return 0;
  const bool vec = view_vec4_ok(src) && (((uintptr_t)dst16.ptr & 7) == 0) && dst16.sn % 4 == 0 && dst16.sh % 4 == 0 &&
                   dst16.sw % 4 == 0 && (src.sw >= ((src.c + 3) / 4) * 4) && (dst16.sw >= ((src.c + 3) / 4) * 4);
  cast_bf16_kernel<<<grid_for(total, 256, 16), 256, 0, ST>>>(dv(src), reinterpret_cast<uint16_t*>(dst16.ptr), dst16.sn,
                                                              dst16.sh, dst16.sw, vec ? 1 : 0);
  TPG_CHECK_LAUNCH("cast_bf16");
  return 0;
}
int tpgan_cast_packed_multi(const tpgan_cast_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream) {
  if (!jobs_dev || njobs <= 0 || total_blocks <= 0) return set_error(TPGAN_ERR_INVALID, "cast_packed_multi: bad args");
  cast_packed_multi_kernel<<<total_blocks, 256, 0, ST>>>(jobs_dev, njobs);
  TPG_CHECK_LAUNCH("cast_packed_multi");
  return 0;
}
int tpgan_bias_grad_multi(const tpgan_bias_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream) {
  if (!jobs_dev || njobs < 1 || total_blocks < 1) return set_error(TPGAN_ERR_INVALID, "bias_grad_multi: bad args");
  bias_grad_multi_kernel<<<total_blocks, 256, 0, ST>>>(jobs_dev, njobs);
  TPG_CHECK_LAUNCH("bias_grad_multi");
  return 0;
}
int tpgan_pack_multi(const tpgan_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, int32_t max_row_len, int32_t unpack,
                     void* stream) {
  if (!jobs_dev || njobs < 1 || total_blocks < 1 || max_row_len < 1 || max_row_len > 48 * 1024)
    return set_error(TPGAN_ERR_INVALID, "pack_multi: bad args");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(pack_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    attr_set = true;
  }
  pack_multi_kernel<<<total_blocks, kPackThreads, (size_t)max_row_len * 4, ST>>>(jobs_dev, njobs, unpack ? 1 : 0);
  TPG_CHECK_LAUNCH("pack_multi");
  return 0;
}
int tpgan_transpose_multi(const tpgan_transpose_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream) {
  if (!jobs_dev || njobs < 1 || total_blocks < 1) return set_error(TPGAN_ERR_INVALID, "transpose_multi: bad args");
  transpose_multi_kernel<<<(total_blocks + kMultiRun - 1) / kMultiRun, dim3(32, 8), 0, ST>>>(jobs_dev, njobs, total_blocks);
  TPG_CHECK_LAUNCH("transpose_multi");
  return 0;
}

int tpgan_maxpool3s2(tpgan_view x, tpgan_view y, uint8_t* argmax, void* stream) {
  if (x.n != y.n || x.c != y.c || y.h != (x.h + 2 - 3) / 2 + 1 || y.w != (x.w + 2 - 3) / 2 + 1)
    return set_error(TPGAN_ERR_INVALID, "maxpool3s2: geometry mismatch");
  long long total = (long long)y.n * y.h * y.w * y.c;
  maxpool3s2_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(x), dv(y), argmax);
  TPG_CHECK_LAUNCH("maxpool3s2");
  return 0;
}
int tpgan_maxpool3s2_backward(tpgan_view dy, const uint8_t* argmax, tpgan_view dx, int32_t accumulate, void* stream) {
  if (!argmax || dx.n != dy.n || dx.c != dy.c) return set_error(TPGAN_ERR_INVALID, "maxpool3s2_backward: bad args");
  long long total = (long long)dx.n * dx.h * dx.w * dx.c;
  maxpool3s2_backward_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(dy), argmax, dv(dx), accumulate);
  TPG_CHECK_LAUNCH("maxpool3s2_backward");
  return 0;
}
int tpgan_avgpool(tpgan_view x, tpgan_view y, void* stream) {
  if (x.n != y.n || x.c != y.c || y.h != 1 || y.w != 1) return set_error(TPGAN_ERR_INVALID, "avgpool: geometry mismatch");
  avgpool_kernel<<<dim3((unsigned)((x.c + 63) / 64), (unsigned)x.n), 64, 0, ST>>>(dv(x), dv(y));
  TPG_CHECK_LAUNCH("avgpool");
  return 0;
}
int tpgan_avgpool_backward(tpgan_view dy, tpgan_view dx, int32_t accumulate, void* stream) {
  if (dx.n != dy.n || dx.c != dy.c || dy.h != 1 || dy.w != 1) return set_error(TPGAN_ERR_INVALID, "avgpool_backward: bad args");
  long long total = (long long)dx.n * dx.h * dx.w * dx.c;
  avgpool_backward_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dv(dy), dv(dx), accumulate);
  TPG_CHECK_LAUNCH("avgpool_backward");
  return 0;
}

}  // extern "C"
