// Device-side input pipeline of the TP-GAN step: what TrainDataset / TestDataset do per sample on
// the host with PIL + numpy (DataAndDataset.py:179-256, UtilityMethods.py:146-164) as three small launches per batch:
//   * uint8 HWC image -> fp32 NHWC in [-1, 1]  (transforms.ToTensor() then *2.0 - 1.0, DataAndDataset.py:214-220,251-255),
//   * 68-point landmark list -> the 5 key points (get_5_landmarks_pixal_position) with the 128/width, 128/height rescale
//     of TestDataset.__getitem__ (:242-245),
//   * the 64x64 / 32x32 targets as average pools of the 128x128 image (the oracle step's convention, oracle/step.py).
// Patch cropping (process()) is tpgan_patch_crop in pointwise.cu.  All HBM-bound, a few MB per batch.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdint.h>

#include "../../include/tpgan_b200.h"
#include "common.cuh"
#include "host_common.h"

namespace tpg {

struct IV {
  float* p;
  long long sn, sh, sw;
  int n, h, w, c;
};
static inline IV iv(const tpgan_view& v) { return IV{v.ptr, v.sn, v.sh, v.sw, v.n, v.h, v.w, v.c}; }
static inline int igrid(long long n, int block) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)std::max(1, device_sm_count() ? device_sm_count() : 148) * 8;
  return (int)std::max(1ll, std::min(want, cap));
}

// ToTensor(): float32(byte) / 255 ; then * 2.0 - 1.0 - three separately rounded fp32 operations, as torch executes them
__global__ void u8_to_nhwc_kernel(const uint8_t* __restrict__ src, IV dst, int round) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dst.c);
    long long r = i / dst.c;
    const int x = (int)(r % dst.w);
    r /= dst.w;
    const int y = (int)(r % dst.h);
    const int n = (int)(r / dst.h);
    const float v = __fsub_rn(__fmul_rn(__fdiv_rn((float)src[i], 255.0f), 2.0f), 1.0f);
    dst.p[(long long)n * dst.sn + (long long)y * dst.sh + (long long)x * dst.sw + c] = round ? round_tf32(v) : v;
  }
}

// out[n][j] = mean of pts[n][lo_j .. hi_j] (inclusive, clipped to the list; empty -> NaN like np.mean), accumulated
// sequentially in fp32 as numpy's axis-0 reduction does, then divided by the count; x scaled by sx, y by sy.
__global__ void landmarks_reduce_kernel(const float* __restrict__ pts, int npts, const int* __restrict__ ranges, int nr,
                                        float sx, float sy, float* __restrict__ out, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * nr) return;
  const int n = i / nr, j = i % nr;
  const int lo = max(ranges[2 * j], 0), hi = min(ranges[2 * j + 1], npts - 1);
  float ax = 0.f, ay = 0.f;
  int cnt = 0;
  for (int k = lo; k <= hi; ++k, ++cnt) {
    ax = __fadd_rn(ax, pts[((long long)n * npts + k) * 2]);
    ay = __fadd_rn(ay, pts[((long long)n * npts + k) * 2 + 1]);
  }
  const float mx = cnt ? __fdiv_rn(ax, (float)cnt) : NAN, my = cnt ? __fdiv_rn(ay, (float)cnt) : NAN;
  out[(long long)i * 2] = __fmul_rn(mx, sx);
  out[(long long)i * 2 + 1] = __fmul_rn(my, sy);
}

// d2 = 2x2 average pool of src, d4 = 4x4 average pool (one thread per 4x4 block and channel: src is read once)
__global__ void pyramid_kernel(IV src, IV d2, IV d4) {
  const long long total = (long long)d4.n * d4.h * d4.w * d4.c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % d4.c);
    long long r = i / d4.c;
    const int x = (int)(r % d4.w);
    r /= d4.w;
    const int y = (int)(r % d4.h);
    const int n = (int)(r / d4.h);
    float s4 = 0.f;
#pragma unroll
    for (int by = 0; by < 2; ++by)
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {
        float s2 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx)
            s2 += src.p[(long long)n * src.sn + (long long)(4 * y + 2 * by + dy) * src.sh +
                        (long long)(4 * x + 2 * bx + dx) * src.sw + c];
        d2.p[(long long)n * d2.sn + (long long)(2 * y + by) * d2.sh + (long long)(2 * x + bx) * d2.sw + c] = s2 * 0.25f;
        s4 += s2;
      }
    d4.p[(long long)n * d4.sn + (long long)y * d4.sh + (long long)x * d4.sw + c] = s4 * 0.0625f;
  }
}

}  // namespace tpg

using namespace tpg;
#define ST ((cudaStream_t)stream)

extern "C" {

int tpgan_u8_to_nhwc(const uint8_t* src, tpgan_view dst, int32_t round_tf32, void* stream) {
  const long long total = (long long)dst.n * dst.h * dst.w * dst.c;
  if (!src || !dst.ptr || total <= 0) return set_error(TPGAN_ERR_INVALID, "u8_to_nhwc: bad args");
  u8_to_nhwc_kernel<<<igrid(total, 256), 256, 0, ST>>>(src, iv(dst), round_tf32);
  TPG_CHECK_LAUNCH("u8_to_nhwc");
  return 0;
}

int tpgan_landmarks_reduce(const float* points, int32_t batch, int32_t npoints, const int32_t* ranges_dev, int32_t nranges,
                           float scale_x, float scale_y, float* out, void* stream) {
  if (!points || !ranges_dev || !out || batch < 1 || npoints < 1 || nranges < 1)
    return set_error(TPGAN_ERR_INVALID, "landmarks_reduce: bad args");
  landmarks_reduce_kernel<<<(batch * nranges + 127) / 128, 128, 0, ST>>>(points, npoints, ranges_dev, nranges, scale_x, scale_y,
                                                                         out, batch);
  TPG_CHECK_LAUNCH("landmarks_reduce");
  return 0;
}

int tpgan_pyramid(tpgan_view src, tpgan_view half, tpgan_view quarter, void* stream) {
  if (!src.ptr || !half.ptr || !quarter.ptr || src.h % 4 || src.w % 4 || half.h * 2 != src.h || half.w * 2 != src.w ||
      quarter.h * 4 != src.h || quarter.w * 4 != src.w || half.c != src.c || quarter.c != src.c || half.n != src.n ||
      quarter.n != src.n)
    return set_error(TPGAN_ERR_INVALID, "pyramid: geometry mismatch");
  const long long total = (long long)quarter.n * quarter.h * quarter.w * quarter.c;
  pyramid_kernel<<<igrid(total, 256), 256, 0, ST>>>(iv(src), iv(half), iv(quarter));
  TPG_CHECK_LAUNCH("pyramid");
  return 0;
}

}  // extern "C"
