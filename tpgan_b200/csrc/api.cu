// C-ABI entry points for the tensor-core convolution kernels: host-side planning (tap lists, parity planes, tile
// geometry, TMA descriptors) and launch.  See include/tpgan_b200.h for the contract of every function.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/tpgan_b200.h"
#include "host_common.h"
#include "kparams.h"

namespace tpg {

template <class Params>
__global__ void tapgemm_kernel(const __grid_constant__ Params P, int* status);
template <class Params>
__global__ void wgrad_kernel(const __grid_constant__ Params P, int* status);
__global__ void rowconv_kernel(const __grid_constant__ RowConvParams P, int* status);

// ------------------------------------------------------------------------------------------------ error state
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------------ device state
struct DeviceState {
  bool ready = false;
  int sm_count = 0;
  int max_smem = 0;
  int* status_dev = nullptr;   // device alias of the mapped host status word
  int* status_host = nullptr;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};
static DeviceState g_dev;
static std::mutex g_mu;

static int ensure_device() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev.ready) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return set_error(TPGAN_ERR_NO_DEVICE, "tpgan_b200 needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
  g_dev.sm_count = prop.multiProcessorCount;
  g_dev.max_smem = (int)prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return set_error(TPGAN_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  g_dev.encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  e = cudaHostAlloc((void**)&g_dev.status_host, sizeof(int), cudaHostAllocMapped);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
  *g_dev.status_host = 0;
  e = cudaHostGetDevicePointer((void**)&g_dev.status_dev, g_dev.status_host, 0);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaHostGetDevicePointer: %s", cudaGetErrorString(e));
  g_dev.ready = true;
  return 0;
}

int device_sm_count() { return g_dev.sm_count; }

// ------------------------------------------------------------------------------------------------ TMA descriptors
// 4D NHWC plane: dims {C, W, H, N}; step = parity-plane subsampling factor along H and W.
static int encode_nhwc(CUtensorMap* m, const float* base, int C, int W, int H, int N, long long sw, long long sh,
                       long long sn, int bw, int bh, int bn, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return set_error(TPGAN_ERR_INVALID, "TMA operand must be 16-byte aligned (ptr %p, strides %lld %lld %lld elements)",
                     (const void*)base, sw, sh, sn);
  if (bw > 256 || bh > 256 || bn > 256 || bw < 1 || bh < 1 || bn < 1)
    return set_error(TPGAN_ERR_INVALID, "bad TMA box %d %d %d", bw, bh, bn);
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TPGAN_ERR_CUDA, "cuTensorMapEncodeTiled(4D C=%d W=%d H=%d N=%d box %d,%d,%d) failed: %d", C, W, H, N,
                     bw, bh, bn, (int)r);
  return 0;
}

static int encode_weights(CUtensorMap* m, const float* base, int k_pad, int rows_pad, int taps, int block_n) {
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)rows_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * 4, (cuuint64_t)k_pad * rows_pad * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)block_n, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((uintptr_t)base & 15) return set_error(TPGAN_ERR_INVALID, "packed weights must be 16-byte aligned");
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TPGAN_ERR_CUDA, "cuTensorMapEncodeTiled(weights k=%d rows=%d taps=%d bn=%d) failed: %d", k_pad,
                     rows_pad, taps, block_n, (int)r);
  return 0;
}

static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static DevView to_dev(const tpgan_view& v) { return DevView{v.ptr, v.sn, v.sh, v.sw}; }
static bool view_vec_ok(const tpgan_view& v) {
  return v.ptr == nullptr || (((uintptr_t)v.ptr & 15) == 0 && v.sn % 4 == 0 && v.sh % 4 == 0 && v.sw % 4 == 0);
}

// Parity planes of `t` for sampling stride s: plane (a,b) holds pixels (s*i + a, s*j + b).
static int encode_planes(CUtensorMap* maps, const tpgan_view& t, int s, int bw, int bh, int bn,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  for (int a = 0; a < s; ++a)
    for (int b = 0; b < s; ++b) {
      int Hp = (t.h - a + s - 1) / s, Wp = (t.w - b + s - 1) / s;
      if (Hp <= 0 || Wp <= 0) return set_error(TPGAN_ERR_INVALID, "empty parity plane");
      int rc = encode_nhwc(&maps[a * s + b], t.ptr + a * t.sh + b * t.sw, t.c, Wp, Hp, t.n, t.sw * s, t.sh * s, t.sn,
                           bw, bh, bn, swz);
      if (rc) return rc;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------ conv planning
static int plan_group(const tpgan_conv_args& a, TapGemmGroup& G) {
  memset(&G, 0, sizeof(G));
  const int k = a.kh;
  if (a.kh != a.kw || k < 1 || k > 8) return set_error(TPGAN_ERR_INVALID, "kernel %dx%d unsupported", a.kh, a.kw);
  const int s = a.stride, p = a.pad;
  if (s != 1 && s != 2 && s != 4) return set_error(TPGAN_ERR_INVALID, "stride %d unsupported", s);
  if (a.in.n != a.out.n) return set_error(TPGAN_ERR_INVALID, "batch mismatch");
  const bool gather = (a.kind == TPGAN_CONV_FWD || a.kind == TPGAN_DECONV_DGRAD);
  const bool phased = (a.kind == TPGAN_CONV_DGRAD || a.kind == TPGAN_DECONV_FWD);
  if (!gather && !phased) return set_error(TPGAN_ERR_INVALID, "bad kind %d", a.kind);
  const int Kc = a.in.c;
  const int taps = k * k;
  if (a.w_k_pad % 32 || a.w_k_pad < Kc || a.w_rows_pad % 16 || a.w_rows_pad < a.out.c)
    return set_error(TPGAN_ERR_INVALID, "packed weight dims (%d x %d) do not cover K=%d N=%d", a.w_rows_pad, a.w_k_pad, Kc,
                     a.out.c);

  G.Nimg = a.in.n;
  int ntap = 0;
  if (gather) {
    // out[o] = sum_r in[s*o + r - p] * w[r]
    const int Ho = (a.in.h + 2 * p - k) / s + 1, Wo = (a.in.w + 2 * p - k) / s + 1;
    if (a.kind == TPGAN_CONV_FWD && (Ho != a.out.h || Wo != a.out.w))
      return set_error(TPGAN_ERR_INVALID, "conv output %dx%d expected %dx%d", a.out.h, a.out.w, Ho, Wo);
    G.Hm = a.out.h;
    G.Wm = a.out.w;
    G.out_sy = G.out_sx = 1;
    G.n_phases = 1;
    G.phase[0].tap_begin = 0;
    G.phase[0].oy = G.phase[0].ox = 0;
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        int ey = r - p, ex = c - p;
        TapDesc t;
        t.plane = (int8_t)(pos_mod(ey, s) * s + pos_mod(ex, s));
        t.dy = (int8_t)floor_div(ey, s);
        t.dx = (int8_t)floor_div(ex, s);
        t.wtap = (uint8_t)(r * k + c);
        G.taps[ntap++] = t;
      }
    G.phase[0].tap_count = (short)ntap;
  } else {
    // out[s*m + a] = sum_{r : (a + p - r) % s == 0} in[m + (a + p - r)/s] * w[r]
    if (a.out.h % s || a.out.w % s) return set_error(TPGAN_ERR_INVALID, "output %dx%d not a multiple of stride %d", a.out.h, a.out.w, s);
    G.Hm = a.out.h / s;
    G.Wm = a.out.w / s;
    G.out_sy = G.out_sx = s;
    G.n_phases = s * s;
    for (int pa = 0; pa < s; ++pa)
      for (int pb = 0; pb < s; ++pb) {
        PhaseDesc& ph = G.phase[pa * s + pb];
        ph.tap_begin = (short)ntap;
        ph.oy = (short)pa;
        ph.ox = (short)pb;
        for (int r = 0; r < k; ++r) {
          if (pos_mod(pa + p - r, s)) continue;
          for (int c = 0; c < k; ++c) {
            if (pos_mod(pb + p - c, s)) continue;
            if (ntap >= kMaxTaps) return set_error(TPGAN_ERR_INVALID, "too many taps");
            TapDesc t;
            t.plane = 0;
            t.dy = (int8_t)((pa + p - r) / s);
            t.dx = (int8_t)((pb + p - c) / s);
            t.wtap = (uint8_t)(r * k + c);
            G.taps[ntap++] = t;
          }
        }
        if (ntap == ph.tap_begin) {  // hole phase (stride > kernel): bias only, through the all-zero tap
          TapDesc t;
          t.plane = 0; t.dy = 0; t.dx = 0; t.wtap = (uint8_t)taps;
          G.taps[ntap++] = t;
        }
        ph.tap_count = (short)(ntap - ph.tap_begin);
      }
  }
  if (G.Wm > 128) return set_error(TPGAN_ERR_INVALID, "tile-space width %d > 128 unsupported", G.Wm);
  G.bw = G.Wm;
  G.bh = std::min(G.Hm, 128 / G.bw);
  G.bn = (G.bh == G.Hm) ? std::max(1, std::min(G.Nimg, 128 / (G.bw * G.bh))) : 1;
  G.tiles_h = ceil_div(G.Hm, G.bh);
  G.m_tiles = G.tiles_h * ceil_div(G.Nimg, G.bn);
  G.n_tiles = ceil_div(a.w_rows_pad, 256);
  G.block_n = ceil_div(ceil_div(a.w_rows_pad, G.n_tiles), 16) * 16;
  G.kchunks = ceil_div(Kc, 32);
  G.last_mmas = ceil_div(Kc - 32 * (G.kchunks - 1), 8);
  G.tile_count = G.n_phases * G.m_tiles * G.n_tiles;

  int rc = gather ? encode_planes(G.amap, a.in, s, G.bw, G.bh, G.bn) : encode_planes(G.amap, a.in, 1, G.bw, G.bh, G.bn);
  if (rc) return rc;
  rc = encode_weights(&G.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, taps + 1, G.block_n);
  if (rc) return rc;

  G.out = to_dev(a.out);
  G.add1 = to_dev(a.add1);
  G.add2 = to_dev(a.add2);
  G.mask = to_dev(a.mask);
  G.bias = a.bias;
  G.slopes = a.slopes;
  G.Hout = a.out.h;
  G.Wout = a.out.w;
  G.cout_valid = a.out.c;
  G.epilogue = a.epilogue;
  G.slope = a.slope;
  G.round_tf32 = a.round_tf32;
  if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) return set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view");
  G.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask);
  return 0;
}

template <class Params>
static int launch_tapgemm(Params& P, cudaStream_t st) {
  int bmax = 0;
  int tiles = 0;
  for (int i = 0; i < P.ngroups; ++i) {
    P.g[i].tile_begin = tiles;
    tiles += P.g[i].tile_count;
    bmax = std::max(bmax, P.g[i].block_n * 128);
  }
  P.total_tiles = tiles;
  P.b_stage_bytes = bmax;
  const int budget = g_dev.max_smem - 1024 - 256;
  // two 32-float K chunks per stage when at least 3 such stages fit (halves the per-stage barrier/issue overhead)
  P.kst = (budget / (2 * (16384 + bmax)) >= 3) ? 2 : 1;
  if (const char* ev = getenv("TPGAN_KST")) P.kst = std::max(1, std::min(2, atoi(ev)));
  const int stage_bytes = P.kst * (16384 + bmax);
  P.stages = std::min(kMaxStages, budget / stage_bytes);
  if (const char* ev = getenv("TPGAN_STAGES")) P.stages = std::min(P.stages, std::max(2, atoi(ev)));
  if (P.stages < 2) return set_error(TPGAN_ERR_INVALID, "not enough shared memory for 2 stages");
  const int smem = P.stages * stage_bytes + 1024;
  static std::once_flag once;
  auto kern = tapgemm_kernel<Params>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 256);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int grid = std::min(tiles, g_dev.sm_count);
  if (const char* ev = getenv("TPGAN_GRID")) grid = std::min(grid, std::max(1, atoi(ev)));
  kern<<<grid, 256, smem, st>>>(P, g_dev.status_dev);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "tapgemm launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// ------------------------------------------------------------------------------------------------ row-tile conv
// Eligible: Conv2d forward / input gradient, stride 1, "same" padding (2p == k-1), width exactly 128 (one GEMM tile = one
// image row), i.e. every full-resolution layer of the global pathway.  Returns 1 when launched, 0 when not eligible.
static int try_rowconv(const tpgan_conv_args& a, cudaStream_t st, int* rc_out) {
  *rc_out = 0;
  static const bool disabled = getenv("TPGAN_NO_ROWCONV") != nullptr;
  if (disabled) return 0;
  if (a.kind != TPGAN_CONV_FWD && a.kind != TPGAN_CONV_DGRAD) return 0;
  const int k = a.kh, p = a.pad;
  if (a.kh != a.kw || a.stride != 1 || 2 * p != k - 1 || k < 3 || k * k > kMaxTaps) return 0;
  if (a.in.w != 128 || a.out.w != 128 || a.in.h != a.out.h || a.in.n != a.out.n) return 0;
  if (a.w_rows_pad > 256) return 0;
  static thread_local RowConvParams P;
  memset(&P, 0, sizeof(P));
  P.H = a.out.h; P.W = 128; P.Nimg = a.in.n; P.k = k;
  P.n_tiles = 1;
  P.block_n = ceil_div(a.w_rows_pad, 16) * 16;
  // T output rows share every weight tile; accumulators stay double-buffered (T * block_n <= 256 columns per buffer) -
  // measured: a single-buffered wider tile loses more to the epilogue bubble than it gains in weight reuse
  int T = std::max(1, std::min(4, 256 / P.block_n));
  T = std::max(1, std::min(T, P.H));
  if (const char* ev = getenv("TPGAN_ROWCONV_T")) T = std::max(1, std::min(atoi(ev), 512 / P.block_n));
  P.T = T;
  P.double_buf = (P.block_n * T <= 256) ? 1 : 0;
  P.row_tiles = ceil_div(P.H, T);
  P.total_tiles = P.Nimg * P.row_tiles * P.n_tiles;
  P.kchunks = ceil_div(a.in.c, 32);
  P.last_mmas = ceil_div(a.in.c - 32 * (P.kchunks - 1), 8);
  const bool fwd = a.kind == TPGAN_CONV_FWD;
  P.dy0 = fwd ? -p : p - k + 1;
  P.dx0 = P.dy0;
  for (int r = 0; r < k; ++r)
    for (int j = 0; j < k; ++j) {
      P.wtap[r * k + j] = (unsigned char)(fwd ? (r * k + j) : ((k - 1 - r) * k + (k - 1 - j)));
      P.dxoff[r * k + j] = (unsigned char)j;
    }
  P.slab_bytes = ceil_div((128 + k - 1) * 128, 1024) * 1024;
  P.b_bytes = P.block_n * 128;
  const int budget = g_dev.max_smem - 1024 - 1024;   // alignment slack + this kernel's static shared memory
  P.a_slots = std::min(16, T + 2);
  P.b_slots = std::min(16, (budget - P.a_slots * P.slab_bytes) / P.b_bytes);
  if (P.b_slots < 3) {
    P.a_slots = T + 1;
    P.b_slots = std::min(16, (budget - P.a_slots * P.slab_bytes) / P.b_bytes);
    if (P.b_slots < 2) return 0;
  }
  int rc = encode_nhwc(&P.amap, a.in.ptr, a.in.c, a.in.w, a.in.h, a.in.n, a.in.sw, a.in.sh, a.in.sn, 128 + k - 1, 1, 1);
  if (rc) { *rc_out = rc; return 1; }
  rc = encode_weights(&P.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, k * k + 1, P.block_n);
  if (rc) { *rc_out = rc; return 1; }
  P.out = to_dev(a.out); P.add1 = to_dev(a.add1); P.add2 = to_dev(a.add2); P.mask = to_dev(a.mask);
  P.bias = a.bias; P.slopes = a.slopes;
  P.cout_valid = a.out.c; P.epilogue = a.epilogue; P.slope = a.slope; P.round_tf32 = a.round_tf32;
  if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) { *rc_out = set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view"); return 1; }
  P.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask);
  const int smem = P.a_slots * P.slab_bytes + P.b_slots * P.b_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(rowconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 1024);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
  const int grid = std::min(P.total_tiles, g_dev.sm_count);
  rowconv_kernel<<<grid, 256, smem, st>>>(P, g_dev.status_dev);
  e = cudaGetLastError();
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "rowconv launch: %s", cudaGetErrorString(e)); return 1; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 1;
}

// ------------------------------------------------------------------------------------------------ wgrad planning
static int plan_wgrad(const tpgan_wgrad_args& a, WgradGroup& G, int box_px) {
  memset(&G, 0, sizeof(G));
  const int k = a.kh;
  if (a.kh != a.kw || k < 1 || k > 8) return set_error(TPGAN_ERR_INVALID, "kernel %dx%d unsupported", a.kh, a.kw);
  const int s = a.stride, p = a.pad;
  if (s != 1 && s != 2 && s != 4) return set_error(TPGAN_ERR_INVALID, "stride %d unsupported", s);
  if (a.x.n != a.dy.n) return set_error(TPGAN_ERR_INVALID, "batch mismatch");
  // conv  : P = dy (Cout), Q = x planes  (Cin)  -> dw[tap][co][ci], m=co n=ci
  // deconv: P = x  (Cin),  Q = dy planes (Cout) -> dw[tap][co][ci], m=ci n=co (transposed write)
  const bool is_conv = (a.kind == TPGAN_CONV_FWD);
  if (!is_conv && a.kind != TPGAN_DECONV_FWD) return set_error(TPGAN_ERR_INVALID, "bad wgrad kind %d", a.kind);
  const tpgan_view& Pt = is_conv ? a.dy : a.x;
  const tpgan_view& Qt = is_conv ? a.x : a.dy;
  G.transpose_out = is_conv ? 0 : 1;
  G.m_valid = Pt.c;
  G.n_valid = Qt.c;
  const int cout = a.dy.c, cin = a.x.c;
  if (a.w_rows_pad < cout || a.w_k_pad < cin) return set_error(TPGAN_ERR_INVALID, "packed dw too small");
  G.rows_pad = a.w_rows_pad;
  G.k_pad = a.w_k_pad;
  G.dw = a.dw_packed;
  G.accumulate = a.accumulate;
  G.Hp = Pt.h;
  G.Wp = Pt.w;
  G.Nimg = Pt.n;
  // pixel boxes of ~32 pixels
  G.bw = (G.Wp <= box_px + box_px / 2) ? G.Wp : box_px;
  G.bh = std::max(1, std::min(G.Hp, box_px / G.bw));
  G.bn = (G.bh == G.Hp && G.bw == G.Wp) ? std::max(1, std::min(G.Nimg, box_px / (G.bw * G.bh))) : 1;
  G.kp = ceil_div(G.bw * G.bh * G.bn, 8) * 8;
  G.tiles_w = ceil_div(G.Wp, G.bw);
  G.tiles_h = ceil_div(G.Hp, G.bh);
  G.chunks = G.tiles_w * G.tiles_h * ceil_div(G.Nimg, G.bn);
  G.m_tiles = ceil_div(G.m_valid, 128);
  const int nch_total = ceil_div(G.n_valid, 32);
  G.ntaps = k * k;
  static const bool no_pack = getenv("TPGAN_WGRAD_NOPACK") != nullptr;
  static const int tap_pack_max = getenv("TPGAN_WGRAD_TAPPACK") ? atoi(getenv("TPGAN_WGRAD_TAPPACK")) : 1;
  if (nch_total <= tap_pack_max && G.ntaps > 1 && !no_pack) {
    // narrow shifted tensor (one 32-channel chunk: the 3-channel image layers): several taps side by side in N, all
    // sharing the loads of P (measured: pays only for a single chunk; wider tensors become L2-bound)
    G.ncpt = nch_total;
    G.tpu = std::min(G.ntaps, 8 / nch_total);
    G.n_tiles = 1;
    G.block_n = G.tpu * G.ncpt * 32;
    G.mpu = 1;
  } else {
    G.tpu = 1;
    G.n_tiles = ceil_div(nch_total, 8);
    G.block_n = ceil_div(nch_total, G.n_tiles) * 32;
    G.ncpt = G.block_n / 32;
    // several M tiles per unit (one accumulator each), all sharing the loads of Q - when the reduction is long enough
    // (>= 4096 pixels per image batch) to amortise the then single-buffered epilogue
    const long long npix = (long long)G.Hp * G.Wp * G.Nimg;
    G.mpu = (no_pack || npix < 32768) ? 1 : std::max(1, std::min(G.m_tiles, 512 / G.block_n));
  }
  G.tap_groups = ceil_div(G.ntaps, G.tpu);
  G.mt_groups = ceil_div(G.m_tiles, G.mpu);
  G.nbuf = (G.mpu * G.block_n <= 256) ? 2 : 1;
  int nt = 0;
  for (int r = 0; r < k; ++r)
    for (int c = 0; c < k; ++c) {
      int ey = r - p, ex = c - p;
      TapDesc t;
      t.plane = (int8_t)(pos_mod(ey, s) * s + pos_mod(ex, s));
      t.dy = (int8_t)floor_div(ey, s);
      t.dx = (int8_t)floor_div(ex, s);
      t.wtap = (uint8_t)(r * k + c);
      G.taps[nt++] = t;
    }
  // MN-major tf32 operands need the 32B-atom flavour of the 128B swizzle
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  int rc = encode_nhwc(&G.pmap, Pt.ptr, Pt.c, Pt.w, Pt.h, Pt.n, Pt.sw, Pt.sh, Pt.sn, G.bw, G.bh, G.bn, swz);
  if (rc) return rc;
  rc = encode_planes(G.qmap, Qt, s, G.bw, G.bh, G.bn, swz);
  if (rc) return rc;
  return 0;
}

// Pixels per pipeline stage: the largest of {128, 64, 32} that still leaves >= 4 stages in shared memory for every group
// (few, large TMA boxes and 8-16 MMAs per barrier round trip; measured 1.8x on the 64-channel 128x128 layers).
static int choose_wgrad_px(const tpgan_wgrad_args* groups, int ngroups) {
  if (const char* ev = getenv("TPGAN_WGRAD_PX")) return std::max(8, std::min(128, atoi(ev)));
  const int budget = g_dev.max_smem - 1024 - 256;
  int chunks = 0;
  static const bool no_pack = getenv("TPGAN_WGRAD_NOPACK") != nullptr;
  for (int i = 0; i < ngroups; ++i) {
    const bool is_conv = groups[i].kind == TPGAN_CONV_FWD;
    const int pc = is_conv ? groups[i].dy.c : groups[i].x.c, qc = is_conv ? groups[i].x.c : groups[i].dy.c;
    const int ntaps = groups[i].kh * groups[i].kw;
    const int m_tiles = ceil_div(pc, 128);
    const int nch_total = ceil_div(qc, 32);
    int a_ch, b_ch;
    static const int tap_pack_max = getenv("TPGAN_WGRAD_TAPPACK") ? atoi(getenv("TPGAN_WGRAD_TAPPACK")) : 1;
    const tpgan_view& Pt = is_conv ? groups[i].dy : groups[i].x;
    const long long npix = (long long)Pt.h * Pt.w * Pt.n;
    if (nch_total <= tap_pack_max && ntaps > 1 && !no_pack) {
      a_ch = std::min(4, ceil_div(pc, 32));
      b_ch = std::min(ntaps, 8 / nch_total) * nch_total;
    } else {
      b_ch = ceil_div(nch_total, ceil_div(nch_total, 8));
      const int mpu = (no_pack || npix < 32768) ? 1 : std::max(1, std::min(m_tiles, 512 / (b_ch * 32)));
      a_ch = std::min(4 * mpu, ceil_div(pc, 32));
    }
    chunks = std::max(chunks, a_ch + b_ch);
  }
  for (int px : {128, 64})
    if (budget / (chunks * px * 128) >= 4) return px;
  if (budget / (chunks * 32 * 128) >= 3) return 32;
  return 16;
}

template <class Params>
static int launch_wgrad(Params& P, cudaStream_t st) {
  int amax = 0, bmax = 0, base_units = 0;
  for (int i = 0; i < P.ngroups; ++i) {
    WgradGroup& G = P.g[i];
    amax = std::max(amax, 4 * G.mpu * G.kp * 128);
    bmax = std::max(bmax, (G.block_n / 32) * G.kp * 128);
    base_units += G.tap_groups * G.mt_groups * G.n_tiles;
  }
  // split the pixel reduction so that ~2 waves of units cover the SMs
  int units = 0;
  for (int i = 0; i < P.ngroups; ++i) {
    WgradGroup& G = P.g[i];
    int want = std::max(1, ceil_div(2 * g_dev.sm_count, std::max(1, base_units)));
    int ks = std::min(want, G.chunks);
    G.chunks_per_split = ceil_div(G.chunks, ks);
    G.ksplits = ceil_div(G.chunks, G.chunks_per_split);
    G.unit_begin = units;
    G.unit_count = G.tap_groups * G.mt_groups * G.n_tiles * G.ksplits;
    units += G.unit_count;
  }
  P.total_units = units;
  P.nbuf = 2;
  for (int i = 0; i < P.ngroups; ++i) P.nbuf = std::min(P.nbuf, P.g[i].nbuf);
  P.a_stage_bytes = amax;
  P.b_stage_bytes = bmax;
  const int stage_bytes = amax + bmax;
  const int budget = g_dev.max_smem - 1024 - 256;
  P.stages = std::min(kMaxStages, budget / stage_bytes);
  if (P.stages < 2) return set_error(TPGAN_ERR_INVALID, "wgrad: not enough shared memory for 2 stages (%d B/stage)", stage_bytes);
  const int smem = P.stages * stage_bytes + 1024;
  auto kern = wgrad_kernel<Params>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 256);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = std::min(units, g_dev.sm_count);
  kern<<<grid, 256, smem, st>>>(P, g_dev.status_dev);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace tpg

using namespace tpg;

extern "C" {

int tpgan_conv2d(const tpgan_conv_args* groups, int32_t ngroups, void* stream) {
  if (!groups || ngroups < 1 || ngroups > kMaxGroups) return set_error(TPGAN_ERR_INVALID, "ngroups must be 1..%d", kMaxGroups);
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (ngroups == 1) {
    int rrc = 0;
    if (try_rowconv(groups[0], st, &rrc)) return rrc;
    static thread_local TapGemmParams1 P;
    P.ngroups = 1;
    rc = plan_group(groups[0], P.g[0]);
    if (rc) return rc;
    return launch_tapgemm(P, st);
  }
  static thread_local TapGemmParams P;
  P.ngroups = ngroups;
  for (int i = 0; i < ngroups; ++i) {
    rc = plan_group(groups[i], P.g[i]);
    if (rc) return rc;
  }
  return launch_tapgemm(P, st);
}

int tpgan_conv2d_wgrad(const tpgan_wgrad_args* groups, int32_t ngroups, void* stream) {
  if (!groups || ngroups < 1 || ngroups > kMaxGroups) return set_error(TPGAN_ERR_INVALID, "ngroups must be 1..%d", kMaxGroups);
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (ngroups == 1) {
    static thread_local WgradParams1 P;
    P.ngroups = 1;
    rc = plan_wgrad(groups[0], P.g[0], choose_wgrad_px(groups, 1));
    if (rc) return rc;
    return launch_wgrad(P, st);
  }
  static thread_local WgradParams P;
  P.ngroups = ngroups;
  bool uniform = true;
  const int px = choose_wgrad_px(groups, ngroups);
  for (int i = 0; i < ngroups; ++i) {
    rc = plan_wgrad(groups[i], P.g[i], px);
    if (rc) return rc;
    const WgradGroup& G = P.g[i];
    // the zero-padded K rows of the smem ring stay zero only if every box fills its rows or all boxes are alike
    if ((G.bw * G.bh * G.bn) % 8) uniform = false;
  }
  if (!uniform) {
    bool same = true;
    for (int i = 1; i < ngroups; ++i)
      same = same && P.g[i].bw == P.g[0].bw && P.g[i].bh == P.g[0].bh && P.g[i].bn == P.g[0].bn &&
             P.g[i].block_n == P.g[0].block_n && P.g[i].mpu == P.g[0].mpu && P.g[i].tpu == P.g[0].tpu;
    if (!same) {
      for (int i = 0; i < ngroups; ++i) {
        rc = tpgan_conv2d_wgrad(groups + i, 1, stream);
        if (rc) return rc;
      }
      return 0;
    }
  }
  return launch_wgrad(P, st);
}

const char* tpgan_last_error(void) { return g_err; }
int tpgan_abi_version(void) { return TPGAN_ABI_VERSION; }
int tpgan_kernel_status(void) {
  if (!g_dev.status_host) return 0;
  int v = *(volatile int*)g_dev.status_host;
  if (v) *(volatile int*)g_dev.status_host = 0;
  return v;
}
int64_t tpgan_launch_count(void) { return (int64_t)g_launches.load(); }

}  // extern "C"
