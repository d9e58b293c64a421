// C-ABI entry points for the tensor-core convolution kernels: host-side planning (tap lists, parity planes, tile
// geometry, TMA descriptors) and launch.  See include/tpgan_b200.h for the contract of every function.
#include <cmath>
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

template <class Params, bool BF16, bool PAIR>
__global__ void tapgemm_kernel(const __grid_constant__ Params P, int* status);
template <class Params, bool BF16>
__global__ void wgrad_kernel(const __grid_constant__ Params P, int* status);
template <bool BF16>
__global__ void flatconv_kernel(const __grid_constant__ FlatConvParams P, int* status);
template <bool BF16>
__global__ void rowconv_kernel(const __grid_constant__ RowConvParams P, int* status);
template <bool BF16>
__global__ void rowstack_kernel(const __grid_constant__ RowStackParams P, int* status);

// Operand geometry by dtype: channels per 128-byte shared-memory row and elements per MMA K step (common.cuh: Opnd).
static inline int chunk_ch(int bf16) { return bf16 ? 64 : 32; }
static inline int mma_k(int bf16) { return bf16 ? 16 : 8; }

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
  void* split_ws = nullptr;    // split-K workspace ring (see kSplitSlots); null = split-K unavailable
  int* split_cnt = nullptr;
  void* im2col_ws = nullptr;   // padded-input ring of the im2col-by-TMA weight gradients; null = that mode is off
};
static DeviceState g_dev;
static std::mutex g_mu;
// Deterministic mode (tpgan_set_deterministic): split-K weight-gradient reductions and multi-block bias sums combine their
// partial results with fp32 atomics, whose order varies from run to run.  With the flag set, a weight-gradient CTA owns
// whole output tiles and a bias sum is one block per channel group: bit-identical results run to run, at a lower speed.
std::atomic<int> g_deterministic{0};
// SMs left free by the persistent tensor-core kernels (tpgan_set_sm_reserve).  These kernels run one CTA per SM with a
// static tile partition; when a concurrent kernel on another stream (an NCCL collective overlapped with backward) holds k
// SMs, k CTAs of the next launch cannot start until an SM frees up and the launch can take up to twice as long.  With a
// reserve >= the collective's CTA count both run side by side at (sm_count - k) / sm_count of the tensor throughput.
std::atomic<int> g_sm_reserve{0};
// Split-K workspace: a ring of slots (partial accumulators + per-tile arrival counters, self-resetting).  Consecutive
// launches take consecutive slots; a slot is reused only after kSplitSlots further split launches, i.e. never while an
// earlier user can still be in flight unless that many whole-GPU launches overlap across streams.
constexpr int kSplitSlots = 8;
constexpr size_t kSplitSlotBytes = 4u << 20;
constexpr int kSplitCounters = 256;
std::atomic<unsigned> g_split_seq{0};
// Zero-padded float4-per-pixel copies of <= 4-channel conv inputs for the im2col-by-TMA weight gradients (same ring
// discipline: written and read by consecutive launches of one stream, reused kIm2colSlots launches later).
constexpr int kIm2colSlots = 4;
constexpr size_t kIm2colSlotBytes = 32u << 20;
std::atomic<unsigned> g_im2col_seq{0};

static inline int persistent_sms() { return std::max(1, g_dev.sm_count - g_sm_reserve.load(std::memory_order_relaxed)); }

// Tensor-core kernels are launched as programmatic dependents of the kernel before them in the stream (captured into CUDA
// graphs as programmatic edges): their CTAs become resident as the previous grid's CTAs exit and run the prologue early;
// griddepcontrol.wait in the kernel orders every global-memory access after the previous grid's completion.
// TPGAN_PDL=0 falls back to plain stream serialisation (for A/B measurements).
std::atomic<int> g_pdl{-1};
template <class Kern, class Params>
static cudaError_t launch_tc(Kern kern, int grid, int smem, cudaStream_t st, const Params& P, int cluster = 1) {
  int pdl = g_pdl.load(std::memory_order_relaxed);
  if (pdl < 0) {
    const char* ev = getenv("TPGAN_PDL");
    pdl = (ev && atoi(ev) == 0) ? 0 : 1;
    g_pdl.store(pdl, std::memory_order_relaxed);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)kConvThreads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster > 1) {   // CTA pairs (tcgen05 cta_group::2): grid is a multiple of the cluster size
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  int* status = g_dev.status_dev;
  void* args[2] = {const_cast<Params*>(&P), &status};
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), args);
}

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
  {   // split-K workspace ring + zeroed arrival counters (the kernels leave every counter at zero again)
    const size_t ws_bytes = (size_t)kSplitSlots * kSplitSlotBytes, cnt_bytes = (size_t)kSplitSlots * kSplitCounters * sizeof(int);
    void* p = nullptr;
    if (cudaMalloc(&p, ws_bytes + cnt_bytes) == cudaSuccess && cudaMemset((char*)p + ws_bytes, 0, cnt_bytes) == cudaSuccess) {
      g_dev.split_ws = p;
      g_dev.split_cnt = reinterpret_cast<int*>((char*)p + ws_bytes);
    } else {
      (void)cudaGetLastError();
    }
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)kIm2colSlots * kIm2colSlotBytes) == cudaSuccess) g_dev.im2col_ws = q;
    else (void)cudaGetLastError();
  }
  g_dev.ready = true;
  return 0;
}

int device_sm_count() { return g_dev.sm_count; }
int* device_status_word() { return ensure_device() == 0 ? g_dev.status_dev : nullptr; }

// ------------------------------------------------------------------------------------------------ TMA descriptors
// 4D NHWC plane: dims {C, W, H, N}; step = parity-plane subsampling factor along H and W.
static int encode_nhwc(CUtensorMap* m, const float* base, int C, int W, int H, int N, long long sw, long long sh,
                       long long sn, int bw, int bh, int bn, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                       int bc = 32, int bf16 = 0) {
  // bf16: `base` addresses 2-byte elements (the caller passes the view's ptr); strides are in elements
  const cuuint64_t eb = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * eb, (cuuint64_t)sh * eb, (cuuint64_t)sn * eb};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return set_error(TPGAN_ERR_INVALID, "TMA operand must be 16-byte aligned (ptr %p, strides %lld %lld %lld elements)",
                     (const void*)base, sw, sh, sn);
  if (bw > 256 || bh > 256 || bn > 256 || bw < 1 || bh < 1 || bn < 1)
    return set_error(TPGAN_ERR_INVALID, "bad TMA box %d %d %d", bw, bh, bn);
  CUresult r = g_dev.encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TPGAN_ERR_CUDA, "cuTensorMapEncodeTiled(4D C=%d W=%d H=%d N=%d box %d,%d,%d) failed: %d", C, W, H, N,
                     bw, bh, bn, (int)r);
  return 0;
}

static int encode_weights(CUtensorMap* m, const float* base, int k_pad, int rows_pad, int taps, int block_n,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int bc = 32, int bf16 = 0) {
  const cuuint64_t eb = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)rows_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * eb, (cuuint64_t)k_pad * rows_pad * eb};
  cuuint32_t box[3] = {(cuuint32_t)bc, (cuuint32_t)block_n, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((uintptr_t)base & 15) return set_error(TPGAN_ERR_INVALID, "packed weights must be 16-byte aligned");
  CUresult r = g_dev.encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
static DevView16 to_dev16(const tpgan_view& v) { return DevView16{reinterpret_cast<uint16_t*>(v.ptr), v.sn, v.sh, v.sw}; }
// bf16 copy of the output: 8-byte stores of 4 channels need an 8-byte aligned view with strides that are multiples of 4
static bool view16_ok(const tpgan_view& v) {
  return v.ptr == nullptr || (((uintptr_t)v.ptr & 7) == 0 && v.sn % 4 == 0 && v.sh % 4 == 0 && v.sw % 4 == 0);
}
// dtype / output-view consistency shared by the three forward-type planners
static int check_dtype(const tpgan_conv_args& a) {
  if (a.dtype != TPGAN_DTYPE_TF32 && a.dtype != TPGAN_DTYPE_BF16) return set_error(TPGAN_ERR_INVALID, "bad dtype %d", a.dtype);
  if (a.dtype == TPGAN_DTYPE_TF32 && a.out16.ptr) return set_error(TPGAN_ERR_INVALID, "out16 needs dtype BF16");
  if (!a.out.ptr && !a.out16.ptr) return set_error(TPGAN_ERR_INVALID, "conv2d: no output view");
  if (a.out16.ptr && (a.out16.n != a.out.n || a.out16.h != a.out.h || a.out16.w != a.out.w || a.out16.c != a.out.c))
    return set_error(TPGAN_ERR_INVALID, "out16 geometry differs from out");
  return 0;
}
static bool view_vec_ok(const tpgan_view& v) {
  return v.ptr == nullptr || (((uintptr_t)v.ptr & 15) == 0 && v.sn % 4 == 0 && v.sh % 4 == 0 && v.sw % 4 == 0);
}

// Parity planes of `t` for sampling stride s: plane (a,b) holds pixels (s*i + a, s*j + b).
static int encode_planes(CUtensorMap* maps, const tpgan_view& t, int s, int bw, int bh, int bn,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int bf16 = 0) {
  for (int a = 0; a < s; ++a)
    for (int b = 0; b < s; ++b) {
      int Hp = (t.h - a + s - 1) / s, Wp = (t.w - b + s - 1) / s;
      if (Hp <= 0 || Wp <= 0) return set_error(TPGAN_ERR_INVALID, "empty parity plane");
      const long long off = a * t.sh + b * t.sw;   // elements
      const float* base = bf16 ? reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(t.ptr) + off) : t.ptr + off;
      int rc = encode_nhwc(&maps[a * s + b], base, t.c, Wp, Hp, t.n, t.sw * s, t.sh * s, t.sn,
                           bw, bh, bn, swz, chunk_ch(bf16), bf16);
      if (rc) return rc;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------ conv planning
// CTA-pair launches (tapgemm_kernel<.., PAIR = true>): set around plan_group / launch_tapgemm by tpgan_conv2d.  The plan of a
// paired group differs in two places: the weight box is half an N tile (each CTA of the pair stages block_n / 2 rows) and the
// tile list counts PAIRS of consecutive M tiles.
static thread_local bool t_pair = false;
static thread_local int g_last_conv_pair = 0;
static int pair_mode() {   // TPGAN_PAIR: 0 = never, 1 = whenever eligible (default)
  static const int mode = [] { const char* ev = getenv("TPGAN_PAIR"); return ev ? atoi(ev) : 1; }();
  return mode;
}

static int plan_group(const tpgan_conv_args& a, TapGemmGroup& G, int n_split = 1) {
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
  int rc0 = check_dtype(a);
  if (rc0) return rc0;
  const int bf16 = a.dtype == TPGAN_DTYPE_BF16;
  const int CH = chunk_ch(bf16);
  const int cout = a.out.c;   // out's geometry is always valid; only its ptr may be null (bf16 mode without an fp32 copy)
  if (a.w_k_pad % CH || a.w_k_pad < Kc || a.w_rows_pad % 16 || a.w_rows_pad < cout)
    return set_error(TPGAN_ERR_INVALID, "packed weight dims (%d x %d) do not cover K=%d N=%d", a.w_rows_pad, a.w_k_pad, Kc,
                     cout);

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
  // One-launch 3xTF32: every phase's tap list three times - (in_hi, w_lo), (in_lo, w_hi), (in_hi, w_hi); the in_lo planes
  // follow the in_hi planes in amap[], bit 7 of wtap selects the residual weight packing.
  const bool exact3 = a.in_lo.ptr != nullptr;
  const int nplanes = gather ? s * s : 1;
  if (exact3) {
    if (bf16 || !a.w_lo_packed) return set_error(TPGAN_ERR_INVALID, "in_lo needs TF32 operands and w_lo_packed");
    if (ntap * 3 > kMaxTaps || nplanes * 2 > kMaxPlanes || taps + 1 > 0x7f)
      return set_error(TPGAN_ERR_INVALID, "3xTF32 in one launch: %d taps x 3 exceed the tap table", ntap);
    if (a.in_lo.n != a.in.n || a.in_lo.h != a.in.h || a.in_lo.w != a.in.w || a.in_lo.c != a.in.c)
      return set_error(TPGAN_ERR_INVALID, "in_lo geometry differs from in");
    TapDesc old[kMaxTaps];
    memcpy(old, G.taps, sizeof(old));
    int nt3 = 0;
    for (int ph = 0; ph < G.n_phases; ++ph) {
      const int b0 = G.phase[ph].tap_begin, cnt = G.phase[ph].tap_count;
      G.phase[ph].tap_begin = (short)nt3;
      for (int v = 0; v < 3; ++v)
        for (int t = 0; t < cnt; ++t) {
          TapDesc d = old[b0 + t];
          if (v == 1) d.plane = (int8_t)(d.plane + nplanes);
          if (v == 0) d.wtap = (uint8_t)(d.wtap | 0x80);
          G.taps[nt3++] = d;
        }
      G.phase[ph].tap_count = (short)(3 * cnt);
    }
    ntap = nt3;
  }
  if (G.Wm > 128) return set_error(TPGAN_ERR_INVALID, "tile-space width %d > 128 unsupported", G.Wm);
  G.bw = G.Wm;
  G.bh = std::min(G.Hm, 128 / G.bw);
  G.bn = (G.bh == G.Hm) ? std::max(1, std::min(G.Nimg, 128 / (G.bw * G.bh))) : 1;
  G.tiles_h = ceil_div(G.Hm, G.bh);
  G.m_tiles = G.tiles_h * ceil_div(G.Nimg, G.bn);
  G.n_tiles = n_split < 0 ? -n_split : ceil_div(a.w_rows_pad, 256) * n_split;   // n_split < 0: explicit tile count
  G.block_n = ceil_div(ceil_div(a.w_rows_pad, G.n_tiles), 16) * 16;
  G.n_tiles = ceil_div(a.w_rows_pad, G.block_n);
  G.kchunks = ceil_div(Kc, CH);
  G.last_mmas = ceil_div(Kc - CH * (G.kchunks - 1), mma_k(bf16));
  G.tile_count = G.n_phases * (t_pair ? ceil_div(G.m_tiles, 2) : G.m_tiles) * G.n_tiles;
  G.ksplit = 1;
  G.kc_per = G.kchunks;
  G.kt_per = 0;

  int rc = encode_planes(G.amap, a.in, gather ? s : 1, G.bw, G.bh, G.bn, CU_TENSOR_MAP_SWIZZLE_128B, bf16);
  if (rc) return rc;
  const int wbox = t_pair ? G.block_n / 2 : G.block_n;
  rc = encode_weights(&G.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, taps + 1, wbox, CU_TENSOR_MAP_SWIZZLE_128B, CH, bf16);
  if (rc) return rc;
  if (exact3) {
    rc = encode_planes(G.amap + nplanes, a.in_lo, gather ? s : 1, G.bw, G.bh, G.bn, CU_TENSOR_MAP_SWIZZLE_128B, 0);
    if (rc) return rc;
    rc = encode_weights(&G.bmap_lo, a.w_lo_packed, a.w_k_pad, a.w_rows_pad, taps + 1, wbox, CU_TENSOR_MAP_SWIZZLE_128B, CH, 0);
    if (rc) return rc;
  }

  G.out = to_dev(a.out);
  G.out16 = to_dev16(a.out16);
  G.add1 = to_dev(a.add1);
  G.add2 = to_dev(a.add2);
  G.mask = to_dev(a.mask);
  G.bias = a.bias;
  G.slopes = a.slopes;
  G.Hout = a.out.h;
  G.Wout = a.out.w;
  G.cout_valid = cout;
  G.epilogue = a.epilogue;
  G.slope = a.slope;
  G.round_tf32 = a.round_tf32;
  if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) return set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view");
  G.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask) && view16_ok(a.out16);
  return 0;
}

// Narrower N tiles for launches that cannot fill the SMs (small feature maps): modelled makespan = waves of tiles x
// (K steps x cycles of one M=128 MMA at that N + a fixed per-tile cost).  Returns the factor by which to multiply n_tiles.
template <class Params>
static int choose_n_split(const Params& P, const tpgan_conv_args* groups) {
  // Returns the number of N tiles per 256 output channels... see choose_n_tiles below (kept: the factor form used by the
  // grouped launches, powers of two)
  static const bool off = getenv("TPGAN_NO_NSPLIT") != nullptr;
  if (off) return 1;
  double best_cost = 0;
  int best = 1;
  for (int f : {1, 2, 4, 8}) {
    double tiles = 0, worst_tile = 0;
    bool ok = true;
    for (int i = 0; i < P.ngroups; ++i) {
      const TapGemmGroup& G = P.g[i];
      const int base_tiles = ceil_div(groups[i].w_rows_pad, 256);
      const int bn = ceil_div(ceil_div(groups[i].w_rows_pad, base_tiles * f), 16) * 16;
      if (f > 1 && bn < 32) { ok = false; break; }
      const int nt = ceil_div(groups[i].w_rows_pad, bn);
      int taps = 0;
      for (int ph = 0; ph < G.n_phases; ++ph) taps += G.phase[ph].tap_count;
      const double ksteps = (double)taps / G.n_phases * G.kchunks * 4;
      const double per_tile = ksteps * std::max(bn / 2.0, 32.0 + bn / 4.0) + 4000.0;
      tiles += (double)G.n_phases * G.m_tiles * nt;
      worst_tile = std::max(worst_tile, per_tile);
    }
    if (!ok) break;
    const double cost = std::ceil(tiles / persistent_sms()) * worst_tile;
    if (f == 1 || cost < best_cost * 0.9) { best_cost = cost; best = f; }
  }
  return best;
}

// Single launches: any number of N tiles (not only powers of two times the 256-column minimum).  Modelled makespan =
// waves of tiles x (K steps x cycles per K step + a fixed per-tile cost); a K step costs the MMA (N/2 clocks) or the
// operand bytes it needs at what a CTA's ~190 KB in flight sustain (~72 B/clk), whichever is larger - the model that
// reproduces the measured tensor-pipe utilisation of these launches.  The point of the
// finer search is wave quantisation: enhance_features_16 (768 channels, 64 M tiles) runs 192 tiles of 256 columns on 148
// SMs, i.e. two waves for 1.3; 256 tiles of 192 columns fill both.  Returns 0 to keep the default tiling.
static int choose_n_tiles(const TapGemmGroup& G, const tpgan_conv_args& a) {
  static const bool off = getenv("TPGAN_NO_NSPLIT") != nullptr;
  if (off) return 0;
  const int base = ceil_div(a.w_rows_pad, 256);
  int taps = 0;
  for (int ph = 0; ph < G.n_phases; ++ph) taps += G.phase[ph].tap_count;
  const double ksteps = (double)taps / G.n_phases * G.kchunks * 4;
  double best_cost = 0;
  int best = 0;
  // CTA pairs (TPGAN_PAIR_MODEL=1): the schedulable unit is a pair of M tiles on a pair of SMs, and a CTA stages half the
  // weight bytes per K step
  static const bool pair_model = getenv("TPGAN_PAIR_MODEL") != nullptr && atoi(getenv("TPGAN_PAIR_MODEL")) != 0;
  const bool pm = pair_model && pair_mode() && G.m_tiles >= 2;
  for (int nt = base; nt <= base * 8; ++nt) {
    const int bn = ceil_div(ceil_div(a.w_rows_pad, nt), 16) * 16;
    if (nt > base && bn < 32) break;
    if (ceil_div(a.w_rows_pad, bn) != nt) continue;   // this count collapses to a smaller one after rounding to 16 columns
    const double per_kstep = std::max(bn / 2.0, (4096.0 + (pm ? 16.0 : 32.0) * bn) / 72.0);
    const double per_tile = ksteps * per_kstep + 4000.0;
    const double tiles = (double)G.n_phases * (pm ? ceil_div(G.m_tiles, 2) : G.m_tiles) * nt;
    const double cost = std::ceil(tiles / (pm ? std::max(1, persistent_sms() / 2) : persistent_sms())) * per_tile;
    if (best == 0 || cost < best_cost * 0.95) { best_cost = cost; best = nt; }
  }
  return best == base ? 0 : best;
}

template <class Params>
static int launch_tapgemm(Params& P, cudaStream_t st, int bf16) {
  int bmax = 0;
  int tiles = 0;
  for (int i = 0; i < P.ngroups; ++i) {
    P.g[i].tile_begin = tiles;
    tiles += P.g[i].tile_count;
    bmax = std::max(bmax, P.g[i].block_n * (t_pair ? 64 : 128));
  }
  P.total_tiles = tiles;
  P.b_stage_bytes = bmax;
  const int budget = g_dev.max_smem - 1024 - 256;
  // two 32-float K chunks per stage when at least 3 such stages fit (halves the per-stage barrier/issue overhead)
  P.kst = (budget / (2 * (16384 + bmax)) >= 3) ? 2 : 1;
  if (t_pair && P.kst == 2 && budget / (2 * (16384 + bmax)) == 3) {
    // Paired launches of the wide layers: three two-chunk stages whose last stage per tap is half empty (odd chunk count -
    // 208 channels = 6.5 chunks) leave too little in flight; seven single-chunk stages measure 11 % faster there
    // (208->208 k3 @64x64: 583 -> 651 TFLOP/s), while an even chunk count (512 channels) is faster with the big stages.
    bool odd = false;
    for (int i = 0; i < P.ngroups; ++i) odd = odd || (P.g[i].kchunks & 1);
    if (odd) P.kst = 1;
  }
  if (const char* ev = getenv("TPGAN_KST")) P.kst = std::max(1, std::min(2, atoi(ev)));
  const int stage_bytes = P.kst * (16384 + bmax);
  static const int max_stages = [] { const char* ev = getenv("TPGAN_TAP_MAXSTAGES"); return ev ? std::max(2, std::min(kTapMaxStages, atoi(ev))) : kMaxStages; }();
  P.stages = std::min(max_stages, budget / stage_bytes);
  if (const char* ev = getenv("TPGAN_STAGES")) P.stages = std::min(P.stages, std::max(2, atoi(ev)));
  if (P.stages < 2) return set_error(TPGAN_ERR_INVALID, "not enough shared memory for 2 stages");
  const int smem = P.stages * stage_bytes + 1024;
  auto kern = t_pair ? (bf16 ? tapgemm_kernel<Params, true, true> : tapgemm_kernel<Params, false, true>)
                     : (bf16 ? tapgemm_kernel<Params, true, false> : tapgemm_kernel<Params, false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 256);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int grid = t_pair ? 2 * std::min(tiles, std::max(1, persistent_sms() / 2)) : std::min(tiles, persistent_sms());
  if (const char* ev = getenv("TPGAN_GRID")) grid = std::min(grid, std::max(t_pair ? 2 : 1, atoi(ev) & (t_pair ? ~1 : ~0)));
  g_last_conv_pair = t_pair ? 1 : 0;
  e = launch_tc(kern, grid, smem, st, P, t_pair ? 2 : 1);
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
  if (check_dtype(a)) { *rc_out = TPGAN_ERR_INVALID; return 1; }
  const int bf16 = a.dtype == TPGAN_DTYPE_BF16;
  const int CH = chunk_ch(bf16);
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
  P.kchunks = ceil_div(a.in.c, CH);
  P.last_mmas = ceil_div(a.in.c - CH * (P.kchunks - 1), mma_k(bf16));
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
  int rc = encode_nhwc(&P.amap, a.in.ptr, a.in.c, a.in.w, a.in.h, a.in.n, a.in.sw, a.in.sh, a.in.sn, 128 + k - 1, 1, 1,
                       CU_TENSOR_MAP_SWIZZLE_128B, CH, bf16);
  if (rc) { *rc_out = rc; return 1; }
  rc = encode_weights(&P.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, k * k + 1, P.block_n, CU_TENSOR_MAP_SWIZZLE_128B, CH, bf16);
  if (rc) { *rc_out = rc; return 1; }
  P.out = to_dev(a.out); P.add1 = to_dev(a.add1); P.add2 = to_dev(a.add2); P.mask = to_dev(a.mask);
  P.out16 = to_dev16(a.out16);
  P.bias = a.bias; P.slopes = a.slopes;
  P.cout_valid = a.out.c; P.epilogue = a.epilogue; P.slope = a.slope; P.round_tf32 = a.round_tf32;
  if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) { *rc_out = set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view"); return 1; }
  P.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask) && view16_ok(a.out16);
  const int smem = P.a_slots * P.slab_bytes + P.b_slots * P.b_bytes + 1024;
  auto kern = bf16 ? rowconv_kernel<true> : rowconv_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 1024);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
  const int grid = std::min(P.total_tiles, persistent_sms());
  e = launch_tc(kern, grid, smem, st, P);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "rowconv launch: %s", cudaGetErrorString(e)); return 1; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 1;
}

// ------------------------------------------------------------------------------------------------ N-stacked row-tile conv
// Eligible: as try_rowconv, with at most 128 (padded) output channels so that at least two output rows share an MMA.
static int try_rowstack(const tpgan_conv_args& a, cudaStream_t st, int* rc_out) {
  *rc_out = 0;
  static const bool disabled = getenv("TPGAN_NO_ROWSTACK") != nullptr;
  if (disabled) return 0;
  if (a.kind != TPGAN_CONV_FWD && a.kind != TPGAN_CONV_DGRAD) return 0;
  const int k = a.kh, p = a.pad;
  if (a.kh != a.kw || a.stride != 1 || 2 * p != k - 1 || k < 3 || k * k > kMaxTaps) return 0;
  if (a.in.w != 128 || a.out.w != 128 || a.in.h != a.out.h || a.in.n != a.out.n) return 0;
  if (a.w_rows_pad > 128) return 0;
  if (check_dtype(a)) { *rc_out = TPGAN_ERR_INVALID; return 1; }
  const int bf16 = a.dtype == TPGAN_DTYPE_BF16;
  const int CH = chunk_ch(bf16) / 2;   // channels per 64-byte K chunk
  static thread_local RowStackParams P;
  memset(&P, 0, sizeof(P));
  P.H = a.out.h; P.W = 128; P.Nimg = a.in.n; P.k = k;
  P.block_n = ceil_div(a.w_rows_pad, 16) * 16;
  int T = std::max(1, std::min(4, 256 / P.block_n));
  T = std::max(1, std::min(T, P.H));
  if (const char* ev = getenv("TPGAN_ROWSTACK_T")) T = std::max(1, std::min(atoi(ev), 256 / P.block_n));
  if (T < 2) return 0;
  P.T = T;
  P.row_tiles = ceil_div(P.H, T);
  P.total_tiles = P.Nimg * P.row_tiles;
  P.kchunks = ceil_div(a.in.c, CH);
  P.last_mmas = ceil_div(a.in.c - CH * (P.kchunks - 1), mma_k(bf16));
  const bool fwd = a.kind == TPGAN_CONV_FWD;
  P.dy0 = fwd ? -p : p - k + 1;
  P.dx0 = P.dy0;
  for (int r = 0; r < k; ++r)
    for (int j = 0; j < k; ++j)
      P.wtap[r * k + j] = (unsigned char)(fwd ? (r * k + j) : ((k - 1 - r) * k + (k - 1 - j)));
  for (int s = 0; s < T + k - 1; ++s) {
    const int t_lo = std::max(0, s - k + 1), t_hi = std::min(T - 1, s);
    const int nn = (t_hi - t_lo + 1) * P.block_n;
    // instruction descriptor, M = 128, N = nn, K-major A and B, operand format 2 = tf32 / 1 = bf16 (Opnd::idesc in common.cuh)
    const uint32_t fmt = bf16 ? 1u : 2u;
    P.s_idesc[s] = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    P.s_boff[s] = (uint32_t)(k - 1 - (s - t_lo)) * (uint32_t)P.block_n * 4u;
    P.s_doff[s] = (uint32_t)(t_lo * P.block_n);
  }
  P.slab_bytes = ceil_div((128 + k - 1) * 64, 1024) * 1024;
  P.wb_bytes = k * P.block_n * 64;
  const int budget = g_dev.max_smem - 1024 - 1024;   // alignment slack + this kernel's static shared memory
  const int nslab = T + k - 1;
  P.b_slots = 3;
  P.a_slots = std::min(16, (budget - P.b_slots * P.wb_bytes) / P.slab_bytes);
  if (P.a_slots < nslab + 1) {
    P.b_slots = 2;
    P.a_slots = std::min(16, (budget - P.b_slots * P.wb_bytes) / P.slab_bytes);
    if (P.a_slots < nslab) return 0;
  }
  P.a_slots = std::min(P.a_slots, nslab + 4);
  if (nslab > 16) return 0;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_64B;
  int rc = encode_nhwc(&P.amap, a.in.ptr, a.in.c, a.in.w, a.in.h, a.in.n, a.in.sw, a.in.sh, a.in.sn, 128 + k - 1, 1, 1, swz, CH, bf16);
  if (rc) { *rc_out = rc; return 1; }
  rc = encode_weights(&P.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, k * k + 1, P.block_n, swz, CH, bf16);
  if (rc) { *rc_out = rc; return 1; }
  P.out = to_dev(a.out); P.add1 = to_dev(a.add1); P.add2 = to_dev(a.add2); P.mask = to_dev(a.mask);
  P.out16 = to_dev16(a.out16);
  P.bias = a.bias; P.slopes = a.slopes;
  P.cout_valid = a.out.c; P.epilogue = a.epilogue; P.slope = a.slope; P.round_tf32 = a.round_tf32;
  if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) { *rc_out = set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view"); return 1; }
  P.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask) && view16_ok(a.out16);
  const int smem = P.a_slots * P.slab_bytes + P.b_slots * P.wb_bytes + 1024;
  auto kern = bf16 ? rowstack_kernel<true> : rowstack_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 1024);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
  const int grid = std::min(P.total_tiles, persistent_sms());
  e = launch_tc(kern, grid, smem, st, P);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "rowstack launch: %s", cudaGetErrorString(e)); return 1; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 1;
}

// ------------------------------------------------------------------------------------------------ flat-slab conv
// Eligible: Conv2d forward / input gradient, stride 1, "same" padding, k >= 3, maps up to 64 pixels wide - every group of the
// launch.  Opt-in: TPGAN_FLATCONV=1 takes the launches where it measures faster than tapgemm in isolation (13-20 %: at most
// 128 (padded) output channels, or maps up to 12 wide - the 64 / 128 / 256-channel layers of the local pathways, the
// 64..128-channel encoder layers of the global pathway at 64x64 and 32x32), =2 every eligible launch.  Off by default: on
// the whole (power-capped) step the A/B difference is below the run-to-run noise (44.05 vs 44.02, 43.70 vs 43.69 ms).  Returns 1 when launched (or failed with *rc_out set), 0 when not eligible.
struct FlatChoice {
  int T, ur, bn;
  double eff;   // real pixels / M rows computed
};
static FlatChoice choose_flat(int H, int W, int N, int k, int block_n) {
  // every (T, unit shape) candidate; the winner is the LARGEST T (most weight reuse) within 3 % of the best efficiency
  FlatChoice cand[8];
  int nc = 0;
  const int Wp = W + k - 1;
  const int Tmax = std::min(4, 256 / block_n);
  for (int T = 1; T <= Tmax; ++T) {
    const int rows = T * 128;
    FlatChoice c{T, 0, 0, 0.0};
    const int ur = std::min(H, rows / Wp);
    if (ur >= 1 && ur + k - 1 <= 256) c = FlatChoice{T, ur, 1, (double)H * W / ((double)ceil_div(H, ur) * rows)};
    const int ipitch = (H + k - 1) * Wp;
    if (H * Wp <= rows && H + k - 1 <= 256) {
      const int bn = std::min(N, 1 + (rows - H * Wp) / ipitch);
      const double eff = (double)N * H * W / ((double)ceil_div(N, bn) * rows);
      if (bn >= 2 && eff > c.eff) c = FlatChoice{T, H, bn, eff};
    }
    if (c.ur > 0) cand[nc++] = c;
  }
  FlatChoice best{0, 0, 0, 0.0};
  double top = 0.0;
  for (int i = 0; i < nc; ++i) top = std::max(top, cand[i].eff);
  for (int i = 0; i < nc; ++i)
    if (cand[i].eff >= 0.97 * top) best = cand[i];
  return best;
}

static int try_flatconv(const tpgan_conv_args* groups, int ngroups, cudaStream_t st, int* rc_out) {
  *rc_out = 0;
  const char* mode_ev = getenv("TPGAN_FLATCONV");   // read per call: tests switch it between launches
  const int mode = mode_ev ? atoi(mode_ev) : 0;
  if (mode == 0) return 0;
  const tpgan_conv_args& a0 = groups[0];
  if (const char* only = getenv("TPGAN_FLAT_ONLY")) {   // debugging aid: restrict the kernel to one class of launches
    if ((!strcmp(only, "fwd") && a0.kind != TPGAN_CONV_FWD) || (!strcmp(only, "dgrad") && a0.kind != TPGAN_CONV_DGRAD) ||
        (!strcmp(only, "single") && ngroups != 1) || (!strcmp(only, "grouped") && ngroups == 1))
      return 0;
  }
  const int k = a0.kh, p = a0.pad;
  if (a0.kind != TPGAN_CONV_FWD && a0.kind != TPGAN_CONV_DGRAD) return 0;
  if (a0.kh != a0.kw || a0.stride != 1 || 2 * p != k - 1 || k < 3 || k * k > kMaxTaps) return 0;
  const int bf16 = a0.dtype == TPGAN_DTYPE_BF16;
  const int CH = chunk_ch(bf16);
  FlatChoice ch[kMaxGroups];
  int block_n[kMaxGroups], n_tiles[kMaxGroups];
  for (int i = 0; i < ngroups; ++i) {
    const tpgan_conv_args& a = groups[i];
    if (a.kind != a0.kind || a.kh != k || a.kw != k || a.stride != 1 || a.pad != p || a.dtype != a0.dtype) return 0;
    if (a.in.w != a.out.w || a.in.h != a.out.h || a.in.n != a.out.n) return 0;
    if (a.in.w > 64 || a.in.w < 8 || a.in.w + k - 1 > 128) return 0;
    n_tiles[i] = ceil_div(a.w_rows_pad, 256);
    block_n[i] = ceil_div(ceil_div(a.w_rows_pad, n_tiles[i]), 16) * 16;
    n_tiles[i] = ceil_div(a.w_rows_pad, block_n[i]);
    if (mode == 1 && block_n[i] > 128 && a.in.w > 12) return 0;
    ch[i] = choose_flat(a.in.h, a.in.w, a.in.n, k, block_n[i]);
    if (ch[i].T < 1 || ch[i].eff < (mode == 1 ? 0.6 : 0.3)) return 0;
  }
  static thread_local FlatConvParams P;
  memset(&P, 0, sizeof(P));
  P.ngroups = ngroups;
  P.k = k;
  const bool fwd = a0.kind == TPGAN_CONV_FWD;
  P.dy0 = fwd ? -p : p - k + 1;
  P.dx0 = P.dy0;
  for (int r = 0; r < k; ++r)
    for (int j = 0; j < k; ++j)
      P.wtap[r * k + j] = (unsigned char)(fwd ? (r * k + j) : ((k - 1 - r) * k + (k - 1 - j)));
  int tiles = 0, slab_rows = 0, bmax = 0;
  for (int i = 0; i < ngroups; ++i) {
    const tpgan_conv_args& a = groups[i];
    FlatGroup& G = P.g[i];
    int rc = check_dtype(a);
    if (rc) { *rc_out = rc; return 1; }
    if (a.w_k_pad % CH || a.w_k_pad < a.in.c || a.w_rows_pad % 16 || a.w_rows_pad < a.out.c) {
      *rc_out = set_error(TPGAN_ERR_INVALID, "packed weight dims (%d x %d) do not cover K=%d N=%d", a.w_rows_pad, a.w_k_pad, a.in.c, a.out.c);
      return 1;
    }
    if (a.epilogue == TPGAN_EPI_MASK && a.mask.ptr == nullptr) { *rc_out = set_error(TPGAN_ERR_INVALID, "EPI_MASK needs a mask view"); return 1; }
    G.H = a.out.h; G.W = a.out.w; G.Nimg = a.in.n;
    G.Wp = G.W + k - 1;
    G.T = ch[i].T; G.ur = ch[i].ur; G.bn = ch[i].bn;
    G.R = G.ur + k - 1;
    G.ipitch = G.R * G.Wp;
    G.units_h = ceil_div(G.H, G.ur);
    G.n_tiles = n_tiles[i];
    G.block_n = block_n[i];
    G.kchunks = ceil_div(a.in.c, CH);
    G.last_mmas = ceil_div(a.in.c - CH * (G.kchunks - 1), mma_k(bf16));
    G.tile_begin = tiles;
    G.tile_count = G.units_h * ceil_div(G.Nimg, G.bn) * G.n_tiles;
    tiles += G.tile_count;
    G.slab_tx = G.bn * G.R * G.Wp * 128;
    // rows an M = 128 MMA of the last tile / last tap may read (garbage rows beyond the box only reach rows never stored)
    slab_rows = std::max(slab_rows, std::max(G.bn * G.R * G.Wp, G.T * 128 + (k - 1) * G.Wp + (k - 1)));
    bmax = std::max(bmax, G.block_n * 128);
    rc = encode_nhwc(&G.amap, a.in.ptr, a.in.c, a.in.w, a.in.h, a.in.n, a.in.sw, a.in.sh, a.in.sn, G.Wp, G.R, G.bn,
                     CU_TENSOR_MAP_SWIZZLE_128B, CH, bf16);
    if (rc) { *rc_out = rc; return 1; }
    rc = encode_weights(&G.bmap, a.w_packed, a.w_k_pad, a.w_rows_pad, k * k + 1, G.block_n, CU_TENSOR_MAP_SWIZZLE_128B, CH, bf16);
    if (rc) { *rc_out = rc; return 1; }
    G.out = to_dev(a.out); G.add1 = to_dev(a.add1); G.add2 = to_dev(a.add2); G.mask = to_dev(a.mask);
    G.out16 = to_dev16(a.out16);
    G.bias = a.bias; G.slopes = a.slopes;
    G.cout_valid = a.out.c; G.epilogue = a.epilogue; G.slope = a.slope; G.round_tf32 = a.round_tf32;
    G.vec_ok = view_vec_ok(a.out) && view_vec_ok(a.add1) && view_vec_ok(a.add2) && view_vec_ok(a.mask) && view16_ok(a.out16);
  }
  P.total_tiles = tiles;
  P.slab_bytes = ceil_div(slab_rows * 128, 1024) * 1024;
  P.b_bytes = bmax;
  const int budget = g_dev.max_smem - 1024 - 1024;   // alignment slack + this kernel's static shared memory
  // two slab slots (the one in use + the next, requested a whole K chunk ahead by its own producer warp); the rest of the
  // shared memory is the weight ring - the stream whose bytes in flight set the pace
  P.a_slots = 2;
  P.b_slots = std::min(16, (budget - P.a_slots * P.slab_bytes) / P.b_bytes);
  if (P.b_slots < 3) return 0;
  const int smem = P.a_slots * P.slab_bytes + P.b_slots * P.b_bytes + 1024;
  auto kern = bf16 ? flatconv_kernel<true> : flatconv_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 1024);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
  const int grid = std::min(P.total_tiles, persistent_sms());
  e = launch_tc(kern, grid, smem, st, P);
  if (e != cudaSuccess) { *rc_out = set_error(TPGAN_ERR_CUDA, "flatconv launch: %s", cudaGetErrorString(e)); return 1; }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 1;
}

// ------------------------------------------------------------------------------------------------ wgrad planning
// How one weight-gradient problem is mapped onto the kernel (see WgradGroup in kparams.h).
struct WgradChoice {
  bool im2col;    // <= 4 input channels: horizontal taps folded into the 32-lane chunk (see WgradGroup::im2col_k)
  bool swap;      // stride-1 conv only: P = x (M = Cin), Q = dy shifted by -tap (N = Cout), dw written transposed
  bool slab;      // taps of one kernel row share one Q slab (N = taps * 32 per MMA)
  int pc, qc;     // channels of P and Q
  int m_tiles, nch_total, n_tiles, block_n, ncpt, mpu, tpu;
  int a_ch;       // 32-channel P chunks per stage
  int b_ch;       // 32-channel Q chunks per stage (non-slab)
  double cost;    // modelled tensor-pipe cycles per 8-pixel K step over all taps / tiles
};

static const double kL2BytesPerClk = 45.0;   // sustained L2->SM bytes/clk/SM with every SM streaming (measured 35-50)

static WgradChoice wgrad_option(const tpgan_wgrad_args& a, bool swap, bool slab) {
  WgradChoice c{};
  static const bool no_pack = getenv("TPGAN_WGRAD_NOPACK") != nullptr;
  static const int tap_pack_max = getenv("TPGAN_WGRAD_TAPPACK") ? atoi(getenv("TPGAN_WGRAD_TAPPACK")) : 1;
  const bool is_conv = (a.kind == TPGAN_CONV_FWD);
  const bool p_is_dy = is_conv && !swap;
  const tpgan_view& Pt = p_is_dy ? a.dy : a.x;
  const tpgan_view& Qt = p_is_dy ? a.x : a.dy;
  const int k = a.kh, ntaps = k * k;
  const int bf16 = a.dtype == TPGAN_DTYPE_BF16;
  const int CH = chunk_ch(bf16);        // channels per 128-byte operand chunk
  const int MCH = 128 / CH;             // P chunks per M = 128 tile
  const int NCH = 256 / CH;             // Q chunks per N = 256 MMA
  c.swap = swap;
  c.slab = slab;
  c.pc = Pt.c;
  c.qc = Qt.c;
  c.m_tiles = ceil_div(c.pc, 128);
  c.nch_total = ceil_div(c.qc, CH);
  const long long npix = (long long)Pt.h * Pt.w * Pt.n;
  if (slab) {
    static const int acc_cols = getenv("TPGAN_WGRAD_SLAB_COLS") ? atoi(getenv("TPGAN_WGRAD_SLAB_COLS")) : 512;
    const int acc_chunks = acc_cols / CH;     // CH-column accumulator blocks in TMEM
    c.cost = 1e30;
    // wide Q tensors are cut into N tiles of ncpt chunks (each re-loads P) so that a whole kernel row of taps still fits TMEM
    const int budget = g_dev.max_smem - 1024 - 256;
    for (int nt = 1; nt <= 4; nt *= 2)
      for (int share_m = 1; share_m >= 0; --share_m) {
        WgradChoice o = c;
        o.n_tiles = nt;
        o.ncpt = ceil_div(c.nch_total, nt);
        if (o.ncpt * nt != c.nch_total && nt > 1) continue;
        o.mpu = (share_m && c.m_tiles * o.ncpt * 2 <= acc_chunks) ? c.m_tiles : 1;   // (acc_chunks below may still shrink tpu)
        if (!share_m && c.m_tiles == 1) continue;
        const int tmax = std::min(std::min(k, NCH), acc_chunks / std::max(1, o.mpu * o.ncpt));   // N = tpu * CH <= 256
        if (o.ncpt > NCH || tmax < 2) continue;
        const int ngrp = ceil_div(k, tmax);
        o.tpu = ceil_div(k, ngrp);
        if (k * ngrp * o.tpu > kMaxTaps) continue;
        o.block_n = o.tpu * CH;
        o.a_ch = std::min(MCH * o.mpu, ceil_div(c.pc, CH));
        o.b_ch = o.ncpt;
        double row = 0;   // MMA cycles per K step for the taps of one kernel row, one M tile, one Q chunk
        for (int g = 0, left = k; g < ngrp; ++g) {
          const int ntp = std::min(o.tpu, left);
          left -= ntp;
          const int nn = ntp * CH;   // N of the slab MMA
          const double mma = std::max(nn / 2.0, 32.0 + nn / 4.0) * o.mpu * o.ncpt;
          const double l2 = (o.a_ch + o.ncpt) * 1024.0 / kL2BytesPerClk;
          row += std::max(mma, l2) * ceil_div(c.m_tiles, o.mpu);
        }
        // K rows spent per real pixel: narrow maps are boxed as whole rows with the padded pitch W + tpu - 1 (plan_wgrad),
        // and every box is rounded up to the MMA K step
        int px = 16;
        for (int cand : {128, 64, 32}) {
          const int sb = (o.a_ch * cand + o.b_ch * ceil_div(cand + o.tpu - 1, 8) * 8) * 128;
          if (budget / sb >= (cand == 32 ? 3 : 4)) { px = cand; break; }
        }
        const int KR = mma_k(bf16);
        const int pbw = Pt.w + o.tpu - 1;
        double keff;
        if (Pt.w <= px + px / 2 && px / pbw >= 2) {
          const int bh = std::min(Pt.h, px / pbw);
          const int bn = (bh == Pt.h) ? std::max(1, std::min(Pt.n, px / (pbw * bh))) : 1;
          const int kp = ceil_div(pbw * bh * bn, KR) * KR;
          keff = (double)Pt.h * Pt.w * bn / ((double)ceil_div(Pt.h, bh) * kp);
        } else {
          const int bw = (Pt.w <= px + px / 2) ? Pt.w : px;
          keff = (double)Pt.w / ((double)ceil_div(Pt.w, bw) * ceil_div(bw, KR) * KR);
        }
        o.cost = row * k * nt / std::max(keff, 0.05);
        if (o.cost < c.cost) { const bool sw = c.swap; c = o; c.swap = sw; c.slab = true; }
      }
    return c;
  }
  if (c.nch_total <= tap_pack_max && ntaps > 1 && !no_pack) {
    // narrow shifted tensor (one 32-channel chunk: the 3-channel image layers): several taps side by side in N, all
    // sharing the loads of P
    c.ncpt = c.nch_total;
    c.tpu = std::min(ntaps, NCH / c.nch_total);
    c.n_tiles = 1;
    c.block_n = c.tpu * c.ncpt * CH;
    c.mpu = 1;
    c.a_ch = std::min(MCH, ceil_div(c.pc, CH));
    c.b_ch = c.tpu * c.ncpt;
  } else {
    c.tpu = 1;
    c.n_tiles = ceil_div(c.nch_total, NCH);
    c.block_n = ceil_div(c.nch_total, c.n_tiles) * CH;
    c.ncpt = c.block_n / CH;
    // several M tiles per unit (one accumulator each), all sharing the loads of Q - when the reduction is long enough
    // (>= 32768 pixels) to amortise the then single-buffered epilogue
    static const long long mpu_min_pix = getenv("TPGAN_WGRAD_MPU_MINPIX") ? atoll(getenv("TPGAN_WGRAD_MPU_MINPIX")) : 32768;
    c.mpu = (no_pack || npix < mpu_min_pix) ? 1 : std::max(1, std::min(c.m_tiles, 512 / c.block_n));
    c.a_ch = std::min(MCH * c.mpu, ceil_div(c.pc, CH));
    c.b_ch = c.ncpt;
  }
  const double mma = std::max(c.block_n / 2.0, 32.0 + c.block_n / 4.0) * c.mpu;
  const double l2 = (c.a_ch + c.b_ch) * 1024.0 / kL2BytesPerClk;
  c.cost = std::max(mma, l2) * ceil_div(ntaps, c.tpu) * ceil_div(c.m_tiles, c.mpu) * c.n_tiles;
  return c;
}

// im2col-by-TMA weight gradients: Conv2d with at most 4 input channels stored as one aligned float4 per pixel (the RGB
// layers: generator conv0.0, local conv0.0, critic layer 0), stride 1 or 2, tf32.  Instead of k*k taps that each use 3 of
// the 32 K lanes of a chunk, the k horizontal taps x 4 channels of a kernel row ARE the chunk.
static thread_local int g_wgrad_ngroups = 1;   // set by tpgan_conv2d_wgrad around planning
static bool wgrad_im2col_ok(const tpgan_wgrad_args& a) {
  static const bool off = getenv("TPGAN_NO_IM2COL") != nullptr;
  if (off || !g_dev.im2col_ws) return false;
  // single launches only: in the grouped local-pathway launch (four 40x40-ish patches) the four padded copies cost more
  // than the better lane use saves (measured 0.045 -> 0.056 ms)
  if (g_wgrad_ngroups != 1) return false;
  if (a.kind != TPGAN_CONV_FWD || a.dtype != TPGAN_DTYPE_TF32) return false;
  if (a.x.c > 4) return false;
  if (a.kh != a.kw || a.kh < 2 || a.kh > 8 || (a.stride != 1 && a.stride != 2)) return false;
  const size_t bytes = (size_t)a.x.n * (a.x.h + 2 * a.pad) * (a.x.w + 2 * a.pad + 8) * 16;
  return bytes <= kIm2colSlotBytes;
}

static WgradChoice choose_wgrad(const tpgan_wgrad_args& a) {
  static const int slab_mode = getenv("TPGAN_WGRAD_SLAB") ? atoi(getenv("TPGAN_WGRAD_SLAB")) : 1;
  if (wgrad_im2col_ok(a)) {
    WgradChoice c{};
    c.im2col = true;
    c.pc = a.dy.c;
    c.qc = 32;
    c.m_tiles = ceil_div(c.pc, 128);
    c.nch_total = c.ncpt = 1;
    c.n_tiles = 1;
    c.tpu = a.kh;
    c.block_n = c.tpu * 32;
    c.mpu = 1;
    c.a_ch = std::min(4, ceil_div(c.pc, 32));
    c.b_ch = c.tpu;
    c.cost = 0;
    return c;
  }
  WgradChoice best = wgrad_option(a, false, false);
  const bool is_conv = (a.kind == TPGAN_CONV_FWD);
  const tpgan_view& Pt = is_conv ? a.dy : a.x;
  // slab mode needs unit-stride taps along W and row-segment boxes worth a pipeline stage
  static const int slab_min_w = getenv("TPGAN_WGRAD_SLAB_MINW") ? atoi(getenv("TPGAN_WGRAD_SLAB_MINW")) : 32;   // narrower maps: slab units lose to plain multi-row boxes (measured)
  if (slab_mode && a.stride == 1 && a.kh >= 2 && Pt.w >= slab_min_w) {
    for (int sw = 0; sw < (is_conv ? 2 : 1); ++sw) {
      WgradChoice c = wgrad_option(a, sw != 0, true);
      if (c.cost < best.cost * 0.95) best = c;
    }
  }
  return best;
}

// xp[n][y + p][x + p][0..3] = x[n][y][x][0..C) (C <= 4; x may be a channel slice of a wider buffer), zero elsewhere; xp is
// dense [N][Hq][Wq] with one float4 per pixel.
__global__ void pad_copy4_kernel(const float* __restrict__ x, long long sn, long long sh, long long sw, int C, int N, int H,
                                 int W, int p, int Hq, int Wq, float4* __restrict__ xp) {
  const long long total = (long long)N * Hq * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % Wq);
    const long long r = i / Wq;
    const int yq = (int)(r % Hq), n = (int)(r / Hq);
    const int xs = xq - p, ys = yq - p;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (xs >= 0 && xs < W && ys >= 0 && ys < H) {
      const float* src = x + n * sn + ys * sh + xs * sw;
      for (int c = 0; c < C; ++c) v[c] = src[c];
    }
    xp[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

static int plan_wgrad_im2col(const tpgan_wgrad_args& a, WgradGroup& G, int box_px, cudaStream_t st) {
  const int k = a.kh, s = a.stride, p = a.pad;
  const int Hq = a.x.h + 2 * p, Wq = a.x.w + 2 * p + 8;
  const unsigned slot = g_im2col_seq.fetch_add(1, std::memory_order_relaxed) % kIm2colSlots;
  float* xp = reinterpret_cast<float*>(reinterpret_cast<char*>(g_dev.im2col_ws) + (size_t)slot * kIm2colSlotBytes);
  {
    const long long total = (long long)a.x.n * Hq * Wq;
    const int grid = (int)std::max(1ll, std::min((total + 255) / 256, (long long)g_dev.sm_count * 8));
    pad_copy4_kernel<<<grid, 256, 0, st>>>(a.x.ptr, a.x.sn, a.x.sh, a.x.sw, a.x.c, a.x.n, a.x.h, a.x.w, p, Hq, Wq,
                                           reinterpret_cast<float4*>(xp));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "pad_copy4 launch: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  const tpgan_view& Pt = a.dy;
  G.transpose_out = 0;
  G.m_valid = Pt.c;
  G.n_valid = 32;
  G.im2col_k = k;
  G.im2col_cin = a.x.c;
  G.Hp = Pt.h; G.Wp = Pt.w; G.Nimg = Pt.n;
  G.bw = (G.Wp <= box_px + box_px / 2) ? G.Wp : box_px;
  G.bh = std::max(1, std::min(G.Hp, box_px / G.bw));
  G.bn = (G.bh == G.Hp && G.bw == G.Wp) ? std::max(1, std::min(G.Nimg, box_px / (G.bw * G.bh))) : 1;
  G.p_rows = G.bw * G.bh * G.bn;
  G.kp = ceil_div(G.p_rows, 8) * 8;
  G.tiles_w = ceil_div(G.Wp, G.bw);
  G.tiles_h = ceil_div(G.Hp, G.bh);
  G.chunks = G.tiles_w * G.tiles_h * ceil_div(G.Nimg, G.bn);
  G.m_tiles = ceil_div(Pt.c, 128);
  G.ncpt = 1; G.tpu = k; G.n_tiles = 1; G.block_n = k * 32; G.mpu = 1; G.slab = 0;
  G.mt_groups = G.m_tiles;
  G.ntaps = k; G.tap_groups = 1;
  G.nbuf = (G.block_n <= 256) ? 2 : 1;
  G.q_chunk_bytes = G.kp * 128;
  G.q_rows = G.p_rows;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  int rc = encode_nhwc(&G.pmap, Pt.ptr, Pt.c, Pt.w, Pt.h, Pt.n, Pt.sw, Pt.sh, Pt.sn, G.bw, G.bh, G.bn, swz, 32, 0);
  if (rc) return rc;
  if (G.kp > G.p_rows) {
    rc = encode_nhwc(&G.pzero, Pt.ptr, Pt.c, Pt.w, Pt.h, Pt.n, Pt.sw, Pt.sh, Pt.sn, G.kp - G.p_rows, 1, 1, swz, 32, 0);
    if (rc) return rc;
  }
  for (int r = 0; r < k; ++r) {
    // vertical tap r: 32 "channels" = the 8-pixel window starting at padded pixel (stride*y + r, stride*x); the pixel
    // stride of the map (16 B x stride) is smaller than its inner extent (128 B) - overlapping windows, by design
    rc = encode_nhwc(&G.qmap[r], xp + (size_t)r * Wq * 4, 32, Pt.w, Pt.h, Pt.n, 4ll * s, 4ll * s * Wq, 4ll * Hq * Wq, G.bw, G.bh, G.bn,
                     swz, 32, 0);
    if (rc) return rc;
    TapDesc t;
    t.plane = (int8_t)r;
    t.dy = t.dx = 0;
    t.wtap = (uint8_t)r;
    G.taps[r] = t;
  }
  return 0;
}

static int plan_wgrad(const tpgan_wgrad_args& a, WgradGroup& G, int box_px, cudaStream_t st = nullptr) {
  memset(&G, 0, sizeof(G));
  const int k = a.kh;
  if (a.kh != a.kw || k < 1 || k > 8) return set_error(TPGAN_ERR_INVALID, "kernel %dx%d unsupported", a.kh, a.kw);
  const int s = a.stride, p = a.pad;
  if (s != 1 && s != 2 && s != 4) return set_error(TPGAN_ERR_INVALID, "stride %d unsupported", s);
  if (a.x.n != a.dy.n) return set_error(TPGAN_ERR_INVALID, "batch mismatch");
  // conv  : P = dy (Cout), Q = x planes  (Cin)  -> dw[tap][co][ci], m=co n=ci
  // conv, swapped (stride 1): P = x (Cin), Q = dy shifted by -tap (Cout) -> m=ci n=co (transposed write)
  // deconv: P = x  (Cin),  Q = dy planes (Cout) -> dw[tap][co][ci], m=ci n=co (transposed write)
  const bool is_conv = (a.kind == TPGAN_CONV_FWD);
  if (!is_conv && a.kind != TPGAN_DECONV_FWD) return set_error(TPGAN_ERR_INVALID, "bad wgrad kind %d", a.kind);
  if (a.dtype != TPGAN_DTYPE_TF32 && a.dtype != TPGAN_DTYPE_BF16) return set_error(TPGAN_ERR_INVALID, "bad dtype %d", a.dtype);
  const int bf16 = a.dtype == TPGAN_DTYPE_BF16;
  const int KR = mma_k(bf16);           // pixel rows per MMA K step
  const WgradChoice ch = choose_wgrad(a);
  const bool p_is_dy = is_conv && !ch.swap;
  const tpgan_view& Pt = p_is_dy ? a.dy : a.x;
  const tpgan_view& Qt = p_is_dy ? a.x : a.dy;
  G.transpose_out = p_is_dy ? 0 : 1;
  G.m_valid = Pt.c;
  G.n_valid = Qt.c;
  const int cout = a.dy.c, cin = a.x.c;
  if (a.w_rows_pad < cout || a.w_k_pad < cin) return set_error(TPGAN_ERR_INVALID, "packed dw too small");
  G.rows_pad = a.w_rows_pad;
  G.k_pad = a.w_k_pad;
  G.dw = a.dw_packed;
  G.accumulate = a.accumulate;
  if (ch.im2col) return plan_wgrad_im2col(a, G, box_px, st);
  G.Hp = Pt.h;
  G.Wp = Pt.w;
  G.Nimg = Pt.n;
  // pixel boxes of ~box_px pixels.  Slab mode: one row segment - or, on maps narrow enough for two or more whole rows per
  // box, rows enumerated with the PADDED pitch W + tpu - 1 (the pad columns are out-of-bounds zeros of P): tap t of every
  // pixel is still the Q slab read t rows further down, and a stage carries ~box_px pixels instead of one short row
  G.bw = (G.Wp <= box_px + box_px / 2) ? G.Wp : box_px;
  int pbw = G.bw;                       // width of the P (and, multi-row, Q) box
  G.bh = ch.slab ? 1 : std::max(1, std::min(G.Hp, box_px / G.bw));
  if (ch.slab && G.bw == G.Wp && box_px / (G.Wp + ch.tpu - 1) >= 2) {
    pbw = G.Wp + ch.tpu - 1;
    G.bh = std::min(G.Hp, box_px / pbw);
  }
  const bool multirow = ch.slab && pbw != G.bw;
  G.bn = ((!ch.slab || multirow) && G.bh == G.Hp && G.bw == G.Wp) ? std::max(1, std::min(G.Nimg, box_px / (pbw * G.bh))) : 1;
  G.p_rows = pbw * G.bh * G.bn;
  G.kp = ceil_div(G.p_rows, KR) * KR;
  G.tiles_w = ceil_div(G.Wp, G.bw);
  G.tiles_h = ceil_div(G.Hp, G.bh);
  G.chunks = G.tiles_w * G.tiles_h * ceil_div(G.Nimg, G.bn);
  G.m_tiles = ch.m_tiles;
  G.ncpt = ch.ncpt;
  G.tpu = ch.tpu;
  G.n_tiles = ch.n_tiles;
  G.block_n = ch.block_n;
  G.mpu = ch.mpu;
  G.slab = ch.slab ? 1 : 0;
  G.mt_groups = ceil_div(G.m_tiles, G.mpu);
  // MN-major tf32 operands need the 32B-atom flavour of the 128B swizzle; bf16 uses the plain 128B swizzle
  const CUtensorMapSwizzle swz = bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const int CH = chunk_ch(bf16);
  int rc = encode_nhwc(&G.pmap, Pt.ptr, Pt.c, Pt.w, Pt.h, Pt.n, Pt.sw, Pt.sh, Pt.sn, pbw, G.bh, G.bn, swz, CH, bf16);
  if (rc) return rc;
  if (G.kp > G.p_rows) {
    rc = encode_nhwc(&G.pzero, Pt.ptr, Pt.c, Pt.w, Pt.h, Pt.n, Pt.sw, Pt.sh, Pt.sn, G.kp - G.p_rows, 1, 1, swz, CH, bf16);
    if (rc) return rc;
  }
  if (ch.slab) {
    // taps: per kernel row, groups of tpu slots (unused slots: wtap 255), horizontal offset increasing inside a group
    const int ngrp = ceil_div(k, G.tpu);
    int nt = 0;
    for (int r = 0; r < k; ++r)
      for (int g = 0; g < ngrp; ++g)
        for (int t = 0; t < G.tpu; ++t) {
          const int j = g * G.tpu + t;          // position along the row in order of increasing offset
          TapDesc td;
          td.plane = 0;
          if (j < k) {
            const int c = ch.swap ? (k - 1 - j) : j;   // swapped: the shifted tensor is dy at offset -(tap - p)
            const int ey = r - p, ex = c - p;
            td.dy = (int8_t)(ch.swap ? -ey : ey);
            td.dx = (int8_t)(ch.swap ? -ex : ex);
            td.wtap = (uint8_t)(r * k + c);
          } else {
            td.dy = td.dx = 0;
            td.wtap = 255;
          }
          G.taps[nt++] = td;
        }
    G.ntaps = nt;
    G.tap_groups = k * ngrp;
    G.nbuf = (G.mpu * G.ncpt * G.tpu * CH <= 256) ? 2 : 1;
    G.q_chunk_bytes = ceil_div(G.kp + G.tpu - 1, 8) * 8 * 128;
    // multi-row: the Q box has the P box's shape (rows past it, read by the last taps of the pad columns, stay zero: the
    // ring is cleared once per launch); single row: one segment + halo
    G.q_rows = multirow ? G.p_rows : G.bw + G.tpu - 1;
    G.zero_ring = multirow ? 1 : 0;
    rc = encode_nhwc(&G.qslab, Qt.ptr, Qt.c, Qt.w, Qt.h, Qt.n, Qt.sw, Qt.sh, Qt.sn, multirow ? pbw : G.bw + G.tpu - 1,
                     multirow ? G.bh : 1, multirow ? G.bn : 1, swz, CH, bf16);
    if (rc) return rc;
    return 0;
  }
  G.ntaps = k * k;
  G.tap_groups = ceil_div(G.ntaps, G.tpu);
  G.nbuf = (G.mpu * G.block_n <= 256) ? 2 : 1;
  G.q_chunk_bytes = G.kp * 128;
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
  rc = encode_planes(G.qmap, Qt, s, G.bw, G.bh, G.bn, swz, bf16);
  if (rc) return rc;
  return 0;
}

// Pixels per pipeline stage: the largest of {128, 64, 32} that still leaves >= 4 stages in shared memory for every group
// (few, large TMA boxes and 8-16 MMAs per barrier round trip; measured 1.8x on the 64-channel 128x128 layers).
static int choose_wgrad_px(const tpgan_wgrad_args* groups, int ngroups) {
  if (const char* ev = getenv("TPGAN_WGRAD_PX")) return std::max(8, std::min(128, atoi(ev)));
  const int budget = g_dev.max_smem - 1024 - 256;
  auto stage_bytes = [&](int px) {
    int worst = 0;
    for (int i = 0; i < ngroups; ++i) {
      const WgradChoice c = choose_wgrad(groups[i]);
      const int q_rows = c.slab ? ceil_div(px + c.tpu - 1, 8) * 8 : px;
      worst = std::max(worst, c.a_ch * px * 128 + c.b_ch * q_rows * 128);   // bytes: a chunk row is 128 B for either dtype
    }
    return worst;
  };
  for (int px : {128, 64})
    if (budget / stage_bytes(px) >= 4) return px;
  // small maps (multi-row boxes): three stages of 64 pixels beat more stages of 32 - the per-stage cost (a dozen TMA
  // requests of 2-4 KB and a barrier round trip per 3-4 K steps) dominates those launches (measured on the grouped
  // 128->128 3x3 layer at 20x20: 81 -> 50 us)
  bool small_maps = true;
  for (int i = 0; i < ngroups; ++i) {
    const bool is_conv = groups[i].kind == TPGAN_CONV_FWD;
    small_maps = small_maps && (is_conv ? groups[i].dy.w : groups[i].x.w) <= 64;
  }
  if (small_maps && budget / stage_bytes(64) >= 3) return 64;
  if (budget / stage_bytes(32) >= 3) return 32;
  return 16;
}

template <class Params>
static int launch_wgrad(Params& P, cudaStream_t st, int bf16) {
  int amax = 0, bmax = 0;
  long long work = 0;
  const int CH = chunk_ch(bf16), MCH = 128 / CH, KR = mma_k(bf16);
  for (int i = 0; i < P.ngroups; ++i) {
    WgradGroup& G = P.g[i];
    amax = std::max(amax, std::min(MCH * G.mpu, ceil_div(G.m_valid, CH)) * G.kp * 128);
    bmax = std::max(bmax, (G.slab ? G.ncpt : G.block_n / CH) * G.q_chunk_bytes);
    G.tiles = G.tap_groups * G.mt_groups * G.n_tiles;
    // K blocks: pixel ranges whose activations (P and Q) fit a slice of L2; every CTA works through block after block
    {
      static const double mb_plain = getenv("TPGAN_WGRAD_KB_MB") ? atof(getenv("TPGAN_WGRAD_KB_MB")) : 64.0;
      // slab units stream little from L2 and pay for every extra epilogue: one K block - unless the activations would be
      // streamed from HBM a dozen times (75->75 k7: 7 kernel rows x 2 tap groups = 14 passes over 0.32 GB = 3.8 GB of DRAM
      // reads at 51 % of the DRAM peak, ncu_full_r2_wgrad_add128): then 64 MB blocks keep a block's pixels in L2 for all
      // its tiles (measured 0.910 -> 0.845 ms; the 64-channel and swapped 206->64 layers, 7-10 passes, lose with blocks)
      static const double mb_slab_env = getenv("TPGAN_WGRAD_KB_MB_SLAB") ? atof(getenv("TPGAN_WGRAD_KB_MB_SLAB")) : 0.0;
      const double mb_slab = mb_slab_env > 0 ? mb_slab_env : (G.tiles >= 12 ? 64.0 : 1e6);
      const double bytes = (double)G.Hp * G.Wp * G.Nimg * (bf16 ? 2.0 : 4.0) * (ceil_div(G.m_valid, 4) * 4 + ceil_div(G.n_valid, 4) * 4);
      int nkb = (int)std::ceil(bytes / ((G.slab ? mb_slab : mb_plain) * 1048576.0));
      nkb = std::max(1, std::min(nkb, G.chunks));
      if (g_deterministic.load(std::memory_order_relaxed)) nkb = 1;
      G.kb_chunks = ceil_div(G.chunks, nkb);
      G.nkb = ceil_div(G.chunks, G.kb_chunks);
    }
    G.work_begin = (int)work;
    work += (long long)G.tiles * G.chunks;
  }
  if (work <= 0 || work > 0x7fffffffll) return set_error(TPGAN_ERR_INVALID, "wgrad: work list of %lld chunks unsupported", work);
  P.total_work = (int)work;
  P.whole_tiles = g_deterministic.load(std::memory_order_relaxed) ? 1 : 0;
  P.nbuf = 2;
  for (int i = 0; i < P.ngroups; ++i) P.nbuf = std::min(P.nbuf, P.g[i].nbuf);
  P.a_stage_bytes = amax;
  P.b_stage_bytes = bmax;
  const int stage_bytes = amax + bmax;
  // an M = 128 MMA always reads four 32-channel P chunks; when fewer are loaded (P narrower than 128 channels) the read
  // runs into the following bytes (rows that are never stored), so the ring is followed by that much slack
  int slack = 0;
  for (int i = 0; i < P.ngroups; ++i) slack = std::max(slack, MCH * P.g[i].mpu * P.g[i].kp * 128 - stage_bytes);
  slack = std::max(0, slack);
  const int budget = g_dev.max_smem - 1024 - 256 - slack;
  P.stages = std::min(kMaxStages, budget / stage_bytes);
  P.ring_bytes = P.stages * stage_bytes + slack;
  P.need_zero = 0;
  for (int i = 0; i < P.ngroups; ++i)
    if ((P.g[i].p_rows % KR) || P.g[i].zero_ring) P.need_zero = 1;
  if (P.stages < 2) return set_error(TPGAN_ERR_INVALID, "wgrad: not enough shared memory for 2 stages (%d B/stage)", stage_bytes);
  const int smem = P.ring_bytes + 1024;
  static const bool debug = getenv("TPGAN_WGRAD_DEBUG") != nullptr;
  if (debug)
    for (int i = 0; i < P.ngroups; ++i) {
      const WgradGroup& G = P.g[i];
      fprintf(stderr, "wgrad[%d/%d] P %dx%dx%d m_valid %d n_valid %d slab %d transpose %d mpu %d tpu %d ncpt %d n_tiles %d block_n %d box %dx%dx%d p_rows %d kp %d "
              "chunks %d tiles %d nkb %d stages %d stage_bytes %d nbuf %d zero_tail %d need_zero %d\n", i, P.ngroups, G.Nimg, G.Hp, G.Wp, G.m_valid,
              G.n_valid, G.slab, G.transpose_out, G.mpu, G.tpu, G.ncpt, G.n_tiles, G.block_n, G.bw, G.bh, G.bn, G.p_rows, G.kp, G.chunks, G.tiles,
              G.nkb, P.stages, stage_bytes, P.nbuf, P.zero_tail, P.need_zero);
    }
  auto kern = bf16 ? wgrad_kernel<Params, true> : wgrad_kernel<Params, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dev.max_smem - 256);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  // balanced schedule: every CTA gets total_work / grid (+-1) chunks (see SegmentWalk in wgrad.cu)
  const int grid = (int)std::min<long long>(work, persistent_sms());
  e = launch_tc(kern, grid, smem, st, P);
  if (e != cudaSuccess) return set_error(TPGAN_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace tpg

using namespace tpg;

static thread_local int g_last_conv_kernel = 0;

extern "C" {

int tpgan_conv2d(const tpgan_conv_args* groups, int32_t ngroups, void* stream) {
  if (!groups || ngroups < 1 || ngroups > kMaxGroups) return set_error(TPGAN_ERR_INVALID, "ngroups must be 1..%d", kMaxGroups);
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int bf16 = groups[0].dtype == TPGAN_DTYPE_BF16;
  for (int i = 1; i < ngroups; ++i)
    if (groups[i].dtype != groups[0].dtype) return set_error(TPGAN_ERR_INVALID, "grouped problems must share one dtype");
  if (ngroups == 1) {
    int rrc = 0;
    if (!groups[0].in_lo.ptr) {      // the three-term product is a tap-list feature of tapgemm
      g_last_conv_kernel = 2;
      if (try_rowstack(groups[0], st, &rrc)) return rrc;
      g_last_conv_kernel = 1;
      if (try_rowconv(groups[0], st, &rrc)) return rrc;
      g_last_conv_kernel = 3;
      if (try_flatconv(groups, 1, st, &rrc)) return rrc;
    }
    g_last_conv_kernel = 0;
    static thread_local TapGemmParams1 P;
    P.ngroups = 1;
    t_pair = false;
    rc = plan_group(groups[0], P.g[0]);
    if (rc) return rc;
    const int nt = choose_n_tiles(P.g[0], groups[0]);
    if (nt > 0) {
      rc = plan_group(groups[0], P.g[0], -nt);
      if (rc) return rc;
    }
    // Split-K for GEMM-like launches (the Linear layers: rows = images) whose few tiles stream a long reduction - fc1, an
    // 8x8 "conv" of 64 taps on the 8x8x512 map, reads 67 MB of weights on 16 CTAs otherwise.  The reduction is cut into
    // ranges of whole taps (or of K chunks of a single tap); partial accumulators go to a workspace slot and the CTA
    // arriving last at the tile's counter adds them in range order and runs the normal epilogue: deterministic, any epilogue.
    {
      const tpgan_conv_args& a = groups[0];
      TapGemmGroup& G = P.g[0];
      static const bool no_splitk = getenv("TPGAN_NO_SPLITK") != nullptr;
      if (!no_splitk && g_dev.split_ws && a.stride == 1 && a.pad == 0 && a.kind == TPGAN_CONV_FWD && a.out.h == 1 && a.out.w == 1 &&
          G.n_phases == 1 && G.tile_count * 4 <= persistent_sms() && G.tile_count <= kSplitCounters) {
        const int ntap = G.phase[0].tap_count;
        const int want = persistent_sms() / G.tile_count;
        int ksplit = 1, kt_per = 0, kc_per = G.kchunks;
        if ((long long)ntap * G.kchunks >= 256) {
          if (ntap >= 2 * want) {
            kt_per = ceil_div(ntap, want);
            ksplit = ceil_div(ntap, kt_per);
          } else if (ntap == 1 && G.kchunks >= 64) {
            const int ks = std::min(want, G.kchunks / 16);
            kc_per = ceil_div(ceil_div(G.kchunks, ks), 2) * 2;   // even: a stage carries up to two chunks
            ksplit = ceil_div(G.kchunks, kc_per);
          }
        }
        if (ksplit > 1 && (size_t)G.tile_count * ksplit * 128 * G.block_n * sizeof(float) <= kSplitSlotBytes) {
          const unsigned slot = g_split_seq.fetch_add(1, std::memory_order_relaxed) % kSplitSlots;
          G.split_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(g_dev.split_ws) + (size_t)slot * kSplitSlotBytes);
          G.split_cnt = g_dev.split_cnt + slot * kSplitCounters;
          G.ksplit = ksplit; G.kt_per = kt_per; G.kc_per = kc_per;
          G.tile_count *= ksplit;
        }
      }
    }
    if (pair_mode() && P.g[0].ksplit == 1 && P.g[0].m_tiles >= 2) {   // same tiling, over CTA pairs
      t_pair = true;
      rc = plan_group(groups[0], P.g[0], -P.g[0].n_tiles);
      if (!rc) rc = launch_tapgemm(P, st, bf16);
      t_pair = false;
      return rc;
    }
    return launch_tapgemm(P, st, bf16);
  }
  {
    int rrc = 0;
    bool any_lo = false;
    for (int i = 0; i < ngroups; ++i) any_lo = any_lo || groups[i].in_lo.ptr != nullptr;
    g_last_conv_kernel = 3;
    if (!any_lo && try_flatconv(groups, ngroups, st, &rrc)) return rrc;
  }
  g_last_conv_kernel = 0;
  static thread_local TapGemmParams P;
  P.ngroups = ngroups;
  t_pair = false;
  for (int i = 0; i < ngroups; ++i) {
    rc = plan_group(groups[i], P.g[i]);
    if (rc) return rc;
  }
  const int f = choose_n_split(P, groups);
  bool pair = pair_mode() != 0;
  for (int i = 0; i < ngroups; ++i) pair = pair && P.g[i].m_tiles >= 2;
  if (f > 1 || pair) {
    t_pair = pair;
    for (int i = 0; i < ngroups && !rc; ++i) rc = plan_group(groups[i], P.g[i], f);
  }
  if (!rc) rc = launch_tapgemm(P, st, bf16);
  t_pair = false;
  return rc;
}

int tpgan_conv2d_wgrad(const tpgan_wgrad_args* groups, int32_t ngroups, void* stream) {
  if (!groups || ngroups < 1 || ngroups > kMaxGroups) return set_error(TPGAN_ERR_INVALID, "ngroups must be 1..%d", kMaxGroups);
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int bf16 = groups[0].dtype == TPGAN_DTYPE_BF16;
  for (int i = 1; i < ngroups; ++i)
    if (groups[i].dtype != groups[0].dtype) return set_error(TPGAN_ERR_INVALID, "grouped problems must share one dtype");
  g_wgrad_ngroups = ngroups;
  if (ngroups == 1) {
    static thread_local WgradParams1 P;
    P.ngroups = 1;
    P.zero_tail = 0;
    rc = plan_wgrad(groups[0], P.g[0], choose_wgrad_px(groups, 1), st);
    if (rc) return rc;
    return launch_wgrad(P, st, bf16);
  }
  static thread_local WgradParams P;
  P.ngroups = ngroups;
  bool uniform = true;
  const int px = choose_wgrad_px(groups, ngroups);
  for (int i = 0; i < ngroups; ++i) {
    rc = plan_wgrad(groups[i], P.g[i], px, st);
    if (rc) return rc;
    const WgradGroup& G = P.g[i];
    if (G.p_rows % mma_k(bf16)) uniform = false;
  }
  // the zero-padded K rows of the smem ring stay zero by themselves only if every box fills its rows or all boxes are
  // alike; otherwise each stage re-zeroes them (WgradGroup::pzero)
  P.zero_tail = 0;
  if (!uniform) {
    bool same = true;
    for (int i = 1; i < ngroups; ++i)
      same = same && P.g[i].bw == P.g[0].bw && P.g[i].bh == P.g[0].bh && P.g[i].bn == P.g[0].bn && P.g[i].p_rows == P.g[0].p_rows &&
             P.g[i].block_n == P.g[0].block_n && P.g[i].mpu == P.g[0].mpu && P.g[i].tpu == P.g[0].tpu &&
             P.g[i].slab == P.g[0].slab;
    if (!same) P.zero_tail = 1;
  }
  return launch_wgrad(P, st, bf16);
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
int tpgan_set_deterministic(int32_t on) {
  const int prev = g_deterministic.exchange(on ? 1 : 0);
  return prev;
}
int tpgan_get_deterministic(void) { return g_deterministic.load(); }
int tpgan_set_sm_reserve(int32_t sms) {
  const int prev = g_sm_reserve.exchange(std::max(0, std::min((int)sms, 64)));
  return prev;
}
int tpgan_last_conv_kernel(void) { return g_last_conv_kernel; }
int tpgan_last_conv_pair(void) { return g_last_conv_pair; }

}  // extern "C"
