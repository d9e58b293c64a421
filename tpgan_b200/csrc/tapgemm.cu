// Multi-tap implicit-GEMM convolution on Blackwell tensor cores (sm_100a).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer   - per (tap, 32-channel chunk) loads one activation box {32ch, bw, bh, bn} straight
//                               from the NHWC tensor (zero padding = TMA out-of-bounds fill, stride-2/4 sampling =
//                               parity-plane tensor maps) and one weight box {32, block_n} into a 128B-swizzled ring
//   warp 1   : MMA issuer     - tcgen05.mma kind::tf32, M=128 x N=block_n x K=8 per instruction, fp32 accumulator in
//                               TMEM (two 256-column accumulators so the epilogue overlaps the next tile)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue       - tcgen05.ld -> bias / addends / (Leaky)ReLU or activation-backward mask / tf32 round
//                               -> NHWC global stores (optionally strided x2/x4 for transposed convs, optionally into a
//                               channel slice of a wider concat buffer)
//
// Serves: Conv2d fwd + dgrad, ConvTranspose2d fwd + dgrad (ModificationLayer.py:101,189 of the reference), and the
// Linear layers (as 1x1 problems).  Up to four independent problems share one launch (grouped local pathways).
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kAStageBytes = 128 * 128;  // 128 rows x 32 fp32
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;

struct TileCoord {
  int gi, ph, h0, n0, nt;
};

template <class Params>
__device__ __forceinline__ TileCoord decode_tile(const Params& P, int tile) {
  TileCoord tc;
  int gi = 0;
  constexpr int kG = (int)(sizeof(P.g) / sizeof(P.g[0]));
#pragma unroll
  for (int i = 1; i < kG; ++i)
    if (i < P.ngroups && tile >= P.g[i].tile_begin) gi = i;
  const TapGemmGroup& G = P.g[gi];
  int local = tile - G.tile_begin;
  tc.gi = gi;
  tc.nt = local % G.n_tiles;
  int rest = local / G.n_tiles;
  int mt = rest % G.m_tiles;
  tc.ph = rest / G.m_tiles;
  int hb = mt % G.tiles_h;
  int nb = mt / G.tiles_h;
  tc.h0 = hb * G.bh;
  tc.n0 = nb * G.bn;
  return tc;
}

template <class Params>
__global__ void __launch_bounds__(256, 1) tapgemm_kernel(const __grid_constant__ Params P, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t stage_bytes = kAStageBytes + P.b_stage_bytes;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  AbortCtl ac{&abort_flag, status};

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        TileCoord tc = decode_tile(P, tile);
        const TapGemmGroup& G = P.g[tc.gi];
        const uint32_t tx = (uint32_t)(G.bw * G.bh * G.bn + G.block_n) * 128u;
        const PhaseDesc ph = G.phase[tc.ph];
        const int rot = (P.tap_rot & 1) ? (int)(blockIdx.x % (unsigned)ph.tap_count) : 0;
        for (int t0 = 0; ok && t0 < ph.tap_count; ++t0) {
          int t = t0 + rot;
          if (t >= ph.tap_count) t -= ph.tap_count;
          const TapDesc tap = G.taps[ph.tap_begin + t];
          const CUtensorMap* am = &G.amap[tap.plane];
          for (int c = 0; c < G.kchunks; ++c) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1u, ac, 1)) { ok = false; break; }
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            const int dbg = P.tap_rot >> 4;
            uint32_t txx = 0;
            if (!(dbg & 1)) txx += (uint32_t)(G.bw * G.bh * G.bn) * 128u;
            if (!(dbg & 2)) txx += (uint32_t)G.block_n * 128u;
            if (txx) mbar_arrive_expect_tx(&full_bar[stage], txx); else mbar_arrive(&full_bar[stage]);
            if (!(dbg & 1)) tma_load_4d(sa, am, &full_bar[stage], c * 32, tap.dx, tc.h0 + tap.dy, tc.n0);
            if (!(dbg & 2)) tma_load_3d(sa + kAStageBytes, &G.bmap, &full_bar[stage], c * 32, tc.nt * G.block_n, tap.wtap);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        TileCoord tc = decode_tile(P, tile);
        const TapGemmGroup& G = P.g[tc.gi];
        if (!mbar_wait(&tempty_bar[as], aphase ^ 1u, ac, 2)) break;
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(128, G.block_n, 0, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccCols);
        const int ntap = G.phase[tc.ph].tap_count;
        uint32_t acc = 0;
        for (int t = 0; ok && t < ntap; ++t) {
          for (int c = 0; c < G.kchunks; ++c) {
            if (!mbar_wait(&full_bar[stage], phase, ac, 3)) { ok = false; break; }
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t b_addr = a_addr + kAStageBytes;
            const int nm = (c == G.kchunks - 1) ? G.last_mmas : 4;
            for (int k = 0; k < nm; ++k) {
              if ((P.tap_rot >> 4) & 4) break;
              mma_tf32_ss(d_tmem, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                          idesc, acc);
              acc = 1;
            }
            tc_commit(&empty_bar[stage]);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
        if (!ok) break;
        tc_commit(&tfull_bar[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      TileCoord tc = decode_tile(P, tile);
      const TapGemmGroup& G = P.g[tc.gi];
      if (!mbar_wait(&tfull_bar[as], aphase, ac, 4)) break;
      tc_fence_after();
      // row -> pixel
      int wl = row % G.bw;
      int r2 = row / G.bw;
      int hl = r2 % G.bh;
      int nl = r2 / G.bh;
      const int y = tc.h0 + hl, n = tc.n0 + nl;
      const PhaseDesc ph = G.phase[tc.ph];
      const int yo = y * G.out_sy + ph.oy, xo = wl * G.out_sx + ph.ox;
      const bool valid = (nl < G.bn) && (y < G.Hm) && (n < G.Nimg) && (yo < G.Hout) && (xo < G.Wout);
      float* po = G.out.ptr + (long long)n * G.out.sn + (long long)yo * G.out.sh + (long long)xo * G.out.sw;
      const float* p1 = G.add1.ptr ? G.add1.ptr + (long long)n * G.add1.sn + (long long)yo * G.add1.sh + (long long)xo * G.add1.sw : nullptr;
      const float* p2 = G.add2.ptr ? G.add2.ptr + (long long)n * G.add2.sn + (long long)yo * G.add2.sh + (long long)xo * G.add2.sw : nullptr;
      const float* pm = G.mask.ptr ? G.mask.ptr + (long long)n * G.mask.sn + (long long)yo * G.mask.sh + (long long)xo * G.mask.sw : nullptr;
      const uint32_t t_addr = tmem_base + (uint32_t)(as * kAccCols) + ((uint32_t)(q * 32) << 16);
      const int col_base = tc.nt * G.block_n;
      for (int c0 = 0; c0 < G.block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const int col = col_base + c0 + j;
            const int nv = G.cout_valid - col;
            if (nv <= 0) break;
            float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3])};
            const bool vec = G.vec_ok && nv >= 4;
            if (G.bias) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < nv) v[i] += __ldg(G.bias + col + i);
            }
            if (p1) {
              if (vec) {
                float4 a = *reinterpret_cast<const float4*>(p1 + col);
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < nv) v[i] += p1[col + i];
              }
            }
            if (p2) {
              if (vec) {
                float4 a = *reinterpret_cast<const float4*>(p2 + col);
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < nv) v[i] += p2[col + i];
              }
            }
            if (G.epilogue == 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * G.slope;
            } else if (G.epilogue == 2) {
              float m[4] = {1.f, 1.f, 1.f, 1.f};
              if (vec) {
                float4 a = *reinterpret_cast<const float4*>(pm + col);
                m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < nv) m[i] = pm[col + i];
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float s = G.slope;
                if (G.slopes && i < nv) s = __ldg(G.slopes + col + i);
                v[i] = m[i] > 0.f ? v[i] : v[i] * s;
              }
            }
            if (G.round_tf32) {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = round_tf32(v[i]);
            }
            if (vec) {
              *reinterpret_cast<float4*>(po + col) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < nv) po[col + i] = v[i];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// explicit instantiations used by api.cu
template __global__ void tapgemm_kernel<TapGemmParams>(const __grid_constant__ TapGemmParams, int*);
template __global__ void tapgemm_kernel<TapGemmParams1>(const __grid_constant__ TapGemmParams1, int*);

}  // namespace tpg
