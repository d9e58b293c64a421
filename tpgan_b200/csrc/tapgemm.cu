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
//
// PAIR = true: the same kernel over CTA pairs (clusters of 2, tcgen05 cta_group::2, common.cuh).  A pair works on two
// consecutive M tiles of one (phase, N tile): every MMA is M = 256 x N = block_n, each CTA stages its own 128 activation
// rows and HALF of the weight tile (block_n / 2 rows), so per MMA a CTA's shared memory takes in and serves half the weight
// bytes - the shared-memory port is what bounds this kernel for N <= 208.
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kAStageBytes = 128 * 128;  // 128 rows x 32 fp32
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue (two warps per TMEM lane quarter, each
// taking every other 16-column group - the epilogue is latency-bound on its addend/mask loads, not on TMEM reads)
constexpr int kTapGemmThreads = kConvThreads;

struct TileCoord {
  int gi, ph, h0, n0, nt;
  int t0, t1, c0, c1;   // taps t0 .. t1 and K chunks c0 .. c1 of this tile (everything unless the group is split)
  int ks, base;         // split-K: range index, un-split tile index inside the group
};

template <class Params, bool PAIR = false>
__device__ __forceinline__ TileCoord decode_tile(const Params& P, int tile, int rank = 0) {
  TileCoord tc;
  int gi = 0;
  constexpr int kG = (int)(sizeof(P.g) / sizeof(P.g[0]));
#pragma unroll
  for (int i = 1; i < kG; ++i)
    if (i < P.ngroups && tile >= P.g[i].tile_begin) gi = i;
  const TapGemmGroup& G = P.g[gi];
  int local = tile - G.tile_begin;
  tc.gi = gi;
  const int ks = local % G.ksplit;
  local /= G.ksplit;
  tc.ks = ks;
  tc.base = local;
  tc.nt = local % G.n_tiles;
  int rest = local / G.n_tiles;
  int mt;
  if constexpr (PAIR) {   // `tile` counts pairs of M tiles (split-K launches are never paired); an odd count leaves the
    const int m_pairs = (G.m_tiles + 1) >> 1;   // last pair's second CTA a tile past the end (loads zero-fill, no stores)
    mt = 2 * (rest % m_pairs) + rank;
    tc.ph = rest / m_pairs;
  } else {
    mt = rest % G.m_tiles;
    tc.ph = rest / G.m_tiles;
  }
  int hb = mt % G.tiles_h;
  int nb = mt / G.tiles_h;
  tc.h0 = hb * G.bh;
  tc.n0 = nb * G.bn;
  const int ntap = G.phase[tc.ph].tap_count;
  tc.t0 = 0; tc.t1 = ntap; tc.c0 = 0; tc.c1 = G.kchunks;
  if (G.ksplit > 1) {
    if (G.kt_per > 0) { tc.t0 = ks * G.kt_per; tc.t1 = min(ntap, tc.t0 + G.kt_per); }
    else { tc.c0 = ks * G.kc_per; tc.c1 = min(G.kchunks, tc.c0 + G.kc_per); }
  }
  return tc;
}

template <class Params, bool BF16, bool PAIR>
__global__ void __launch_bounds__(kTapGemmThreads, 1) tapgemm_kernel(const __grid_constant__ Params P, int* status) {
  using Op = Opnd<BF16>;
  // PAIR: cluster rank (0 = leader: issues the MMAs, owns the full / accumulator-empty barriers), pair index, pair count
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int CH = Op::kChunk;   // channels per 128-byte K chunk
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTapMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kTapMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;
  __shared__ int split_last;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = P.stages;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      // pairs: one arrival per epilogue warp of BOTH CTAs, on the leader's barrier
      mbar_init(&tempty_bar[i], PAIR ? 2 * (kTapGemmThreads - 128) / 32 : kTapGemmThreads - 128);
    }
    abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc2(&tmem_base_s, kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_s, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  pdl_wait();   // everything above touched only shared / tensor memory; global memory from here on
  const uint32_t tmem_base = tmem_base_s;
  AbortCtl ac{&abort_flag, status};

  const int kst = P.kst;  // 32-float K chunks per pipeline stage (1 or 2)
  const uint32_t a_region = (uint32_t)kst * kAStageBytes;
  const uint32_t b_chunk = (uint32_t)P.b_stage_bytes;
  const uint32_t stage_bytes = a_region + (uint32_t)kst * b_chunk;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t sa = smem_base, fb = full0, eb = empty0;
      bool ok = true;
      for (int tile = tile0; ok && tile < P.total_tiles; tile += tstep) {
        const TileCoord tc = decode_tile<Params, PAIR>(P, tile, rank);
        const TapGemmGroup& G = P.g[tc.gi];
        // pairs: the leader's barrier expects the bytes of both CTAs (2 x (A box + half a weight box))
        const uint32_t chunk_tx = PAIR ? (uint32_t)(2 * G.bw * G.bh * G.bn + G.block_n) * 128u
                                       : (uint32_t)(G.bw * G.bh * G.bn + G.block_n) * 128u;
        const PhaseDesc ph = G.phase[tc.ph];
        const int kchunks = G.kchunks;
        const int bn0 = tc.nt * G.block_n + (PAIR ? rank * (G.block_n >> 1) : 0);
        const TapDesc* taps = &G.taps[ph.tap_begin];
        for (int t = tc.t0; ok && t < tc.t1; ++t) {
          const TapDesc tap = taps[t];
          const CUtensorMap* am = &G.amap[tap.plane];
          const CUtensorMap* bm = (tap.wtap & 0x80) ? &G.bmap_lo : &G.bmap;   // bit 7: the residual weight packing
          const int ax = tap.dx, ay = tc.h0 + tap.dy, wt = tap.wtap & 0x7f;
          for (int c = tc.c0; c < tc.c1; c += kst) {
            if (!mbar_wait_a(eb, phase ^ 1u, ac, 1)) { ok = false; break; }
            const uint32_t sb = sa + a_region;
            const bool two = (kst > 1) && (c + 1 < tc.c1);
            if constexpr (PAIR) {
              if (rank == 0) mbar_arrive_expect_tx_a(fb, two ? 2u * chunk_tx : chunk_tx);
              tma_load_4d_pair(sa, am, fb, c * CH, ax, ay, tc.n0);
              tma_load_3d_pair(sb, bm, fb, c * CH, bn0, wt);
              if (two) {
                tma_load_4d_pair(sa + kAStageBytes, am, fb, c * CH + CH, ax, ay, tc.n0);
                tma_load_3d_pair(sb + b_chunk, bm, fb, c * CH + CH, bn0, wt);
              }
            } else {
              mbar_arrive_expect_tx_a(fb, two ? 2u * chunk_tx : chunk_tx);
              tma_load_4d_a(sa, am, fb, c * CH, ax, ay, tc.n0);
              tma_load_3d_a(sb, bm, fb, c * CH, bn0, wt);
              if (two) {
                tma_load_4d_a(sa + kAStageBytes, am, fb, c * CH + CH, ax, ay, tc.n0);
                tma_load_3d_a(sb + b_chunk, bm, fb, c * CH + CH, bn0, wt);
              }
            }
            sa += stage_bytes; fb += 8; eb += 8;
            if (++stage == S) { stage = 0; phase ^= 1u; sa = smem_base; fb = full0; eb = empty0; }
          }
        }
      }
      if constexpr (PAIR) {
        // tail: every multicast stage-release of the leader has landed on this CTA's barriers before it may exit
        for (int i = 0; ok && i < S; ++i) {
          if (!mbar_wait_a(eb, phase ^ 1u, ac, 1)) break;
          eb += 8;
          if (++stage == S) { stage = 0; phase ^= 1u; eb = empty0; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane; pairs: the leader's)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool ok = true;
      const uint32_t dhi = desc_hi(1024, 2);
      const uint32_t a_lo0 = desc_lo(smem_base, 16);
      const uint32_t stage16 = stage_bytes >> 4, areg16 = a_region >> 4, b16 = b_chunk >> 4;
      uint32_t a_lo = a_lo0, fb = full0, eb = empty0;
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc_) {
        if constexpr (PAIR) Op::mma2(d, a, b, id, acc_);
        else Op::mma(d, a, b, id, acc_);
      };
      for (int tile = tile0; ok && tile < P.total_tiles; tile += tstep) {
        const TileCoord tc = decode_tile<Params, PAIR>(P, tile, 0);
        const TapGemmGroup& G = P.g[tc.gi];
        if constexpr (PAIR) {
          if (!mbar_wait_cluster_a(smem_u32(&tempty_bar[as]), aphase ^ 1u, ac.smem_flag, ac.status, 2)) break;
        } else {
          if (!mbar_wait(&tempty_bar[as], aphase ^ 1u, ac, 2)) break;
        }
        tc_fence_after();
        const uint32_t idesc = Op::idesc(PAIR ? 256 : 128, G.block_n, 0, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccCols);
        const int ntap = G.phase[tc.ph].tap_count;
        const int kchunks = G.kchunks;
        // a stage carries kst chunks except the last one of the K range; the last chunk of the whole K may be partial
        const int c0 = tc.c0, c1 = tc.c1;
        uint32_t acc = 0;
        for (int t = tc.t0; ok && t < tc.t1; ++t) {
          for (int c = c0; c < c1; c += kst) {
            if (!mbar_wait_a(fb, phase, ac, 3)) { ok = false; break; }
            tc_fence_after();
            const uint32_t b_lo = a_lo + areg16;
            const int nch = min(kst, c1 - c);
            const int n_mma = (nch - 1) * 4 + ((c + nch == kchunks) ? G.last_mmas : 4);
            if (n_mma == 4 * kst) {  // a full stage: kst * 4 MMAs
              mma(d_tmem, desc_join(a_lo, dhi), desc_join(b_lo, dhi), idesc, acc);
              mma(d_tmem, desc_join(a_lo + 2, dhi), desc_join(b_lo + 2, dhi), idesc, 1);
              mma(d_tmem, desc_join(a_lo + 4, dhi), desc_join(b_lo + 4, dhi), idesc, 1);
              mma(d_tmem, desc_join(a_lo + 6, dhi), desc_join(b_lo + 6, dhi), idesc, 1);
              if (kst > 1) {
                const uint32_t a2 = a_lo + (kAStageBytes >> 4), b2 = b_lo + b16;
                mma(d_tmem, desc_join(a2, dhi), desc_join(b2, dhi), idesc, 1);
                mma(d_tmem, desc_join(a2 + 2, dhi), desc_join(b2 + 2, dhi), idesc, 1);
                mma(d_tmem, desc_join(a2 + 4, dhi), desc_join(b2 + 4, dhi), idesc, 1);
                mma(d_tmem, desc_join(a2 + 6, dhi), desc_join(b2 + 6, dhi), idesc, 1);
              }
            } else {
#pragma unroll
              for (int m = 0; m < 8; ++m) {
                if (m < n_mma) {
                  const uint32_t off = (uint32_t)(m & 3) * 2u;
                  const uint32_t aj = a_lo + off + ((m >> 2) ? (kAStageBytes >> 4) : 0u);
                  const uint32_t bj = b_lo + off + ((m >> 2) ? b16 : 0u);
                  mma(d_tmem, desc_join(aj, dhi), desc_join(bj, dhi), idesc, m ? 1u : acc);
                }
              }
            }
            acc = 1;
            if constexpr (PAIR) tc_commit2_a(eb);
            else tc_commit_a(eb);
            a_lo += stage16; fb += 8; eb += 8;
            if (++stage == S) { stage = 0; phase ^= 1u; a_lo = a_lo0; fb = full0; eb = empty0; }
          }
        }
        if (!ok) break;
        if constexpr (PAIR) tc_commit2_a(smem_u32(&tfull_bar[as]));
        else tc_commit(&tfull_bar[as]);
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
    for (int tile = tile0; tile < P.total_tiles; tile += tstep) {
      TileCoord tc = decode_tile<Params, PAIR>(P, tile, rank);
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
      uint16_t* po16 = G.out16.ptr ? G.out16.ptr + (long long)n * G.out16.sn + (long long)yo * G.out16.sh + (long long)xo * G.out16.sw : nullptr;
      if (!G.out.ptr) po = nullptr;
      const uint32_t t_addr = tmem_base + (uint32_t)(as * kAccCols) + ((uint32_t)(q * 32) << 16);
      const int col_base = tc.nt * G.block_n;
      const EpiArgs E{G.bias, G.slopes, G.cout_valid, G.epilogue, G.round_tf32, G.vec_ok, G.slope};
      if (!PAIR && G.ksplit > 1) {
        // ---- split-K: park the partial accumulator, then the last of the tile's ksplit CTAs finishes the tile
        float* ws = G.split_ws + ((size_t)(tc.base * G.ksplit + tc.ks) * 128 + row) * G.block_n;
        for (int c0 = ((warp - 4) >> 2) * 16; c0 < G.block_n; c0 += 16 * kEpiPerQuarter) {
          uint32_t r[16];
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(ws + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);      // the accumulator is free: the next tile's MMAs overlap the fix-up
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"r"(kTapGemmThreads - 128) : "memory");
        if (threadIdx.x == 128) {
          const int prev = atomicAdd(G.split_cnt + tc.base, 1);
          split_last = (prev == G.ksplit - 1) ? 1 : 0;
          if (split_last) G.split_cnt[tc.base] = 0;   // ready for the next launch that uses this workspace slot
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kTapGemmThreads - 128) : "memory");
        if (split_last) {
          __threadfence();
          const float* w0 = G.split_ws + ((size_t)(tc.base * G.ksplit) * 128 + row) * G.block_n;
          const size_t rstride = (size_t)128 * G.block_n;
          for (int c0 = ((warp - 4) >> 2) * 16; c0 < G.block_n; c0 += 16 * kEpiPerQuarter) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            for (int k2 = 0; k2 < G.ksplit; ++k2) {      // fixed order: bit-reproducible whatever the arrival order
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(w0 + k2 * rstride + c0 + j));
                acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
              }
            }
            uint32_t r[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(acc[j]);
            if (valid) epilogue_store16(r, E, col_base + c0, po, p1, p2, pm, po16);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kTapGemmThreads - 128) : "memory");   // split_last is rewritten by the next tile
        as ^= 1;
        if (as == 0) aphase ^= 1u;
        continue;
      }
      for (int c0 = ((warp - 4) >> 2) * 16; c0 < G.block_n; c0 += 16 * kEpiPerQuarter) {
        uint32_t r[16];
        tmem_ld16(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (valid) epilogue_store16(r, E, col_base + c0, po, p1, p2, pm, po16);
      }
      tc_fence_before();
      if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(smem_u32(&tempty_bar[as]), 0);
      } else {
        mbar_arrive(&tempty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while the other may still signal it
  tc_fence_after();
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// explicit instantiations used by api.cu
template __global__ void tapgemm_kernel<TapGemmParams, false, false>(const __grid_constant__ TapGemmParams, int*);
template __global__ void tapgemm_kernel<TapGemmParams1, false, false>(const __grid_constant__ TapGemmParams1, int*);
template __global__ void tapgemm_kernel<TapGemmParams, true, false>(const __grid_constant__ TapGemmParams, int*);
template __global__ void tapgemm_kernel<TapGemmParams1, true, false>(const __grid_constant__ TapGemmParams1, int*);
template __global__ void tapgemm_kernel<TapGemmParams, false, true>(const __grid_constant__ TapGemmParams, int*);
template __global__ void tapgemm_kernel<TapGemmParams1, false, true>(const __grid_constant__ TapGemmParams1, int*);
template __global__ void tapgemm_kernel<TapGemmParams, true, true>(const __grid_constant__ TapGemmParams, int*);
template __global__ void tapgemm_kernel<TapGemmParams1, true, true>(const __grid_constant__ TapGemmParams1, int*);

}  // namespace tpg
