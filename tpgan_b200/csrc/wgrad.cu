// Weight-gradient GEMM on Blackwell tensor cores (sm_100a).
//
//   dW[tap][m][n] += sum over pixels  P[pix, m] * Q_tap[pix, n]
//
// P is the un-shifted tensor (dY for a conv, x for a transposed conv), Q the tap-shifted one (x / dY), both NHWC.
// The reduction (GEMM K) dimension is the pixel index, so both operands are "MN-major" in shared memory: a TMA box
// {32 channels, bw, bh, bn} lands as [pixel][32ch] rows of 128 B, which is exactly the canonical MN-major tf32
// layout (128B swizzle with 32B atoms: 4 K-rows x 128 B per atom, TMA mode SWIZZLE_128B_ATOM_32B).  TF32 UMMA accepts
// MN-major A and B in that layout, so no transpose is ever materialised.
// Zero padding and the conv stride are again TMA out-of-bounds fill and parity-plane tensor maps.
//
// Output tile = (tap group, 128-channel M tile group, N tile); the (tile, pixel chunk) work list is cut into equal
// contiguous slices, one per persistent CTA (one CTA per SM), and the partial sums of tiles that straddle CTAs are
// combined with fp32 red.global.add into the forward-packed weight-gradient tensor.
// Replaces the weight half of aten::convolution_backward for ModificationLayer.py:101,189.
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kWTmemCols = 512;
constexpr int kWAccCols = 256;
constexpr int kWgradThreads = kConvThreads;   // warps 4..11 are the epilogue (two per TMEM lane quarter, alternating 16-column groups)

// Balanced ("stream-K") schedule.  The reduction of every group is cut into K blocks (pixel ranges sized so that the
// activations of one block stay in L2); the work of one K block is the list of (output tile, pixel chunk) pairs,
// tile-major, and CTA b owns the contiguous slice [b*S/grid, (b+1)*S/grid) of EVERY block's list.  All CTAs therefore
// stream the same pixel range at the same time (each activation byte comes from HBM once) and every CTA gets the same
// number of chunks (+-1) per block whatever the tile count - no partial last wave.  A slice is walked as segments (one
// output tile, a chunk range); a segment that covers its tile's whole reduction is written with plain stores, every
// other one is added with vector reds.
struct Segment {
  int gi, tg, mg, nt;
  int c_begin, c_end;
};

template <class Params>
struct SegmentWalk {
  const Params& P;
  int gi, kb, pos, end, base, len;
  __device__ __forceinline__ SegmentWalk(const Params& P_) : P(P_), gi(0), kb(-1), pos(0), end(0), base(0), len(1) {}
  __device__ __forceinline__ bool next(Segment& s) {
    while (pos >= end) {
      if (gi >= P.ngroups) return false;
      if (++kb >= P.g[gi].nkb) {
        kb = 0;
        if (++gi >= P.ngroups) return false;
      }
      const WgradGroup& G = P.g[gi];
      base = kb * G.kb_chunks;
      len = min(G.kb_chunks, G.chunks - base);
      const long long S = (long long)G.tiles * len;
      if (P.whole_tiles) {   // deterministic mode: contiguous runs of WHOLE tiles, every dW element has one writer
        pos = (int)(((long long)G.tiles * blockIdx.x) / gridDim.x) * len;
        end = (int)(((long long)G.tiles * (blockIdx.x + 1)) / gridDim.x) * len;
      } else {
        pos = (int)((S * blockIdx.x) / gridDim.x);
        end = (int)((S * (blockIdx.x + 1)) / gridDim.x);
      }
    }
    const WgradGroup& G = P.g[gi];
    int tile = pos / len;
    s.gi = gi;
    s.c_begin = base + (pos - tile * len);
    s.c_end = min(base + len, s.c_begin + (end - pos));
    s.tg = tile % G.tap_groups;
    tile /= G.tap_groups;
    s.nt = tile % G.n_tiles;
    s.mg = tile / G.n_tiles;
    pos += s.c_end - s.c_begin;
    return true;
  }
};

// Slab mode pads every tap group to tpu slots; unused slots carry wtap == 255.
__device__ __forceinline__ int slab_group_ntap(const WgradGroup& G, int tap0) {
  int n = G.tpu;
  while (n > 1 && G.taps[tap0 + n - 1].wtap == 255) --n;
  return n;
}

template <class Params, bool BF16>
__global__ void __launch_bounds__(kWgradThreads, 1) wgrad_kernel(const __grid_constant__ Params P, int* status) {
  using Op = Opnd<BF16>;
  constexpr int CH = Op::kChunk;        // channels per 128-byte row of an MN-major operand chunk (32 tf32 / 64 bf16)
  constexpr int MCH = 128 / CH;         // chunks that make up the M = 128 rows of one accumulator
  constexpr int KR = Op::kMmaK;         // pixel rows (GEMM K) per MMA
  constexpr uint32_t KSTEP16 = (uint32_t)KR * 128u / 16u;   // descriptor advance per MMA (16-byte units)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = P.stages;
  const uint32_t stage_bytes = (uint32_t)(P.a_stage_bytes + P.b_stage_bytes);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  // Zero the whole ring once when some box does not fill its K rows (pixel count not a multiple of 8): those rows must
  // read as zeros for the lifetime of the kernel.  (M rows / N columns beyond the loaded channel chunks may hold anything:
  // they only reach accumulator rows / columns that are never stored.)
  if (P.need_zero) {
    uint4* p = reinterpret_cast<uint4*>(smem);
    const int n16 = P.ring_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kWgradThreads - 128);
    }
    abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, kWTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // everything above touched only shared / tensor memory; global memory from here on
  const uint32_t tmem_base = tmem_base_s;
  AbortCtl ac{&abort_flag, status};
  const int nbuf = P.nbuf;   // accumulator buffers (uniform over the groups of a launch)

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t sa = smem_base, fb = full0, eb = empty0;
      bool ok = true;
      SegmentWalk<Params> walk(P);
      Segment u;
      while (ok && walk.next(u)) {
        const WgradGroup& G = P.g[u.gi];
        const CUtensorMap* pm = &G.pmap;
        const int mt0 = u.mg * G.mpu;
        const int pc0 = mt0 * 128;
        const int mch = min(MCH * G.mpu, (G.m_valid - pc0 + CH - 1) / CH);   // P chunks of this unit (all its M tiles)
        const int tap0 = u.tg * G.tpu;
        const int ntap = min(G.tpu, G.ntaps - tap0);
        const int qc0 = G.slab ? u.nt * G.ncpt * CH : (G.tpu > 1 ? 0 : u.nt * G.block_n);
        const int ncpt = (G.tpu > 1 || G.slab) ? G.ncpt : min(G.block_n / CH, (G.n_valid - qc0 + CH - 1) / CH);
        const uint32_t box_bytes = (uint32_t)G.p_rows * 128u;
        const bool slab = G.slab != 0;
        const uint32_t q_stride = (uint32_t)G.q_chunk_bytes;
        // K rows the pixel box leaves unwritten: in a launch that mixes box shapes they would hold another group's pixels,
        // so every stage overwrites them (P side only: 0 * finite = 0) with a box fetched fully out of bounds
        const uint32_t tail_bytes = P.zero_tail ? (uint32_t)G.kp * 128u - box_bytes : 0u;
        const uint32_t tx = (slab ? box_bytes * (uint32_t)mch + (uint32_t)G.q_rows * 128u * (uint32_t)G.ncpt
                                  : box_bytes * (uint32_t)(mch + ntap * ncpt)) + tail_bytes * (uint32_t)mch;
        const uint32_t chunk_stride = (uint32_t)G.kp * 128u;
        const int c_begin = u.c_begin, c_end = u.c_end;
        const int tiles_w = G.tiles_w, tiles_h = G.tiles_h, bw = G.bw, bh = G.bh, bn = G.bn;
        const TapDesc* taps = &G.taps[tap0];
        const bool single_tap = (G.tpu == 1);
        const TapDesc tapA = taps[0];
        const CUtensorMap* qm0 = &G.qmap[tapA.plane];
        const int tdx0 = tapA.dx, tdy0 = tapA.dy;
        int wb = c_begin % tiles_w, r = c_begin / tiles_w;
        int hb = r % tiles_h, nb = r / tiles_h;
        for (int c = c_begin; c < c_end; ++c) {
          const int x0 = wb * bw, y0 = hb * bh, n0 = nb * bn;
          if (!mbar_wait_a(eb, phase ^ 1u, ac, 11)) { ok = false; break; }
          uint32_t sb = sa + (uint32_t)P.a_stage_bytes;
          mbar_arrive_expect_tx_a(fb, tx);
          for (int i = 0; i < mch; ++i) tma_load_4d_a(sa + (uint32_t)i * chunk_stride, pm, fb, pc0 + i * CH, x0, y0, n0);
          if (tail_bytes)
            for (int i = 0; i < mch; ++i)
              tma_load_4d_a(sa + (uint32_t)i * chunk_stride + box_bytes, &G.pzero, fb, pc0 + i * CH, G.Wp, 0, 0);
          if (slab) {
            for (int i = 0; i < ncpt; ++i)
              tma_load_4d_a(sb + (uint32_t)i * q_stride, &G.qslab, fb, qc0 + i * CH, x0 + tdx0, y0 + tdy0, n0);
          } else if (single_tap) {
            for (int i = 0; i < ncpt; ++i)
              tma_load_4d_a(sb + (uint32_t)i * chunk_stride, qm0, fb, qc0 + i * CH, x0 + tdx0, y0 + tdy0, n0);
          } else {
            for (int t = 0; t < ntap; ++t) {
              const TapDesc tap = taps[t];
              const CUtensorMap* qm = &G.qmap[tap.plane];
              for (int i = 0; i < ncpt; ++i) {
                tma_load_4d_a(sb, qm, fb, qc0 + i * CH, x0 + tap.dx, y0 + tap.dy, n0);
                sb += chunk_stride;
              }
            }
          }
          sa += stage_bytes; fb += 8; eb += 8;
          if (++stage == S) { stage = 0; phase ^= 1u; sa = smem_base; fb = full0; eb = empty0; }
          if (++wb == tiles_w) { wb = 0; if (++hb == tiles_h) { hb = 0; ++nb; } }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      bool ok = true;
      // MN-major operands: tf32 = 32-byte-atom 128B swizzle (layout type 1, 4-row K atoms of 512 B); bf16 = plain 128B
      // swizzle (layout type 2, 8-row K atoms of 1024 B)
      const uint32_t dhi = BF16 ? desc_hi(1024, 2) : desc_hi(512, 1);
      const uint32_t stage16 = stage_bytes >> 4, areg16 = (uint32_t)P.a_stage_bytes >> 4;
      uint32_t fb = full0, eb = empty0, sbase16 = (smem_base & 0x3FFFFu) >> 4;
      const uint32_t sbase16_0 = sbase16;
      SegmentWalk<Params> walk(P);
      Segment u;
      while (ok && walk.next(u)) {
        const WgradGroup& G = P.g[u.gi];
        if (!mbar_wait(&tempty_bar[as], aphase ^ 1u, ac, 12)) break;
        tc_fence_after();
        const uint32_t idesc = Op::idesc(128, G.block_n, 1, 1);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kWAccCols);
        const uint32_t chunk16 = ((uint32_t)G.kp * 128u) >> 4;
        const uint32_t lbo_field = (chunk16 & 0x3FFFu) << 16;
        const int kgroups = G.kp / KR;
        const int nm = min(G.mpu, G.m_tiles - u.mg * G.mpu);   // M tiles (accumulators) of this unit
        const int c_begin = u.c_begin, c_end = u.c_end;
        // slab mode constants of this segment (hoisted: the issuing thread is the critical resource)
        const int s_ntap = G.slab ? slab_group_ntap(G, u.tg * G.tpu) : 0;
        const uint32_t idesc_s = Op::idesc(128, s_ntap * CH, 1, 1);
        const uint32_t qchunk16 = (uint32_t)G.q_chunk_bytes >> 4;
        const int s_ncpt = G.ncpt;
        const uint32_t s_dstep = (uint32_t)(G.tpu * CH);
        const bool is_slab = G.slab != 0;
        uint32_t acc = 0;
        for (int c = c_begin; c < c_end; ++c) {
          if (!mbar_wait_a(fb, phase, ac, 13)) { ok = false; break; }
          tc_fence_after();
          const uint32_t b_lo = (sbase16 + areg16) | lbo_field;
          if (is_slab) {
            // one MMA per 32-channel Q chunk covers all taps of the group: the N dimension walks the taps with a leading
            // byte offset of ONE pixel row (128 B) through the same slab (im2col by descriptor), D columns = [tap][32 ch]
            const uint32_t bs_lo = (sbase16 + areg16) | (8u << 16);
            uint32_t d = d_tmem;
            for (int mi = 0; mi < nm; ++mi) {
              const uint32_t a_lo = (sbase16 + (uint32_t)mi * (uint32_t)MCH * chunk16) | lbo_field;
              uint32_t bi_lo = bs_lo;
              for (int i = 0; i < s_ncpt; ++i) {
                Op::mma(d, desc_join(a_lo, dhi), desc_join(bi_lo, dhi), idesc_s, acc);
                for (int k = 1; k < kgroups; ++k)
                  Op::mma(d, desc_join(a_lo + KSTEP16 * k, dhi), desc_join(bi_lo + KSTEP16 * k, dhi), idesc_s, 1);
                d += s_dstep;
                bi_lo += qchunk16;
              }
            }
          } else if (nm == 1 && kgroups == 8) {
            const uint32_t a_lo = sbase16 | lbo_field;
            Op::mma(d_tmem, desc_join(a_lo, dhi), desc_join(b_lo, dhi), idesc, acc);
#pragma unroll
            for (int k = 1; k < 8; ++k)
              Op::mma(d_tmem, desc_join(a_lo + KSTEP16 * k, dhi), desc_join(b_lo + KSTEP16 * k, dhi), idesc, 1);
          } else
          for (int mi = 0; mi < nm; ++mi) {
            const uint32_t a_lo = (sbase16 + (uint32_t)mi * (uint32_t)MCH * chunk16) | lbo_field;
            const uint32_t d = d_tmem + (uint32_t)(mi * G.block_n);
            if (kgroups == 4) {
              Op::mma(d, desc_join(a_lo, dhi), desc_join(b_lo, dhi), idesc, acc);
              Op::mma(d, desc_join(a_lo + KSTEP16, dhi), desc_join(b_lo + KSTEP16, dhi), idesc, 1);
              Op::mma(d, desc_join(a_lo + 2 * KSTEP16, dhi), desc_join(b_lo + 2 * KSTEP16, dhi), idesc, 1);
              Op::mma(d, desc_join(a_lo + 3 * KSTEP16, dhi), desc_join(b_lo + 3 * KSTEP16, dhi), idesc, 1);
            } else {
              for (int k = 0; k < kgroups; ++k)
                Op::mma(d, desc_join(a_lo + KSTEP16 * k, dhi), desc_join(b_lo + KSTEP16 * k, dhi), idesc, k ? 1u : acc);
            }
          }
          acc = 1;
          tc_commit_a(eb);
          sbase16 += stage16; fb += 8; eb += 8;
          if (++stage == S) { stage = 0; phase ^= 1u; sbase16 = sbase16_0; fb = full0; eb = empty0; }
        }
        if (!ok) break;
        tc_commit(&tfull_bar[as]);
        if (++as == nbuf) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> dW (stores or vector reds)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    SegmentWalk<Params> walk(P);
    Segment u;
    while (walk.next(u)) {
      const WgradGroup& G = P.g[u.gi];
      if (!mbar_wait(&tfull_bar[as], aphase, ac, 14)) break;
      tc_fence_after();
      const bool nonempty = true;
      // sole writer of this dW tile (whole reduction in this segment, nothing to add to): plain stores
      const bool single = (u.c_begin == 0) && (u.c_end == G.chunks) && (G.accumulate == 0);
      const int nm = min(G.mpu, G.m_tiles - u.mg * G.mpu);
      const int tap0 = u.tg * G.tpu;
      const int cols_per_tap = (G.tpu > 1) ? G.ncpt * CH : G.block_n;
      // 16 accumulator columns of row m -> dw (columns n0 .. n0+15 of tap `tap`)
      auto store16 = [&](const uint32_t (&r)[16], const TapDesc tap, int m, int n0) {
        if (G.im2col_k) {        // column n of vertical tap r = (horizontal tap n / 4, input channel n % 4)
          const int kk = G.im2col_k, rr = tap.wtap;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + j, jt = n >> 2, c = n & 3;
            if (jt < kk && c < G.im2col_cin) {
              float* d = G.dw + ((size_t)(rr * kk + jt) * G.rows_pad + m) * G.k_pad + c;
              if (single) *d = __uint_as_float(r[j]); else atomicAdd(d, __uint_as_float(r[j]));
            }
          }
          return;
        }
        float* dw = G.dw + (size_t)tap.wtap * G.rows_pad * G.k_pad;
        if (G.transpose_out) {   // dw[n][m]: lanes hold consecutive m -> every red is a coalesced 128 B row segment
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + j;
            if (n < G.n_valid) {
              float* d = dw + (size_t)n * G.k_pad + m;
              if (single) *d = __uint_as_float(r[j]); else atomicAdd(d, __uint_as_float(r[j]));
            }
          }
        } else {                 // dw[m][n]: each lane owns 16 consecutive floats of its row -> 16-byte vector ops
          float* d = dw + (size_t)m * G.k_pad + n0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (n0 + j + 3 < G.n_valid) {
              if (single) {
                *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + j), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                             : "memory");
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (n0 + j + i < G.n_valid) {
                  if (single) d[j + i] = __uint_as_float(r[j + i]); else atomicAdd(d + j + i, __uint_as_float(r[j + i]));
                }
            }
          }
        }
      };
      if (G.slab) {
        const int ntap = slab_group_ntap(G, tap0);
        for (int mi = 0; mi < nm; ++mi) {
          const int m = (u.mg * G.mpu + mi) * 128 + row;
          const bool mvalid = m < G.m_valid;
          // 16-column blocks (chunk i, tap t, part h of the chunk's CH columns) of this M tile, round-robin over the warps
          // of the quarter
          constexpr int HP = CH / 16;
          const int nblk = G.ncpt * ntap * HP;
          for (int blk = (warp - 4) >> 2; blk < nblk; blk += kEpiPerQuarter) {
            const int h = blk % HP;
            const int t = (blk / HP) % ntap;
            const int i = (blk / HP) / ntap;
            const uint32_t t_addr = tmem_base + (uint32_t)(as * kWAccCols + ((mi * G.ncpt + i) * G.tpu + t) * CH + h * 16) +
                                    ((uint32_t)(q * 32) << 16);
            uint32_t r[16];
            tmem_ld16(t_addr, r);
            tmem_ld_wait();
            const int n0 = (u.nt * G.ncpt + i) * CH + h * 16;
            if (mvalid && n0 < G.n_valid) store16(r, G.taps[tap0 + t], m, n0);
          }
        }
      } else
      for (int mi = 0; mi < nm; ++mi) {
        const int m = (u.mg * G.mpu + mi) * 128 + row;
        const bool mvalid = m < G.m_valid;
        const uint32_t t_addr = tmem_base + (uint32_t)(as * kWAccCols + mi * G.block_n) + ((uint32_t)(q * 32) << 16);
        for (int c0 = ((warp - 4) >> 2) * 16; c0 < G.block_n; c0 += 16 * kEpiPerQuarter) {
          uint32_t r[16];
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          const int ti = c0 / cols_per_tap;                 // cols_per_tap is a multiple of 32 (or block_n itself)
          if (!(mvalid && nonempty) || tap0 + ti >= G.ntaps) continue;
          const int n0 = ((G.tpu > 1) ? 0 : u.nt * G.block_n) + (c0 - ti * cols_per_tap);
          store16(r, G.taps[tap0 + ti], m, n0);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == nbuf) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kWTmemCols);
}

template __global__ void wgrad_kernel<WgradParams, false>(const __grid_constant__ WgradParams, int*);
template __global__ void wgrad_kernel<WgradParams1, false>(const __grid_constant__ WgradParams1, int*);
template __global__ void wgrad_kernel<WgradParams, true>(const __grid_constant__ WgradParams, int*);
template __global__ void wgrad_kernel<WgradParams1, true>(const __grid_constant__ WgradParams1, int*);

}  // namespace tpg
