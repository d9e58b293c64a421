// Flat-slab convolution for small feature maps (stride 1, "same" padding) on Blackwell tensor cores (sm_100a): the local
// pathways' 3x3 layers at 40x40 ... 10x10 (D_and_G_model.py:18-160 of the reference), forward and input gradient.
//
// The multi-tap GEMM kernel (tapgemm.cu) loads one shifted activation box and one weight tile per (tap, K chunk): on these
// layers it moves 20x the unique operand bytes from L2 and the tensor pipe waits for it (ncu: 33 % busy).  Here the M
// dimension enumerates the pixels of a unit with the PADDED pitch Wp = W + k - 1, so that ONE slab per K chunk - a TMA box
// {chunk, Wp, R rows, bn images} with its zero halo produced by out-of-bounds fill - serves all k*k taps: tap (r, j) is the
// same slab read (r * Wp + j) pixel rows (of 128 B) further down, i.e. just another shared-memory descriptor start address
// (the 128B swizzle is a function of the address, so a row-shifted start reads the same swizzled data).  The T 128-row
// tiles of a unit share every weight tile (one accumulator each).  M rows that fall into the halo columns / rows compute
// garbage that is never stored (78-89 % of the rows are real pixels on the local-pathway shapes).
//
//   warp 0: weight-tile TMA producer   warp 3: slab TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4..: epilogue
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kFcMaxSlots = 16;

template <class Params>
__device__ __forceinline__ int flat_group_of(const Params& P, int tile) {
  int gi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups; ++i)
    if (i < P.ngroups && tile >= P.g[i].tile_begin) gi = i;
  return gi;
}

struct FlatUnit {
  int nt, y0, n0;
};
__device__ __forceinline__ FlatUnit flat_unit(const FlatGroup& G, int local) {
  FlatUnit u;
  u.nt = local % G.n_tiles;
  const int unit = local / G.n_tiles;
  u.y0 = (unit % G.units_h) * G.ur;
  u.n0 = (unit / G.units_h) * G.bn;
  return u;
}

template <bool BF16>
__global__ void __launch_bounds__(kConvThreads, 1) flatconv_kernel(const __grid_constant__ FlatConvParams P, int* status) {
  using Op = Opnd<BF16>;
  constexpr int CH = Op::kChunk;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kFcMaxSlots];
  __shared__ __align__(8) uint64_t a_empty[kFcMaxSlots];
  __shared__ __align__(8) uint64_t b_full[kFcMaxSlots];
  __shared__ __align__(8) uint64_t b_empty[kFcMaxSlots];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFcMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kConvThreads - 128);
    }
    abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // everything above touched only shared / tensor memory; global memory from here on
  const uint32_t tmem_base = tmem_base_s;
  AbortCtl ac{&abort_flag, status};

  const int k = P.k, ntaps = k * k;
  const int AS = P.a_slots, BS = P.b_slots;
  const uint32_t slab_bytes = (uint32_t)P.slab_bytes, b_bytes = (uint32_t)P.b_bytes;
  const uint32_t smem_a = smem_u32(smem), smem_b = smem_a + (uint32_t)AS * slab_bytes;
  const uint32_t af0 = smem_u32(&a_full[0]), ae0 = smem_u32(&a_empty[0]);
  const uint32_t bf0 = smem_u32(&b_full[0]), be0 = smem_u32(&b_empty[0]);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: weight tiles
    if (elect_one()) {
      int bs_ = 0;
      uint32_t bph = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        const FlatGroup& G = P.g[flat_group_of(P, tile)];
        const FlatUnit u = flat_unit(G, tile - G.tile_begin);
        const int bn0 = u.nt * G.block_n;
        const uint32_t b_tx = (uint32_t)G.block_n * 128u;
        for (int c = 0; ok && c < G.kchunks; ++c) {
          for (int t = 0; t < ntaps; ++t) {
            if (!mbar_wait_a(be0 + 8u * bs_, bph ^ 1u, ac, 32)) { ok = false; break; }
            mbar_arrive_expect_tx_a(bf0 + 8u * bs_, b_tx);
            tma_load_3d_a(smem_b + (uint32_t)bs_ * b_bytes, &G.bmap, bf0 + 8u * bs_, c * CH, bn0, P.wtap[t]);
            if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer: activation slabs.  Its own warp, so
    // that the slab of the next K chunk (or of the next unit) is requested as soon as a slab slot frees up instead of
    // behind the k*k weight tiles of the current chunk - the slab is the long-latency load of this kernel
    if (elect_one()) {
      int as_ = 0;
      uint32_t aph = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        const FlatGroup& G = P.g[flat_group_of(P, tile)];
        const FlatUnit u = flat_unit(G, tile - G.tile_begin);
        for (int c = 0; c < G.kchunks; ++c) {
          if (!mbar_wait_a(ae0 + 8u * as_, aph ^ 1u, ac, 31)) { ok = false; break; }
          mbar_arrive_expect_tx_a(af0 + 8u * as_, (uint32_t)G.slab_tx);
          tma_load_4d_a(smem_a + (uint32_t)as_ * slab_bytes, &G.amap, af0 + 8u * as_, c * CH, P.dx0, u.y0 + P.dy0, u.n0);
          if (++as_ == AS) { as_ = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int as_ = 0, bs_ = 0, buf = 0;
      uint32_t aph = 0, bph = 0, tph = 0;
      bool ok = true;
      const uint32_t dhi = desc_hi(1024, 2);
      const uint32_t a_lo0 = desc_lo(smem_a, 16), b_lo0 = desc_lo(smem_b, 16);
      const uint32_t slab16 = slab_bytes >> 4, b16 = b_bytes >> 4;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        const FlatGroup& G = P.g[flat_group_of(P, tile)];
        if (!mbar_wait(&tempty_bar[buf], tph ^ 1u, ac, 33)) break;
        tc_fence_after();
        const uint32_t idesc = Op::idesc(128, G.block_n, 0, 0);
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
        const int T = G.T, Wp = G.Wp, kchunks = G.kchunks;
        const uint32_t dstep = (uint32_t)G.block_n;
        for (int c = 0; ok && c < kchunks; ++c) {
          const int nm = (c == kchunks - 1) ? G.last_mmas : 4;
          if (!mbar_wait_a(af0 + 8u * as_, aph, ac, 34)) { ok = false; break; }
          tc_fence_after();
          const uint32_t a_slab = a_lo0 + (uint32_t)as_ * slab16;
          uint32_t first = (c == 0) ? 0u : 1u;
          for (int r = 0; ok && r < k; ++r) {
            uint32_t a_tap = a_slab + (uint32_t)(r * Wp) * 8u;   // pixel rows of 128 B = 8 x 16 B
            for (int j = 0; j < k; ++j, a_tap += 8u) {
              if (!mbar_wait_a(bf0 + 8u * bs_, bph, ac, 35)) { ok = false; break; }
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + (uint32_t)bs_ * b16;
              uint32_t a_lo = a_tap, d = d0;
              for (int t = 0; t < T; ++t, a_lo += 128u * 8u, d += dstep) {
                if (nm == 4) {
                  Op::mma(d, desc_join(a_lo, dhi), desc_join(b_lo, dhi), idesc, first);
                  Op::mma(d, desc_join(a_lo + 2, dhi), desc_join(b_lo + 2, dhi), idesc, 1);
                  Op::mma(d, desc_join(a_lo + 4, dhi), desc_join(b_lo + 4, dhi), idesc, 1);
                  Op::mma(d, desc_join(a_lo + 6, dhi), desc_join(b_lo + 6, dhi), idesc, 1);
                } else {
                  for (int q = 0; q < nm; ++q)
                    Op::mma(d, desc_join(a_lo + 2 * q, dhi), desc_join(b_lo + 2 * q, dhi), idesc, q ? 1u : first);
                }
              }
              first = 1u;
              tc_commit_a(be0 + 8u * bs_);
              if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
            }
          }
          if (!ok) break;
          tc_commit_a(ae0 + 8u * as_);
          if (++as_ == AS) { as_ = 0; aph ^= 1u; }
        }
        if (!ok) break;
        tc_commit(&tfull_bar[buf]);
        buf ^= 1;
        if (buf == 0) tph ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    int buf = 0;
    uint32_t tph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const FlatGroup& G = P.g[flat_group_of(P, tile)];
      const FlatUnit u = flat_unit(G, tile - G.tile_begin);
      if (!mbar_wait(&tfull_bar[buf], tph, ac, 36)) break;
      tc_fence_after();
      const int col_base = u.nt * G.block_n;
      const EpiArgs E{G.bias, G.slopes, G.cout_valid, G.epilogue, G.round_tf32, G.vec_ok, G.slope};
      for (int t = 0; t < G.T; ++t) {
        const int m = t * 128 + q * 32 + lane;   // flat (padded-pitch) pixel of this TMEM lane
        const int img = m / G.ipitch;
        const int rem = m - img * G.ipitch;
        const int yy = rem / G.Wp;
        const int x = rem - yy * G.Wp;
        const int y = u.y0 + yy, n = u.n0 + img;
        const bool valid = (img < G.bn) && (yy < G.ur) && (y < G.H) && (x < G.W) && (n < G.Nimg);
        const long long pix = (long long)n * G.out.sn + (long long)y * G.out.sh + (long long)x * G.out.sw;
        float* po = G.out.ptr ? G.out.ptr + pix : nullptr;
        const float* p1 = G.add1.ptr ? G.add1.ptr + (long long)n * G.add1.sn + (long long)y * G.add1.sh + (long long)x * G.add1.sw : nullptr;
        const float* p2 = G.add2.ptr ? G.add2.ptr + (long long)n * G.add2.sn + (long long)y * G.add2.sh + (long long)x * G.add2.sw : nullptr;
        const float* pm = G.mask.ptr ? G.mask.ptr + (long long)n * G.mask.sn + (long long)y * G.mask.sh + (long long)x * G.mask.sw : nullptr;
        uint16_t* po16 = G.out16.ptr ? G.out16.ptr + (long long)n * G.out16.sn + (long long)y * G.out16.sh + (long long)x * G.out16.sw : nullptr;
        const uint32_t t_addr = tmem_base + (uint32_t)(buf * 256 + t * G.block_n) + ((uint32_t)(q * 32) << 16);
        for (int c0 = ((warp - 4) >> 2) * 16; c0 < G.block_n; c0 += 16 * kEpiPerQuarter) {
          uint32_t r[16];
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (valid) epilogue_store16(r, E, col_base + c0, po, p1, p2, pm, po16);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      buf ^= 1;
      if (buf == 0) tph ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template __global__ void flatconv_kernel<false>(const __grid_constant__ FlatConvParams, int*);
template __global__ void flatconv_kernel<true>(const __grid_constant__ FlatConvParams, int*);

}  // namespace tpg
