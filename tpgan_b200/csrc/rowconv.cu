// Row-tile convolution for the full-resolution (W = 128) stride-1 layers of the global pathway - the layers that carry
// 65 % of the step's FLOPs (206->206 k5, 75->75 k7, 64->64 k7, 206->64 k5, 64->64 k3, ... all at 128x128).
//
// The generic multi-tap kernel (tapgemm.cu) re-reads a shifted 128-pixel activation tile and a weight tile from L2 for
// every tap, which makes those layers L2->SM bandwidth bound (measured ~70 B/clk/SM where 100-190 are needed).  Here a
// CTA owns T consecutive output rows of one image and
//   * loads each input row ONCE per 32-channel chunk as a "slab" of W + k - 1 pixels; the k horizontal taps are k
//     shared-memory descriptors into the same slab, shifted by one 128-byte pixel row each (the 128B swizzle is a
//     function of the shared-memory address, so a row-shifted start address addresses the same swizzled data);
//   * keeps T + 1.. slabs resident, so the k vertical taps of the T output rows reuse them (T + k - 1 row loads for
//     T*k row uses);
//   * loads each weight tile once per (tap, chunk) and multiplies it into T accumulators (T * N <= 512 TMEM columns).
// L2->SM traffic drops from ~190 to ~22 B/clk/SM for N = 64 and from ~100 to ~40 for N = 208.
//
// Roles as in tapgemm: warp 0 TMA producer, warp 1 MMA issuer (tcgen05.mma kind::tf32), warp 2 TMEM allocator,
// warps 4-7 epilogue (bias / addends / (Leaky)ReLU or activation-backward mask / tf32 rounding -> NHWC stores).
// Serves Conv2d forward and its stride-1 input gradient (ModificationLayer.py:101 and aten::convolution_backward).
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kRcMaxSlots = 16;
constexpr int kRowConvThreads = kConvThreads;   // warps 4..11 are the epilogue (two per TMEM lane quarter, alternating 16-column groups)

template <bool BF16>
__global__ void __launch_bounds__(kRowConvThreads, 1) rowconv_kernel(const __grid_constant__ RowConvParams P, int* status) {
  using Op = Opnd<BF16>;
  constexpr int CH = Op::kChunk;   // channels per 128-byte K chunk
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kRcMaxSlots];
  __shared__ __align__(8) uint64_t a_empty[kRcMaxSlots];
  __shared__ __align__(8) uint64_t b_full[kRcMaxSlots];
  __shared__ __align__(8) uint64_t b_empty[kRcMaxSlots];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRcMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kRowConvThreads - 128);
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

  const int T = P.T, k = P.k, kchunks = P.kchunks;
  const int AS = P.a_slots, BS = P.b_slots;
  const int nslab = T + k - 1;
  const uint32_t slab_bytes = (uint32_t)P.slab_bytes, b_bytes = (uint32_t)P.b_bytes;
  const uint32_t smem_a = smem_u32(smem), smem_b = smem_a + (uint32_t)AS * slab_bytes;
  const uint32_t af0 = smem_u32(&a_full[0]), ae0 = smem_u32(&a_empty[0]);
  const uint32_t bf0 = smem_u32(&b_full[0]), be0 = smem_u32(&b_empty[0]);
  const int nbuf = P.double_buf ? 2 : 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int as_ = 0, bs_ = 0;           // ring slots
      uint32_t aph = 0, bph = 0;      // ring phases
      bool ok = true;
      const uint32_t slab_tx = (uint32_t)(P.W + k - 1) * 128u;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        const int nt = tile % P.n_tiles;
        const int rt = (tile / P.n_tiles) % P.row_tiles;
        const int n = tile / (P.n_tiles * P.row_tiles);
        const int y_in0 = rt * T + P.dy0;   // input row of slab 0
        const int bn0 = nt * P.block_n;
        for (int c = 0; ok && c < kchunks; ++c) {
          int next_slab = 0;
          auto load_slab = [&]() -> bool {
            if (!mbar_wait_a(ae0 + 8u * as_, aph ^ 1u, ac, 21)) return false;
            mbar_arrive_expect_tx_a(af0 + 8u * as_, slab_tx);
            tma_load_4d_a(smem_a + (uint32_t)as_ * slab_bytes, &P.amap, af0 + 8u * as_, c * CH, P.dx0, y_in0 + next_slab, n);
            ++next_slab;
            if (++as_ == AS) { as_ = 0; aph ^= 1u; }
            return true;
          };
          for (int i = 0; ok && i < T; ++i) ok = load_slab();
          for (int r = 0; ok && r < k; ++r) {
            for (int j = 0; j < k; ++j) {
              if (!mbar_wait_a(be0 + 8u * bs_, bph ^ 1u, ac, 22)) { ok = false; break; }
              mbar_arrive_expect_tx_a(bf0 + 8u * bs_, (uint32_t)P.block_n * 128u);
              tma_load_3d_a(smem_b + (uint32_t)bs_ * b_bytes, &P.bmap, bf0 + 8u * bs_, c * CH, bn0, P.wtap[r * k + j]);
              if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
            }
            if (ok && next_slab < nslab) ok = load_slab();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int aw = 0;                      // slot of the next slab to wait for
      uint32_t awph = 0;
      int ar = 0;                      // slot of the next slab to release
      int bs_ = 0;
      uint32_t bph = 0;
      int buf = 0;
      uint32_t tph = 0;
      bool ok = true;
      const uint32_t dhi = desc_hi(1024, 2);
      const uint32_t idesc = Op::idesc(128, P.block_n, 0, 0);
      const uint32_t a_lo0 = desc_lo(smem_a, 16), b_lo0 = desc_lo(smem_b, 16);
      const uint32_t slab16 = slab_bytes >> 4, b16 = b_bytes >> 4;
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        if (!mbar_wait(&tempty_bar[buf], tph ^ 1u, ac, 23)) break;
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
        for (int c = 0; ok && c < kchunks; ++c) {
          const int nm = (c == kchunks - 1) ? P.last_mmas : 4;
          int slot_r = ar;             // slot of slab r (slabs are released in order, so slab 0 of this chunk = ar)
          for (int r = 0; ok && r < k; ++r) {
            // slabs r .. r+T-1 must have landed: T of them for r == 0, one more for each later r
            const int need = (r == 0) ? T : 1;
            for (int w = 0; w < need; ++w) {
              if (!mbar_wait_a(af0 + 8u * aw, awph, ac, 24)) { ok = false; break; }
              if (++aw == AS) { aw = 0; awph ^= 1u; }
            }
            if (!ok) break;
            tc_fence_after();
            for (int j = 0; j < k; ++j) {
              if (!mbar_wait_a(bf0 + 8u * bs_, bph, ac, 25)) { ok = false; break; }
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + (uint32_t)bs_ * b16;
              const uint32_t a_off = (uint32_t)P.dxoff[r * k + j] * 8u;  // pixel rows of 128 B = 8 x 16 B
              const uint32_t first = (c == 0 && r == 0 && j == 0) ? 0u : 1u;
              int slot = slot_r;
              for (int t = 0; t < T; ++t) {
                const uint32_t a_lo = a_lo0 + (uint32_t)slot * slab16 + a_off;
                const uint32_t d = d0 + (uint32_t)(t * P.block_n);
                if (nm == 4) {
                  Op::mma(d, desc_join(a_lo, dhi), desc_join(b_lo, dhi), idesc, first);
                  Op::mma(d, desc_join(a_lo + 2, dhi), desc_join(b_lo + 2, dhi), idesc, 1);
                  Op::mma(d, desc_join(a_lo + 4, dhi), desc_join(b_lo + 4, dhi), idesc, 1);
                  Op::mma(d, desc_join(a_lo + 6, dhi), desc_join(b_lo + 6, dhi), idesc, 1);
                } else {
                  for (int q = 0; q < nm; ++q)
                    Op::mma(d, desc_join(a_lo + 2 * q, dhi), desc_join(b_lo + 2 * q, dhi), idesc, q ? 1u : first);
                }
                if (++slot == AS) slot = 0;
              }
              tc_commit_a(be0 + 8u * bs_);
              if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
            }
            if (!ok) break;
            // release slab r (its last use was this tap row), or all remaining slabs after the last tap row
            const int nrel = (r < k - 1) ? 1 : T;
            for (int w = 0; w < nrel; ++w) {
              tc_commit_a(ae0 + 8u * ar);
              if (++ar == AS) ar = 0;
            }
            if (++slot_r == AS) slot_r = 0;
          }
        }
        if (!ok) break;
        tc_commit(&tfull_bar[buf]);
        if (++buf == nbuf) { buf = 0; tph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int x = q * 32 + lane;   // pixel column of this TMEM lane
    int buf = 0;
    uint32_t tph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int nt = tile % P.n_tiles;
      const int rt = (tile / P.n_tiles) % P.row_tiles;
      const int n = tile / (P.n_tiles * P.row_tiles);
      if (!mbar_wait(&tfull_bar[buf], tph, ac, 26)) break;
      tc_fence_after();
      const int col_base = nt * P.block_n;
      const EpiArgs E{P.bias, P.slopes, P.cout_valid, P.epilogue, P.round_tf32, P.vec_ok, P.slope};
      for (int t = 0; t < T; ++t) {
        const int y = rt * T + t;
        const bool valid = (y < P.H) && (x < P.W);
        const long long pix = (long long)n * P.out.sn + (long long)y * P.out.sh + (long long)x * P.out.sw;
        float* po = P.out.ptr + pix;
        const float* p1 = P.add1.ptr ? P.add1.ptr + (long long)n * P.add1.sn + (long long)y * P.add1.sh + (long long)x * P.add1.sw : nullptr;
        const float* p2 = P.add2.ptr ? P.add2.ptr + (long long)n * P.add2.sn + (long long)y * P.add2.sh + (long long)x * P.add2.sw : nullptr;
        const float* pm = P.mask.ptr ? P.mask.ptr + (long long)n * P.mask.sn + (long long)y * P.mask.sh + (long long)x * P.mask.sw : nullptr;
        uint16_t* po16 = P.out16.ptr ? P.out16.ptr + (long long)n * P.out16.sn + (long long)y * P.out16.sh + (long long)x * P.out16.sw : nullptr;
        if (!P.out.ptr) po = nullptr;
        const uint32_t t_addr = tmem_base + (uint32_t)(buf * 256 + t * P.block_n) + ((uint32_t)(q * 32) << 16);
        for (int c0 = ((warp - 4) >> 2) * 16; c0 < P.block_n; c0 += 16 * kEpiPerQuarter) {
          uint32_t r[16];
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (valid) epilogue_store16(r, E, col_base + c0, po, p1, p2, pm, po16);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      if (++buf == nbuf) { buf = 0; tph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template __global__ void rowconv_kernel<false>(const __grid_constant__ RowConvParams, int*);
template __global__ void rowconv_kernel<true>(const __grid_constant__ RowConvParams, int*);

}  // namespace tpg
