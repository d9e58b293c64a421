// Row-tile convolution with N-STACKED vertical taps, for the narrow (N <= 128 output channels) full-resolution layers:
// 64->64 k7, 75->75 k7, 206->64 k5, 64->64 k3, 64->32, 32->3 ... all stride 1 at W = 128.
//
// Why: an SS-mode tcgen05.mma (M = 128, K = 8 tf32) reads 4 KB of A and N*32 B of B from shared memory per issue, so for
// N = 64 the shared-memory read (48 clk) and not the tensor pipe (32 clk) sets the pace - the plain row-tile kernel
// (rowconv.cu) sits at ~40-45 % of the tensor peak on these layers with no memory traffic left to remove.  Here the SAME
// A operand (an input-row slab shifted to horizontal tap j) is multiplied into ALL output rows it contributes to in one
// instruction: input row s feeds output rows t = s - r through the vertical taps r, so B is the stack
//   [ W(r = s - t_lo, j) ; W(r - 1, j) ; ... ; W(s - t_hi, j) ]      (N' = (t_hi - t_lo + 1) * N <= 256 rows)
// and D is the adjacent accumulators of rows t_lo .. t_hi in TMEM.  A is then read once per up to T output rows and the
// instruction is tensor-pipe bound (N' = 256: 128 clk of math against 96 clk of operand reads).
//
// To keep all T + k - 1 slabs of a channel chunk resident next to a whole k-tap weight stack the K chunk is 16 channels
// (64-byte rows, 64B swizzle; two K = 8 MMAs per chunk):
//   slab  = {16 ch, W + k - 1 px} = ~9 KB, ring of T + k - 1 + 2 slots; horizontal tap j = descriptor start + j * 64 B
//   stack = k boxes {16 ch, N rows} of the packed weights, taps (r = k-1 .. 0, j), k * N * 64 B, 3-stage ring
// Roles, barriers, epilogue and bounded waits as in rowconv.cu.  Replaces aten::convolution / convolution_backward
// (input gradient) for ModificationLayer.py:101.
#include "common.cuh"
#include "kparams.h"

namespace tpg {

constexpr int kRsMaxSlots = 16;

template <bool BF16>
__global__ void __launch_bounds__(kConvThreads, 1) rowstack_kernel(const __grid_constant__ RowStackParams P, int* status) {
  using Op = Opnd<BF16>;
  constexpr int CH = Op::kChunk / 2;   // channels per 64-byte K chunk (16 tf32 / 32 bf16)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kRsMaxSlots];
  __shared__ __align__(8) uint64_t a_empty[kRsMaxSlots];
  __shared__ __align__(8) uint64_t b_full[kMaxStages];
  __shared__ __align__(8) uint64_t b_empty[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_flag;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRsMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxStages; ++i) {
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

  const int T = P.T, k = P.k, kchunks = P.kchunks, N = P.block_n;
  const int AS = P.a_slots, BS = P.b_slots;
  const int nslab = T + k - 1;
  const uint32_t slab_bytes = (uint32_t)P.slab_bytes, wb_bytes = (uint32_t)P.wb_bytes;
  const uint32_t smem_a = smem_u32(smem), smem_b = smem_a + (uint32_t)AS * slab_bytes;
  const uint32_t af0 = smem_u32(&a_full[0]), ae0 = smem_u32(&a_empty[0]);
  const uint32_t bf0 = smem_u32(&b_full[0]), be0 = smem_u32(&b_empty[0]);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      bool ok = true;
      const uint32_t slab_tx = (uint32_t)(P.W + k - 1) * 64u;
      const uint32_t box_bytes = (uint32_t)N * 64u;
      auto load_stack = [&](int c, int j, int bn0) -> bool {
        if (!mbar_wait_a(be0 + 8u * bs_, bph ^ 1u, ac, 31)) return false;
        mbar_arrive_expect_tx_a(bf0 + 8u * bs_, box_bytes * (uint32_t)k);
        const uint32_t dst = smem_b + (uint32_t)bs_ * wb_bytes;
        for (int i = 0; i < k; ++i)   // stack order: vertical tap r = k-1 first
          tma_load_3d_a(dst + (uint32_t)i * box_bytes, &P.bmap, bf0 + 8u * bs_, c * CH, bn0, P.wtap[(k - 1 - i) * k + j]);
        if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
        return true;
      };
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        const int rt = tile % P.row_tiles;
        const int n = tile / P.row_tiles;
        const int y_in0 = rt * T + P.dy0;
        for (int c = 0; ok && c < kchunks; ++c) {
          ok = load_stack(c, 0, 0);
          for (int s = 0; ok && s < nslab; ++s) {
            if (!mbar_wait_a(ae0 + 8u * as_, aph ^ 1u, ac, 32)) { ok = false; break; }
            mbar_arrive_expect_tx_a(af0 + 8u * as_, slab_tx);
            tma_load_4d_a(smem_a + (uint32_t)as_ * slab_bytes, &P.amap, af0 + 8u * as_, c * CH, P.dx0, y_in0 + s, n);
            if (++as_ == AS) { as_ = 0; aph ^= 1u; }
          }
          for (int j = 1; ok && j < k; ++j) ok = load_stack(c, j, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int a0 = 0;                      // ring slot of slab 0 of the current chunk
      uint32_t a0ph = 0;               // ring phase of that slot
      int bs_ = 0;
      uint32_t bph = 0;
      int buf = 0;
      uint32_t tph = 0;
      bool ok = true;
      const uint32_t dhi = desc_hi(512, 4);          // K-major, 64B swizzle: 8-row groups of 512 B
      const uint32_t a_lo0 = desc_lo(smem_a, 16), b_lo0 = desc_lo(smem_b, 16);
      const uint32_t slab16 = slab_bytes >> 4, wb16 = wb_bytes >> 4, nrow16 = (uint32_t)N * 4u;   // N rows of 64 B, in 16 B
      // The issuing thread is the critical resource (one thread feeds the tensor pipe): everything that depends only on
      // the slab index s - instruction descriptor (N' = rows * N), B offset inside the weight stack, accumulator column
      // offset - comes from host-built tables in the parameter bank, and all ring addresses advance arithmetically.
      const uint32_t ring16 = (uint32_t)AS * slab16;
      uint32_t a_ring = 0;             // offset (16 B units) of slab 0 of the current chunk inside the slab ring
      for (int tile = blockIdx.x; ok && tile < P.total_tiles; tile += gridDim.x) {
        if (!mbar_wait(&tempty_bar[buf], tph ^ 1u, ac, 33)) break;
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256);
        for (int c = 0; ok && c < kchunks; ++c) {
          const bool two = (c != kchunks - 1) || (P.last_mmas == 2);
          for (int j = 0; ok && j < k; ++j) {
            if (!mbar_wait_a(bf0 + 8u * bs_, bph, ac, 34)) { ok = false; break; }
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + (uint32_t)bs_ * wb16;
            uint32_t a_off = a_ring;
            if (j == 0) {
              // first pass over this chunk's slabs: they may still be in flight; on the first chunk of a tile the first
              // contribution to output row t = s overwrites its accumulator (split off as its own N-wide MMA)
              int slot = a0;
              uint32_t sph = a0ph;
              for (int s = 0; s < nslab; ++s) {
                if (!mbar_wait_a(af0 + 8u * slot, sph, ac, 35)) { ok = false; break; }
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + a_off;
                const bool fresh = (c == 0) && (s <= T - 1);
                const int t_lo = max(0, s - k + 1);
                const int t_acc_hi = fresh ? s - 1 : min(T - 1, s);
                if (t_acc_hi >= t_lo) {
                  const uint32_t idesc = Op::idesc(128, (t_acc_hi - t_lo + 1) * N, 0, 0);
                  const uint32_t bb = b_lo + P.s_boff[s];
                  const uint32_t d = d0 + P.s_doff[s];
                  Op::mma(d, desc_join(a_lo, dhi), desc_join(bb, dhi), idesc, 1);
                  if (two) Op::mma(d, desc_join(a_lo + 2, dhi), desc_join(bb + 2, dhi), idesc, 1);
                }
                if (fresh) {
                  const uint32_t idesc = Op::idesc(128, N, 0, 0);
                  const uint32_t bb = b_lo + (uint32_t)(k - 1) * nrow16;   // vertical tap r = 0
                  const uint32_t d = d0 + (uint32_t)(s * N);
                  Op::mma(d, desc_join(a_lo, dhi), desc_join(bb, dhi), idesc, 0);
                  if (two) Op::mma(d, desc_join(a_lo + 2, dhi), desc_join(bb + 2, dhi), idesc, 1);
                }
                a_off += slab16;
                if (a_off >= ring16) a_off -= ring16;
                if (++slot == AS) { slot = 0; sph ^= 1u; }
              }
              if (!ok) break;
            } else {
              const uint32_t a_j = a_lo0 + (uint32_t)j * 4u;     // tap j: + j pixel rows of 64 B
              const bool last = (j == k - 1);
              uint32_t ae = ae0 + 8u * (uint32_t)a0;
              const uint32_t ae_end = ae0 + 8u * (uint32_t)AS;
              for (int s = 0; s < nslab; ++s) {
                const uint32_t a_lo = a_j + a_off;
                const uint32_t idesc = P.s_idesc[s];
                const uint32_t bb = b_lo + P.s_boff[s];
                const uint32_t d = d0 + P.s_doff[s];
                Op::mma(d, desc_join(a_lo, dhi), desc_join(bb, dhi), idesc, 1);
                if (two) Op::mma(d, desc_join(a_lo + 2, dhi), desc_join(bb + 2, dhi), idesc, 1);
                if (last) tc_commit_a(ae);                       // last use of this slab
                a_off += slab16;
                if (a_off >= ring16) a_off -= ring16;
                ae += 8;
                if (ae == ae_end) ae = ae0;
              }
              if (last) {
                a_ring = a_off;
                a0 += nslab;
                if (a0 >= AS) { a0 -= AS; a0ph ^= 1u; }
              }
            }
            tc_commit_a(be0 + 8u * bs_);
            if (++bs_ == BS) { bs_ = 0; bph ^= 1u; }
          }
        }
        if (!ok) break;
        tc_commit(&tfull_bar[buf]);
        if (++buf == 2) { buf = 0; tph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (as rowconv.cu)
    const int q = warp & 3;
    const int x = q * 32 + lane;
    int buf = 0;
    uint32_t tph = 0;
    const EpiArgs E{P.bias, P.slopes, P.cout_valid, P.epilogue, P.round_tf32, P.vec_ok, P.slope};
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int rt = tile % P.row_tiles;
      const int n = tile / P.row_tiles;
      if (!mbar_wait(&tfull_bar[buf], tph, ac, 36)) break;
      tc_fence_after();
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
        const uint32_t t_addr = tmem_base + (uint32_t)(buf * 256 + t * N) + ((uint32_t)(q * 32) << 16);
        for (int c0 = ((warp - 4) >> 2) * 16; c0 < N; c0 += 16 * kEpiPerQuarter) {
          uint32_t r[16];
          tmem_ld16(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (valid) epilogue_store16(r, E, c0, po, p1, p2, pm, po16);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      if (++buf == 2) { buf = 0; tph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template __global__ void rowstack_kernel<false>(const __grid_constant__ RowStackParams, int*);
template __global__ void rowstack_kernel<true>(const __grid_constant__ RowStackParams, int*);

}  // namespace tpg
