// Device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX).
// Everything here is Blackwell-only; the library is compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a converged warp (ptxas recognises elect.sync and treats the guarded region as single-threaded, so
// uniform-datapath instructions - UTMALDG, UTCHMMA, UTCBAR - are emitted without per-instruction election loops).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// The tensor-core kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may
// become resident (on SMs this grid's CTAs have already left) and run their prologue - barrier init, TMEM allocation -
// while this grid is still finishing.  pdl_wait() returns once every prerequisite grid has COMPLETED and its memory
// operations are visible; nothing before it may touch global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}

// Bounded wait: a kernel of this library can never hang the GPU.  If a barrier does not flip within
// ~2^31 cycles the waiter raises the CTA-wide abort flag (shared) and records a code in the
// host-visible status word; every other waiter sees the flag and leaves its loop too.
struct AbortCtl {
  volatile int* smem_flag;  // shared, one per CTA
  int* status;              // global (host-mapped) status word, may be null
};
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, const AbortCtl& ac, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      if (*ac.smem_flag) return false;
      if (clock64() - t0 > (1ll << 31)) {
        *ac.smem_flag = 1;
        if (ac.status) {
          atomicCAS(ac.status, 0, code | (int(blockIdx.x) << 8));
          __threadfence_system();
        }
        return false;
      }
    }
  }
  return true;
}

// ---- address-based variants: the hot loops compute shared addresses once and advance them arithmetically (taking
// &bar[i] per use makes ptxas re-derive the shared-window address - S2UR CgaCtaId + ULEA - on every wait/arrive).
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
static __device__ __noinline__ bool mbar_wait_slow_a(uint32_t bar, uint32_t parity, volatile int* smem_flag, int* status, int code) {
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_a(bar, parity)) {
    if ((++spins & 255u) == 0) {
      if (*smem_flag) return false;
      if (clock64() - t0 > (1ll << 31)) {
        *smem_flag = 1;
        if (status) {
          atomicCAS(status, 0, code | (int(blockIdx.x) << 8));
          __threadfence_system();
        }
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ bool mbar_wait_a(uint32_t bar, uint32_t parity, const AbortCtl& ac, int code) {
  if (mbar_try_wait_a(bar, parity)) return true;
  if (mbar_try_wait_a(bar, parity)) return true;
  return mbar_wait_slow_a(bar, parity, ac.smem_flag, ac.status, code);
}
__device__ __forceinline__ void tc_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                              int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma of M = 256: each CTA holds its 128 rows of A and of the
// accumulator and HALF of B (N/2 rows) - the tensor cores exchange the B halves, so a CTA's shared memory receives and
// serves half the weight bytes per MMA.  Protocol (the one CUTLASS's 2-SM kernels use):
//   * both CTAs issue their own TMA loads, but every load signals the LEADER's (cluster rank 0) full barrier: the barrier
//     operand of cp.async.bulk.tensor.cta_group::2 is the local barrier's address with the cluster-rank bit cleared;
//   * the leader's elected thread issues the MMAs and commits with .multicast::cluster to the barrier at the same offset
//     in BOTH CTAs (stage-empty and accumulator-full);
//   * the epilogue warps of both CTAs release an accumulator by arriving on the LEADER's barrier (mapa + remote arrive).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address -> the same offset in cluster rank 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one arrival on the barrier at this offset in both CTAs of the pair once every MMA issued so far has retired
__device__ __forceinline__ void tc_commit2_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrival on the barrier at this offset in cluster rank `rank`
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
// wait whose acquire covers arrivals made by the peer CTA
__device__ __forceinline__ uint32_t mbar_try_wait_cluster_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
static __device__ __noinline__ bool mbar_wait_cluster_a(uint32_t bar, uint32_t parity, volatile int* smem_flag, int* status, int code) {
  if (mbar_try_wait_cluster_a(bar, parity)) return true;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster_a(bar, parity)) {
    if ((++spins & 255u) == 0) {
      if (*smem_flag) return false;
      if (clock64() - t0 > (1ll << 31)) {
        *smem_flag = 1;
        if (status) {
          atomicCAS(status, 0, code | (int(blockIdx.x) << 8));
          __threadfence_system();
        }
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival once every tcgen05.mma issued so far by this thread retired.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate. One thread issues for the CTA.
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 inputs (kind::f16), fp32 accumulate; K = 16 elements (32 bytes) per instruction.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Operand-type traits of the tensor-core kernels.  Both operand types move through shared memory in 128-byte rows and
// advance 32 bytes per MMA K step, so the kernels are identical in BYTES; what changes is the number of channels per row
// (32 tf32 / 64 bf16), the elements per K step (8 / 16), the MMA kind and the instruction-descriptor format bits.
template <bool BF16>
struct Opnd {
  static constexpr int kChunk = BF16 ? 64 : 32;   // channels per 128-byte shared-memory row
  static constexpr int kMmaK = BF16 ? 16 : 8;     // elements per MMA K step
  static __device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    if constexpr (BF16) mma_bf16_ss(tmem_d, desc_a, desc_b, idesc, accumulate);
    else mma_tf32_ss(tmem_d, desc_a, desc_b, idesc, accumulate);
  }
  // M = 256 over a CTA pair (issued by the leader CTA only)
  static __device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    if constexpr (BF16) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
          : "memory");
    }
  }
  // instruction descriptor: fp32 accumulate (c_format 1), a/b format 2 = TF32 or 1 = BF16, major bits, N >> 3, M >> 4
  static __host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((BF16 ? 1u : 2u) << 7) | ((BF16 ? 1u : 2u) << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  }
};
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format, version 1), 128-byte swizzle.
//   K-major  operand: rows of 128 B (32 tf32) stacked; SBO = byte distance between 8-row groups.
//   MN-major tf32 operand: only the 32B-atom 128B swizzle exists (layout type 1, TMA SWIZZLE_128B_ATOM_32B):
//                     atom = 4 K-rows x 128 B; LBO = distance between 32-element MN chunks, SBO between 4-row K atoms.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B (16B atoms), 1 = SWIZZLE_128B_BASE32B (32B atoms)
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Descriptor halves: hi word is constant per layout, lo word = (addr >> 4) | LBO field; advancing the operand by `bytes`
// inside the tile is an add of bytes >> 4 on the lo word.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}


// ---------------------------------------------------------------- fused conv epilogue (shared by tapgemm / rowconv)
// One 16-column group of one accumulator row: + bias, + up to two addends, (Leaky)ReLU or activation-backward mask with
// a scalar or per-channel slope, optional tf32 rounding, store.  All global loads of the group are issued before the first
// store (an addend may legitimately alias the output - in-place accumulation - so the pointers are not __restrict__; with
// loads interleaved between stores they serialise and the epilogue becomes a chain of ~8 dependent memory round trips).
struct EpiArgs {
  const float* bias;
  const float* slopes;
  int cout_valid, epilogue, round_tf32, vec_ok;
  float slope;
};
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
__device__ __forceinline__ uint16_t to_bf16(float x) {
  uint16_t r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
// po: fp32 destination of the pixel (may be null), po16: bf16 copy of the same values (may be null; bf16 operand storage
// for the tensor-core consumers of this tensor).  Both receive the value BEFORE any rounding of the other.
__device__ __forceinline__ void epilogue_store16(const uint32_t (&r)[16], const EpiArgs& E, int col0, float* po,
                                                 const float* p1, const float* p2, const float* pm,
                                                 uint16_t* po16 = nullptr) {
  const int nv16 = E.cout_valid - col0;
  if (nv16 <= 0) return;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  if (E.vec_ok && nv16 >= 16) {
    float4 a1[4], a2[4], mk[4];
    if (p1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a1[j] = *reinterpret_cast<const float4*>(p1 + col0 + 4 * j);
    }
    if (p2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a2[j] = *reinterpret_cast<const float4*>(p2 + col0 + 4 * j);
    }
    if (E.epilogue == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mk[j] = *reinterpret_cast<const float4*>(pm + col0 + 4 * j);
    }
    if (E.bias) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += __ldg(E.bias + col0 + j);
    }
    if (p1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[4 * j] += a1[j].x; v[4 * j + 1] += a1[j].y; v[4 * j + 2] += a1[j].z; v[4 * j + 3] += a1[j].w; }
    }
    if (p2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[4 * j] += a2[j].x; v[4 * j + 1] += a2[j].y; v[4 * j + 2] += a2[j].z; v[4 * j + 3] += a2[j].w; }
    }
    if (E.epilogue == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * E.slope;
    } else if (E.epilogue == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float m4[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float s = E.slopes ? __ldg(E.slopes + col0 + 4 * j + i) : E.slope;
          v[4 * j + i] = m4[i] > 0.f ? v[4 * j + i] : v[4 * j + i] * s;
        }
      }
    }
    if (po16) {   // the view starts on a 4-channel boundary: 8-byte stores always, 16-byte when aligned
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      uint16_t* d = po16 + col0;
      if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(d + 8) = make_uint4(h[4], h[5], h[6], h[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint2*>(d + 4 * j) = make_uint2(h[2 * j], h[2 * j + 1]);
      }
    }
    if (po) {
      if (E.round_tf32) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(po + col0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    return;
  }
  // ragged tail / unaligned views: element-wise
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < nv16) {
      float x = v[j];
      const int col = col0 + j;
      if (E.bias) x += __ldg(E.bias + col);
      if (p1) x += p1[col];
      if (p2) x += p2[col];
      if (E.epilogue == 1) {
        x = x > 0.f ? x : x * E.slope;
      } else if (E.epilogue == 2) {
        const float s = E.slopes ? __ldg(E.slopes + col) : E.slope;
        x = pm[col] > 0.f ? x : x * s;
      }
      if (po16) po16[col] = to_bf16(x);
      if (E.round_tf32) x = round_tf32(x);
      if (po) po[col] = x;
    }
  }
}

}  // namespace tpg
