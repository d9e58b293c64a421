// Kernel-parameter structures shared by the host planner (api.cu) and the tcgen05 kernels.
// They are passed by value as __grid_constant__ kernel parameters, so the TMA descriptors they hold live in
// the parameter bank and can be handed to cp.async.bulk.tensor directly.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace tpg {

constexpr int kMaxTaps = 64;    // 7x7 conv = 49, 8x8 deconv-as-GEMM is a 1-tap problem
constexpr int kMaxPhases = 16;  // stride-4 deconv: 4x4 output phases
constexpr int kMaxPlanes = 16;  // stride-4 gather: 4x4 input parity planes
constexpr int kMaxGroups = 4;   // the four local pathways
constexpr int kMaxStages = 8;
constexpr int kTapMaxStages = 12;  // tapgemm: single-chunk stages of a paired launch are 20-30 KB
// Epilogue warps per TMEM lane quarter in the three tensor-core kernels (warps 4 .. 4 + 4*kEpiPerQuarter - 1); the warps
// of a quarter take the 16-column groups of an accumulator round-robin.  The epilogue is latency-bound on its global
// loads/stores, so more warps = more of them in flight.
#ifndef TPG_EPI_PER_QUARTER
#define TPG_EPI_PER_QUARTER 3
#endif
constexpr int kEpiPerQuarter = TPG_EPI_PER_QUARTER;
constexpr int kConvThreads = 128 + 128 * kEpiPerQuarter;

struct TapDesc {
  int8_t plane;  // which A tensor map (parity plane)
  int8_t dy, dx; // box origin offset in the plane
  uint8_t wtap;  // weight tap slice in the packed tensor
};

struct DevView {
  float* ptr;
  long long sn, sh, sw;
};

struct DevView16 {   // bf16 copy of an output view (operand storage for tensor-core consumers); ptr null = not written
  uint16_t* ptr;
  long long sn, sh, sw;
};

struct PhaseDesc {
  short tap_begin, tap_count;
  short oy, ox;  // output pixel = (y*out_sy + oy, x*out_sx + ox)
};

// ---------------------------------------------------------------- forward-type multi-tap GEMM
struct TapGemmGroup {
  CUtensorMap amap[kMaxPlanes];  // 4D {C, W, H, N} fp32, box {32, bw, bh, bn}, 128B swizzle
  CUtensorMap bmap;              // 3D {k_pad, rows_pad, taps+1}, box {32, block_n, 1}
  CUtensorMap bmap_lo;           // one-launch 3xTF32: the residual packing (taps whose wtap has bit 7 set read it)
  int Hm, Wm, Nimg;              // tile space (pixels enumerated by the M dimension)
  int bw, bh, bn;                // pixels of one 128-row tile = bw*bh*bn (<= 128)
  int tiles_h, m_tiles;          // ceil(Hm/bh), tiles_h * ceil(Nimg/bn)
  int n_tiles, block_n;
  int kchunks, last_mmas;        // K = C of the A operand, in 128-byte chunks (32 tf32 / 64 bf16 channels); MMAs in the last chunk
  int ksplit, kc_per, kt_per;    // split-K (GEMM-like launches with few tiles and a long reduction: the Linear layers): a tile
                                 // is cut into ksplit ranges of kt_per taps (kt_per > 0) or of kc_per chunks of its single tap.
  float* split_ws;               // Partial accumulators [tile][range][128 rows][block_n] in a library workspace; the CTA that
  int* split_cnt;                // arrives last at the tile's counter adds them IN RANGE ORDER (deterministic, no float
                                 // atomics), runs the normal fused epilogue and resets the counter
  int n_phases;
  int tile_begin, tile_count;    // this group's slice of the persistent tile list
  PhaseDesc phase[kMaxPhases];
  TapDesc taps[kMaxTaps];
  // epilogue
  DevView out, add1, add2, mask;
  DevView16 out16;
  const float* bias;
  const float* slopes;
  int out_sy, out_sx, Hout, Wout;
  int cout_valid;
  int epilogue;
  float slope;
  int round_tf32;
  int vec_ok;
};

struct TapGemmParams {
  int ngroups;
  int total_tiles;
  int stages;
  int b_stage_bytes;  // max over groups of block_n*128
  int kst;            // 32-float K chunks per pipeline stage (1 or 2)
  TapGemmGroup g[kMaxGroups];
};

struct TapGemmParams1 {  // single-group variant (keeps the parameter block small for the common launch)
  int ngroups;
  int total_tiles;
  int stages;
  int b_stage_bytes;
  int kst;
  TapGemmGroup g[1];
};

// ---------------------------------------------------------------- row-tile convolution (W = 128, stride 1)
struct RowConvParams {
  CUtensorMap amap;        // 4D {C, W, H, N} fp32, box {32, W + k - 1, 1, 1}, 128B swizzle (an input-row slab)
  CUtensorMap bmap;        // 3D {k_pad, rows_pad, taps+1}, box {32, block_n, 1}
  int H, W, Nimg, T, k;    // T output rows per tile; k x k tap grid
  int dy0, dx0;            // input pixel of grid tap (0,0) relative to the output pixel
  int kchunks, last_mmas;
  int n_tiles, block_n, row_tiles, total_tiles;
  int a_slots, b_slots, slab_bytes, b_bytes, double_buf;
  unsigned char wtap[kMaxTaps];   // weight tap slice for grid position r*k + j
  unsigned char dxoff[kMaxTaps];  // slab pixel-row offset (j) for grid position r*k + j
  DevView out, add1, add2, mask;
  DevView16 out16;
  const float* bias;
  const float* slopes;
  int cout_valid, epilogue, round_tf32, vec_ok;
  float slope;
};

// ---------------------------------------------------------------- row-tile convolution, N-stacked vertical taps (N <= 128)
struct RowStackParams {
  CUtensorMap amap;        // 4D {C, W, H, N} fp32, box {16, W + k - 1, 1, 1}, 64B swizzle (an input-row slab, 16 channels)
  CUtensorMap bmap;        // 3D {k_pad, rows_pad, taps+1}, box {16, block_n, 1}, 64B swizzle
  int H, W, Nimg, T, k;    // T output rows per tile (T * block_n <= 256 TMEM columns, double-buffered)
  int dy0, dx0;            // input pixel of grid tap (0,0) relative to the output pixel
  int kchunks, last_mmas;  // 16-channel K chunks; K = 8 MMAs in the last one (1 or 2)
  int block_n, row_tiles, total_tiles;
  int a_slots, b_slots, slab_bytes, wb_bytes;
  unsigned char wtap[kMaxTaps];   // weight tap slice for grid position r*k + j
  // per input-row slab s of a tile (it feeds output rows t_lo(s) .. t_hi(s)): instruction descriptor with
  // N' = (t_hi - t_lo + 1) * block_n, offset of vertical tap r = s - t_lo inside the weight stack (16 B units), first
  // accumulator column t_lo * block_n
  uint32_t s_idesc[16], s_boff[16], s_doff[16];
  DevView out, add1, add2, mask;
  DevView16 out16;
  const float* bias;
  const float* slopes;
  int cout_valid, epilogue, round_tf32, vec_ok;
  float slope;
};

// ---------------------------------------------------------------- flat-slab convolution (small maps, stride 1, "same" padding)
// A unit = T consecutive 128-row M tiles over the PADDED-PITCH pixel enumeration of `ur` output rows (or `bn` whole images):
// flat row m <-> image m / ipitch, row (m % ipitch) / Wp, column (m % ipitch) % Wp, with Wp = W + k - 1.  One TMA box
// {chunk, Wp, R, bn} (R = ur + k - 1, zero padding = out-of-bounds fill) lands in shared memory in exactly that enumeration,
// so tap (r, j) of EVERY pixel is the same slab read r * Wp + j rows further down: the k*k taps share one activation load,
// and the T tiles of the unit share every weight tile.
struct FlatGroup {
  CUtensorMap amap;              // 4D {C, W, H, N}, box {chunk, Wp, R, bn}, 128B swizzle
  CUtensorMap bmap;              // 3D {k_pad, rows_pad, taps+1}, box {chunk, block_n, 1}
  int H, W, Nimg;
  int Wp, R, bn, ur, ipitch;     // ipitch = R * Wp flat rows per image of a unit
  int T;                         // 128-row tiles per unit (T * block_n <= 256 TMEM columns, double-buffered)
  int units_h;                   // ceil(H / ur); units = units_h * ceil(Nimg / bn)
  int n_tiles, block_n;
  int kchunks, last_mmas;
  int tile_begin, tile_count;    // units * n_tiles
  int slab_tx;                   // bn * R * Wp * 128 bytes per slab load
  DevView out, add1, add2, mask;
  DevView16 out16;
  const float* bias;
  const float* slopes;
  int cout_valid, epilogue, round_tf32, vec_ok;
  float slope;
};

struct FlatConvParams {
  int ngroups, total_tiles;
  int k, dy0, dx0;               // shared by the groups of a launch (same layer type)
  int a_slots, b_slots, slab_bytes, b_bytes;
  unsigned char wtap[kMaxTaps];  // weight tap slice for grid position r*k + j
  FlatGroup g[kMaxGroups];
};

// ---------------------------------------------------------------- weight-gradient GEMM
// D[m = channel of P][n = channel of Q] (per tap) = sum over pixels P[pix, m] * Qtap[pix, n]
struct WgradGroup {
  CUtensorMap pmap;              // unshifted tensor, 4D {C, W, H, N}, box {32, bw, bh, bn}
  CUtensorMap qmap[kMaxPlanes];  // shifted tensor planes, same box
  CUtensorMap qslab;             // slab mode: box {32, bw + tpu - 1, 1, 1} of the shifted tensor (one row segment + halo)
  CUtensorMap pzero;             // box {32, kp - bw*bh*bn, 1, 1} of the unshifted tensor, only ever fetched fully out of bounds:
                                 // zero-fills the K rows a pixel box leaves unwritten (launches that mix box shapes)
  int Hp, Wp, Nimg;
  int bw, bh, bn, kp;            // kp = K rows per stage (pixels rounded up to 8)
  int p_rows, q_rows;            // pixel rows one P box / one Q slab box writes (multi-row slab boxes include the pad columns)
  int zero_ring;                 // the kernel must clear the ring before the first load (rows no box ever writes are read)
  int tiles_w, tiles_h, chunks;  // pixel boxes: tiles_w * tiles_h * ceil(Nimg/bn)
  int m_tiles, n_tiles, block_n; // M tile = 128 P-channels, N tile = block_n columns
  // A unit covers `mpu` consecutive M tiles (one accumulator each; they share every Q load) and `tpu` consecutive taps
  // (their Q chunks sit side by side in the N dimension; they share every P load).  mpu > 1 and tpu > 1 are exclusive.
  int mpu, mt_groups;            // M tiles per unit, ceil(m_tiles / mpu)
  int tpu, tap_groups, ncpt;     // taps per unit, ceil(ntaps / tpu), 32-channel Q chunks per tap (tpu > 1: block_n = tpu*ncpt*32)
  int nbuf;                      // TMEM accumulator buffers (2 when the unit's accumulators fit 256 columns)
  // Slab mode (stride 1, row-segment boxes): the taps of a unit are `tpu` horizontally consecutive taps of ONE kernel row;
  // Q is loaded once per chunk as a slab of bw + tpu - 1 pixels and tap t reads it shifted by t pixel rows (128 B), all
  // taps in one MMA per 32-channel chunk (N = taps * 32, leading byte offset 128).  Accumulator columns:
  // ((mi * ncpt + chunk) * tpu + tap) * 32 + channel.
  int slab, q_chunk_bytes;
  int ntaps;
  int work_begin, tiles;         // tiles = tap_groups * mt_groups * n_tiles; work = tiles * chunks, prefix over groups
  int kb_chunks, nkb;            // the work list is ordered (K block of kb_chunks chunks, tile, chunk); nkb = ceil(chunks / kb_chunks)
  TapDesc taps[kMaxTaps];
  float* dw;                     // [taps+1][rows_pad][k_pad]
  int rows_pad, k_pad;
  int transpose_out;             // 0: dw[tap][m][n]   1: dw[tap][n][m]
  int m_valid, n_valid;          // bounds in the dw tensor (rows_pad/k_pad by orientation)
  int accumulate;                // 1: dW += (always atomics); 0: dW = (plain stores when the K range is not split)
  // im2col-by-TMA mode (inputs of <= 4 channels, e.g. the RGB layers): Q is a zero-padded copy of x seen through tensor maps
  // whose pixel stride (16 B x stride) is smaller than their 128-byte inner extent - row p of the operand is the window of 8
  // consecutive pixels x 4 channels that starts at input pixel stride*p, i.e. ALL horizontal taps of a kernel row in one
  // 32-lane chunk.  The taps of a unit are the k VERTICAL taps (one map each, plane = r); accumulator column
  // r*32 + j*4 + c is dW[tap (r, j)][m][c].  im2col_k = k (0 = off).
  int im2col_k, im2col_cin;
};

struct WgradParams {
  int ngroups;
  int total_work;                // sum over groups of tiles * chunks
  int stages;
  int a_stage_bytes, b_stage_bytes;
  int ring_bytes;                // stages * (a + b) + slack read (never written) by M = 128 MMAs over narrow P tensors
  int need_zero;                 // some group's pixel box leaves K rows unwritten: zero the ring first
  int nbuf;
  int whole_tiles;               // deterministic mode: a CTA owns whole output tiles (no split reduction, no float atomics race)
  int zero_tail;                 // groups with different box shapes share the ring: every stage zero-fills the unwritten K rows of P
  WgradGroup g[kMaxGroups];
};
struct WgradParams1 {
  int ngroups;
  int total_work;
  int stages;
  int a_stage_bytes, b_stage_bytes;
  int ring_bytes;
  int need_zero;
  int nbuf;
  int whole_tiles;
  int zero_tail;
  WgradGroup g[1];
};

}  // namespace tpg
