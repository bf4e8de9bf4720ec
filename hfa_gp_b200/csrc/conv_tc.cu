// Tensor-core implicit-GEMM convolution for sm_100a: TMA (cp.async.bulk.tensor, SWIZZLE_128B) feeds a
// 3-stage shared-memory ring, one elected thread issues tcgen05.mma with the fp32 accumulator in TMEM, and
// four epilogue warps read it back with tcgen05.ld and apply the fused StyleGAN2 / encoder epilogue.
//
// Precision: fp32-class results on the bf16 tensor pipe.  Every fp32 operand is stored as a pair of bf16
// tensors (hi = bf16(x), lo = bf16(x - hi), ~16 mantissa bits together) and each K step issues three MMAs
// into the same accumulator:  hi*hi + lo*hi + hi*lo  (the dropped lo*lo term is ~2^-18 relative).
//
// GEMM view (channels-last):  M = a BHxBW patch of 128 output pixels, N = BN output channels,
// K = taps x cin in chunks of 64 channels.  The A tile of tap (dy,dx) is the TMA box {64ch, BW, BH, 1} at
// (c0, x0+dx, y0+dy, n): out-of-range coordinates are zero-filled by the TMA unit, which *is* the conv padding.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "epilogue.cuh"
#include "tc_common.cuh"

namespace hfagp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // bf16 elements per K chunk = one 128 B swizzle row
constexpr int TC_ROW = TC_BK * 2;               // bytes of one pixel row of a chunk (= the swizzle span)
constexpr int TC_THREADS = 320;                 // warp0 TMA, warp1 MMA + TMEM alloc, warps 2..9 epilogue
constexpr int TC_MAX_CLASSES = 4;               // sub-problems of one launch (output parity classes)
constexpr int TC_MAX_STAGES = 4;

// One sub-problem of a launch: its own output window and tap list, sharing operands and epilogue with the others.
// Taps are ordered in GROUPS: the taps of a group read the same shared-memory input patch (loaded once per K chunk)
// at different row offsets, so a 3x3 stride-1 convolution loads 3 patches instead of 9 tiles.
struct TcClass {
  int oh, ow, off_y, off_x;
  int tiles_x, tiles_y, tile_begin;             // M tiles of this class inside one frame (x splits = work items)
  int ntaps;
  int ngroups;                                  // tap groups (patches) per K chunk
  int splits;                                   // split-K factor of this class: its kchunks*ngroups units are dealt
                                                // out to `splits` CTAs per tile (1 = no split)
  int8_t gstart[HFAGP_MAX_TAPS + 1];            // taps of group g: [gstart[g], gstart[g+1])
  int8_t dx[HFAGP_MAX_TAPS];                    // patch origin (x) of the tap's group
  int8_t dy0[HFAGP_MAX_TAPS];                   // patch origin (y) of the tap's group
  int8_t row_off[HFAGP_MAX_TAPS];               // tap's first patch row (dy - dy0)
  int8_t first[HFAGP_MAX_TAPS];                 // 1: first tap of its group (a new patch is loaded)
  int8_t last[HFAGP_MAX_TAPS];                  // 1: last tap of its group (the patch slot is released)
  int8_t wtap[HFAGP_MAX_TAPS];
};

struct TcParams {
  ConvParams cp;          // epilogue operands + shared geometry (x / w pointers unused here)
  TcClass cls[TC_MAX_CLASSES];
  int ncls;
  int tile_w, tile_h;     // BW x BH = 128, BW % 8 == 0
  int patch_rows;         // rows of one A patch: BH + max dy-span of a group
  int bn;                 // N tile (multiple of 16, <= 128)
  int n_tiles;            // N tiles
  int tiles_per_frame;    // M tiles of all classes
  int total_tiles;        // batch * tiles_per_frame * n_tiles
  int taps_per_frame;     // weight taps stored per batch sample (B map z-coordinate stride)
  int w_batched;          // 1: weights are per sample
  int a_stages, b_stages;
  int accumulate;         // 1: split-K mode — raw fp32 partial sums are atomically added to cp.y, no epilogue
  __nv_bfloat16* y_hi;    // split output (or null -> cp.y fp32)
  __nv_bfloat16* y_lo;
  // fused small ToRGB (super-resolution blocks): rgb_acc[n][oy][ox][o] += sum_co y[..][co] * rgb_w[n][o][co], o < rgb_k,
  // taken from the finished activations while they are still in registers (saves re-reading the layer output)
  const float* rgb_w;
  float* rgb_acc;
  int rgb_k;
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a TMA box (no shared-memory destination): issued one tile ahead for the input patches
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

#ifdef HFAGP_TC_TIMING
// per-CTA cycle counters (debug builds only): [0] MMA loop total, [1] wait acc_empty, [2] wait a_full, [3] wait b_full,
// [4] epilogue warp 2 total, [5] its wait on acc_full, [6] tiles
__device__ unsigned long long tc_dbg[256][8];
#define TC_T0(var) const long long var = clock64()
#define TC_ADD(slot, var) dbg_local[slot] += clock64() - var
#else
#define TC_T0(var)
#define TC_ADD(slot, var)
#endif

// ---- CTA-pair (cta_group::2) forms.  A shared::cta address of the executing CTA is a valid shared::cluster address;
// bit 24 of it is the CTA's rank inside the pair, so clearing it names the same object in the LEADER (rank 0) CTA.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(const CUtensorMap* map, void* smem, uint32_t bar_addr, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(const CUtensorMap* map, void* smem, uint32_t bar_addr, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this address in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// tile id -> (class, batch sample, tile origin, N offset).  N tiles are the fastest index so that the CTAs running
// side by side read the same input patch (L2 hits), then M tiles, then classes, then batch samples.
struct TileCoord { int c, n, y0, x0, n0, u0, u1; };   // [u0, u1): (K chunk, tap group) units of this work item
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t, int kchunks, int pair_rank = -1) {
  TileCoord tc;
  const int nt = t % p.n_tiles;
  int r = t / p.n_tiles;
  tc.n = r / p.tiles_per_frame;
  r -= tc.n * p.tiles_per_frame;
  int c = 0;
  while (c + 1 < p.ncls && r >= p.cls[c + 1].tile_begin) ++c;
  r -= p.cls[c].tile_begin;
  const int splits = p.cls[c].splits;
  const int ks = r % splits;
  r /= splits;
  const int units = kchunks * p.cls[c].ngroups;
  tc.u0 = (int)((long long)units * ks / splits);
  tc.u1 = (int)((long long)units * (ks + 1) / splits);
  const int ty = r / p.cls[c].tiles_x;
  tc.c = c;
  tc.y0 = ty * p.tile_h;
  // pair mode (cta_group::2): tiles_x counts PAIRS of horizontally adjacent tiles; this CTA takes tile 2*px + rank
  // (a tile beyond the image is harmless: TMA zero-fills its patches and the epilogue finds no valid pixel)
  const int px = r - ty * p.cls[c].tiles_x;
  tc.x0 = (pair_rank < 0 ? px : 2 * px + pair_rank) * p.tile_w;
  tc.n0 = nt * p.bn;
  return tc;
}

// Persistent kernel: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Two shared-memory rings (input
// patches, weight tiles) run ahead across tile boundaries and the fp32 accumulator is double-buffered in TMEM, so
// the epilogue of tile i overlaps the MMAs of tile i+1.
//
// TWO = true is the CTA-pair form (cluster of 2, tcgen05 cta_group::2): the pair computes two horizontally adjacent
// M tiles as ONE M=256 MMA issued by the leader CTA.  Each CTA stages its own input patches and only part of every
// weight tile — rank 0 the hi half (x_hi*w_hi columns), rank 1 the lo half (x_hi*w_lo columns), and each one half of
// the w_hi rows for the x_lo*w_hi MMA: 24 KB per tap instead of 32 KB written by TMA and 14 KB instead of 20 KB read
// by the MMAs per K step and SM.  SS-mode MMAs at this tile size are shared-memory-bandwidth bound
// (profiles/r1_conv_tc_notes.md), so that is what buys speed.  TMA completions of both CTAs land on the leader's
// full barriers; the leader's commits are multicast to both CTAs' empty / accumulator-full barriers; both CTAs'
// epilogue warps arrive on the leader's accumulator-empty barrier.
template <bool TWO>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const __grid_constant__ CUtensorMap map_b_half, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_half = p.patch_rows * p.tile_w * TC_ROW;          // bytes of one (hi | lo) patch, multiple of 1024
  const int a_stage = 2 * a_half;
  const int b_bytes = p.bn * TC_ROW;
  const int b_half = (b_bytes + 1023) & ~1023;
  // pair form: [bn rows: this CTA's half of (w_hi ; w_lo)] [bn/2 rows: this CTA's half of w_hi]
  const int b_stage = TWO ? b_bytes + b_bytes / 2 : 2 * b_half;
  const int rank = TWO ? (int)cluster_ctarank() : 0;
  const int pair_rank = TWO ? rank : -1;
  const int wid = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // work-item lane: cluster or CTA index
  const int wstride = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * a_stage;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + p.b_stages * b_stage);
  uint64_t* a_empty = a_full + TC_MAX_STAGES;
  uint64_t* b_full = a_empty + TC_MAX_STAGES;
  uint64_t* b_empty = b_full + TC_MAX_STAGES;
  uint64_t* acc_full = b_empty + TC_MAX_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // per-channel epilogue scale / shift of the current N tile (512 B past the barrier block)
  float* epi_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512);
  float* epi_sh = epi_sc + TC_BM;
  float* epi_rgb = epi_sh + TC_BM;          // [4][TC_BM] ToRGB weights of the current (sample, N tile)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const HfagpConvDesc& d = p.cp.d;
  const int kchunks = (d.cin + TC_BK - 1) / TC_BK;   // a partial last chunk is zero-filled by TMA (both operands)
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 4 * p.bn) tmem_cols <<= 1;   // 2 buffers x (hi|lo weight halves: 2*bn columns)

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) {
      mbar_init(&a_full[s], TWO ? 2 : 1);        // pair form: one expect_tx arrival per CTA of the pair
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], TWO ? 2 : 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], TWO ? 16 : 8);    // one arrival per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (TWO) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs initialised before any remote arrival
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only shared / tensor memory: under programmatic dependent launch it overlapped the
  // previous kernel's tail; global memory is read and written only from here on
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, elected lane issues): patches and weight tiles in the exact order the
    // MMA issuer consumes them
    {
      uint32_t ia = 0, ib = 0;                     // running slot counters of the two rings
      for (int t = wid; t < p.total_tiles; t += wstride) {
        const TileCoord tc = decode_tile(p, t, kchunks, pair_rank);
        const TcClass& c = p.cls[tc.c];
        const int wz0 = p.w_batched ? tc.n * p.taps_per_frame : 0;
        if (t + wstride < p.total_tiles) {
          // pull the NEXT work item's input patches into L2 while this one streams (they are first-touch data)
          const TileCoord nx = decode_tile(p, t + wstride, kchunks, pair_rank);
          const TcClass& cn = p.cls[nx.c];
          for (int u = nx.u0; u < nx.u1; ++u) {
            const int kc = u / cn.ngroups, g = u - kc * cn.ngroups;
            const int tp = cn.gstart[g];
            const int ax = nx.x0 * d.in_stride + cn.dx[tp], ay = nx.y0 * d.in_stride + cn.dy0[tp];
            if (elect_one()) {
              tma_prefetch_4d(&map_a_hi, kc * TC_BK, ax, ay, nx.n);
              tma_prefetch_4d(&map_a_lo, kc * TC_BK, ax, ay, nx.n);
            }
          }
        }
        for (int u = tc.u0; u < tc.u1; ++u) {
          const int kc = u / c.ngroups, g = u - kc * c.ngroups;
          const int c0 = kc * TC_BK;
          for (int tp = c.gstart[g]; tp < c.gstart[g + 1]; ++tp) {
            if (c.first[tp]) {
              const int s = ia % p.a_stages;
              mbar_wait(&a_empty[s], ((ia / p.a_stages) & 1) ^ 1);
              const int ax = tc.x0 * d.in_stride + c.dx[tp], ay = tc.y0 * d.in_stride + c.dy0[tp];
              if (elect_one()) {
                if (TWO) {
                  const uint32_t bar = smem_u32(&a_full[s]) & PEER_BIT_MASK;       // the LEADER's barrier
                  mbar_expect_tx_cluster(bar, 2 * a_half);
                  tma_load_4d_2cta(&map_a_hi, smem_a + s * a_stage, bar, c0, ax, ay, tc.n);
                  tma_load_4d_2cta(&map_a_lo, smem_a + s * a_stage + a_half, bar, c0, ax, ay, tc.n);
                } else {
                  mbar_expect_tx(&a_full[s], 2 * a_half);
                  tma_load_4d(&map_a_hi, smem_a + s * a_stage, &a_full[s], c0, ax, ay, tc.n);
                  tma_load_4d(&map_a_lo, smem_a + s * a_stage + a_half, &a_full[s], c0, ax, ay, tc.n);
                }
              }
              ++ia;
            }
            const int s = ib % p.b_stages;
            mbar_wait(&b_empty[s], ((ib / p.b_stages) & 1) ^ 1);
            if (elect_one()) {
              if (TWO) {
                const uint32_t bar = smem_u32(&b_full[s]) & PEER_BIT_MASK;         // the LEADER's barrier
                mbar_expect_tx_cluster(bar, b_bytes + b_bytes / 2);
                tma_load_3d_2cta(rank == 0 ? &map_b_hi : &map_b_lo, smem_b + s * b_stage, bar, c0, tc.n0, wz0 + c.wtap[tp]);
                tma_load_3d_2cta(&map_b_half, smem_b + s * b_stage + b_bytes, bar, c0, tc.n0 + rank * (p.bn / 2),
                                 wz0 + c.wtap[tp]);
              } else {
                mbar_expect_tx(&b_full[s], 2 * b_bytes);
                tma_load_3d(&map_b_hi, smem_b + s * b_stage, &b_full[s], c0, tc.n0, wz0 + c.wtap[tp]);
                tma_load_3d(&map_b_lo, smem_b + s * b_stage + b_half, &b_full[s], c0, tc.n0, wz0 + c.wtap[tp]);
              }
            }
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, elected lane issues); pair form: the leader CTA only
    if (!TWO || rank == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((TC_BM >> 4) << 24);
      // the same with N = 2*bn: the hi and lo halves of a weight tile are adjacent in shared memory, so
      // x_hi * [w_hi ; w_lo] is ONE instruction writing accumulator columns [0,bn) and [bn,2bn) — x_hi is read from
      // shared memory once instead of twice (SS-mode MMAs at N=128 are shared-memory-read bound: 8 KB per 64 cycles)
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 2) << 17) | ((TC_BM >> 4) << 24);
      // pair form: M = 256 across the two CTAs
      const uint32_t idesc_p = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((2 * TC_BM >> 4) << 24);
      const uint32_t idesc2_p = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 2) << 17) | ((2 * TC_BM >> 4) << 24);
      const bool wide = b_half == b_bytes;         // halves contiguous (bn*128 B is a multiple of 1024: bn % 8 == 0)
      const int row_bytes = p.tile_w * TC_ROW;     // one patch row of pixels (multiple of 1024: BW % 8 == 0)
      uint32_t ia = 0, ib = 0, j = 0;
#ifdef HFAGP_TC_TIMING
      long long dbg_local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
      TC_T0(t_loop);
      for (int t = wid; t < p.total_tiles; t += wstride, ++j) {
        const TileCoord tc = decode_tile(p, t, kchunks, pair_rank);
        const TcClass& c = p.cls[tc.c];
        const uint32_t buf = j & 1;
        TC_T0(t_ae);
        mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
        TC_ADD(1, t_ae);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * 2 * p.bn;
        uint32_t a_addr = 0, sa = 0;
        uint32_t started = 0;
        for (int u = tc.u0; u < tc.u1; ++u) {
          const int g = u % c.ngroups;
          for (int tp = c.gstart[g]; tp < c.gstart[g + 1]; ++tp) {
            if (c.first[tp]) {
              sa = ia % p.a_stages;
              TC_T0(t_af);
              mbar_wait(&a_full[sa], (ia / p.a_stages) & 1);
              TC_ADD(2, t_af);
              a_addr = smem_u32(smem_a + sa * a_stage);
              ++ia;
            }
            const uint32_t sb = ib % p.b_stages;
            TC_T0(t_bf);
            mbar_wait(&b_full[sb], (ib / p.b_stages) & 1);
            TC_ADD(3, t_bf);
            ++ib;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors of k-step 0; a k-step advances the start-address field by 32 B >> 4 = 2
            const uint64_t da_hi = umma_desc(a_addr + c.row_off[tp] * row_bytes);
            const uint64_t da_lo = umma_desc(a_addr + c.row_off[tp] * row_bytes + a_half);
            const uint64_t db_hi = umma_desc(smem_u32(smem_b + sb * b_stage));
            const uint64_t db_lo = umma_desc(smem_u32(smem_b + sb * b_stage) + (TWO ? b_bytes : b_half));
            const bool last = c.last[tp] != 0;
            if (TWO) {
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  // cols [0,bn) += x_hi*w_hi (rank 0's rows), [bn,2bn) += x_hi*w_lo (rank 1's rows)
                  umma_bf16_2cta(acc, da_hi + 2 * k, db_hi + 2 * k, idesc2_p, (k == 0) ? started : 1u);
                  // cols [0,bn) += x_lo*w_hi: each CTA holds bn/2 of the w_hi rows behind its hi|lo block
                  umma_bf16_2cta(acc, da_lo + 2 * k, db_lo + 2 * k, idesc_p, 1u);
                }
                umma_commit_2cta(&b_empty[sb]);
                if (last) umma_commit_2cta(&a_empty[sa]);
              }
            } else if (elect_one()) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                const uint32_t first = (k == 0) ? started : 1u;
                if (wide) {
                  umma_bf16(acc, da_hi + 2 * k, db_hi + 2 * k, idesc2, first);         // cols [0,bn) += hi*hi, [bn,2bn) += hi*lo
                } else {
                  umma_bf16(acc, da_hi + 2 * k, db_hi + 2 * k, idesc, first);
                  umma_bf16(acc + p.bn, da_hi + 2 * k, db_lo + 2 * k, idesc, first);
                }
                umma_bf16(acc, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);               // cols [0,bn) += lo*hi
              }
              umma_commit(&b_empty[sb]);               // frees the weight slot once these MMAs have read it
              if (last) umma_commit(&a_empty[sa]);     // ... and the patch slot after its last tap
            }
            __syncwarp();
            started = 1;
          }
        }
        if (elect_one()) {                             // accumulator complete
          if (TWO) umma_commit_2cta(&acc_full[buf]); else umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
#ifdef HFAGP_TC_TIMING
      TC_ADD(0, t_loop);
      if (lane == 0) {
        for (int i = 0; i < 4; ++i) tc_dbg[blockIdx.x][i] = dbg_local[i];
        tc_dbg[blockIdx.x][6] = j;
      }
#endif
    }
  } else {
    // ===== epilogue: 8 warps; warp w may touch TMEM lanes [32*(w%4), +32), the two warps of a quadrant split the
    // accumulator columns in 32-wide chunks.  Per-channel scale/shift (demodulation, bias) are staged in shared
    // memory once per (sample, N tile); the common epilogue is branch-free.
    const int ew = warp - 2;
    const int q = warp & 3, half = ew >> 2;
    const int r = q * 32 + lane;                     // accumulator row = pixel inside the tile
    const int ly = r / p.tile_w, lx = r - ly * p.tile_w;
    const int et = threadIdx.x - 64;                 // 0..255
    const bool generic = (p.cp.residual != nullptr || p.cp.up_img != nullptr) && (d.cout & 3) != 0;
    const bool add_res = p.cp.residual != nullptr && !generic;
    const bool add_up = p.cp.up_img != nullptr && !generic;
    const float slope = act_slope(d.act);
    const float gain = d.act_gain;
    const float cl = d.clamp > 0.f ? d.clamp : __int_as_float(0x7f800000);
    int staged_n = -1, staged_n0 = -1;
    uint32_t j = 0;
#ifdef HFAGP_TC_TIMING
    long long dbg_local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    TC_T0(t_epi);
    for (int t = wid; t < p.total_tiles; t += wstride, ++j) {
      const TileCoord tc = decode_tile(p, t, kchunks, pair_rank);
      const TcClass& c = p.cls[tc.c];
      const uint32_t buf = j & 1;
      if (tc.n != staged_n || tc.n0 != staged_n0) {  // uniform over the epilogue threads
        asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone is done reading the previous vectors
        if (et < p.bn) {
          const int co = tc.n0 + et;
          const bool in = co < d.cout;
          epi_sc[et] = (in && p.cp.dcoef) ? __ldg(p.cp.dcoef + (size_t)tc.n * d.cout + co) : 1.f;
          epi_sh[et] = (in && p.cp.bias) ? __ldg(p.cp.bias + co) : 0.f;
          if (p.rgb_acc)
            for (int o = 0; o < p.rgb_k; ++o)
              epi_rgb[o * TC_BM + et] = in ? __ldg(p.rgb_w + ((size_t)tc.n * p.rgb_k + o) * d.cout + co) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        staged_n = tc.n;
        staged_n0 = tc.n0;
      }
      const int my = tc.y0 + ly, mx = tc.x0 + lx;
      const bool valid = my < c.oh && mx < c.ow;
      EpiCtx ec;
      epi_setup_at(ec, p.cp, tc.n, valid ? my : 0, valid ? mx : 0, c.off_y, c.off_x);
      TC_T0(t_afl);
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      TC_ADD(5, t_afl);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * p.bn;
      bool arrived = false;
      for (int cb = half * 32; cb < p.bn; cb += 64) {
        float v[32];
        {
          float v2[32];
          tmem_ld32(acc + cb, v);                      // warp-collective: no early exit before this
          tmem_ld32(acc + p.bn + cb, v2);              // the x_hi * w_lo partial sums
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] += v2[jj];
        }
        if (cb + 64 >= p.bn) {
          // last read of this accumulator by this warp: hand it back to the MMA issuer before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (TWO) mbar_arrive_cluster(smem_u32(&acc_empty[buf]) & PEER_BIT_MASK); else mbar_arrive(&acc_empty[buf]);
          }
          arrived = true;
        }
        if (!valid) continue;
        const int co0 = tc.n0 + cb;
        if (co0 >= d.cout) continue;
        const bool full = co0 + 32 <= d.cout;
        if (p.accumulate) {
          float* o = p.cp.y + ec.out_base + co0;
          if (full && (d.cout & 3) == 0) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              atomicAdd(reinterpret_cast<float4*>(o + jj), make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (co0 + jj < d.cout) atomicAdd(o + jj, v[jj]);
          }
          continue;
        }
        if (!generic && full) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(&epi_sc[cb + jj]);
            const float4 sh = *reinterpret_cast<const float4*>(&epi_sh[cb + jj]);
            float a0 = fmaf(v[jj], sc.x, sh.x + ec.nz), a1 = fmaf(v[jj + 1], sc.y, sh.y + ec.nz);
            float a2 = fmaf(v[jj + 2], sc.z, sh.z + ec.nz), a3 = fmaf(v[jj + 3], sc.w, sh.w + ec.nz);
            a0 = fmaxf(a0, slope * a0) * gain; a1 = fmaxf(a1, slope * a1) * gain;
            a2 = fmaxf(a2, slope * a2) * gain; a3 = fmaxf(a3, slope * a3) * gain;
            v[jj] = fminf(fmaxf(a0, -cl), cl); v[jj + 1] = fminf(fmaxf(a1, -cl), cl);
            v[jj + 2] = fminf(fmaxf(a2, -cl), cl); v[jj + 3] = fminf(fmaxf(a3, -cl), cl);
          }
          if (add_res) {                               // ResBlock merge (v + skip) * 1/sqrt2: 16-byte loads
            const float rs = d.residual_scale;
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(ec.res + co0 + jj));
              v[jj] = (v[jj] + r4.x) * rs; v[jj + 1] = (v[jj + 1] + r4.y) * rs;
              v[jj + 2] = (v[jj + 2] + r4.z) * rs; v[jj + 3] = (v[jj + 3] + r4.w) * rs;
            }
          }
          if (add_up) {                                // + upsample2d(previous skip image): 4 taps x 16-byte loads
            const UpTaps ut = upsample_taps(d.up_h, d.up_w, ec.oy, ec.ox);
            const float* ub = ec.up + co0;
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
#pragma unroll
              for (int tp = 0; tp < 4; ++tp) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(ub + (size_t)ut.off[tp] * d.cout + jj));
                v[jj] = fmaf(ut.w[tp], u.x, v[jj]); v[jj + 1] = fmaf(ut.w[tp], u.y, v[jj + 1]);
                v[jj + 2] = fmaf(ut.w[tp], u.z, v[jj + 2]); v[jj + 3] = fmaf(ut.w[tp], u.w, v[jj + 3]);
              }
            }
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (co0 + jj < d.cout) v[jj] = epi_apply(ec, p.cp, v[jj], co0 + jj);
        }
        if (p.rgb_acc) {
          float pr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (o < p.rgb_k) {
                const float4 w4 = *reinterpret_cast<const float4*>(&epi_rgb[o * TC_BM + cb + jj]);
                pr[o] = fmaf(v[jj], w4.x, fmaf(v[jj + 1], w4.y, fmaf(v[jj + 2], w4.z, fmaf(v[jj + 3], w4.w, pr[o]))));
              }
            }
          }
          float* ra = p.rgb_acc + (ec.out_base / d.cout) * p.rgb_k;
          for (int o = 0; o < p.rgb_k; ++o) atomicAdd(ra + o, pr[o]);
        }
        if (p.y_hi) {
          __nv_bfloat16* oh = p.y_hi + ec.out_base + co0;
          __nv_bfloat16* ol = p.y_lo + ec.out_base + co0;
          if (full && (d.cout & 7) == 0) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 8) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat16 h0 = __float2bfloat16_rn(v[jj + 2 * e]), h1 = __float2bfloat16_rn(v[jj + 2 * e + 1]);
                __nv_bfloat16 l0 = __float2bfloat16_rn(v[jj + 2 * e] - __bfloat162float(h0));
                __nv_bfloat16 l1 = __float2bfloat16_rn(v[jj + 2 * e + 1] - __bfloat162float(h1));
                hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
              *reinterpret_cast<uint4*>(oh + jj) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(ol + jj) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (co0 + jj < d.cout) {
                __nv_bfloat16 h = __float2bfloat16_rn(v[jj]);
                oh[jj] = h;
                ol[jj] = __float2bfloat16_rn(v[jj] - __bfloat162float(h));
              }
          }
        } else {
          float* o = p.cp.y + ec.out_base + co0;
          if (full && (d.cout & 3) == 0) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              *reinterpret_cast<float4*>(o + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (co0 + jj < d.cout) o[jj] = v[jj];
          }
        }
      }
      if (!arrived) {                                  // this warp had no column chunk (bn <= 32)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_cluster(smem_u32(&acc_empty[buf]) & PEER_BIT_MASK); else mbar_arrive(&acc_empty[buf]);
        }
      }
    }
#ifdef HFAGP_TC_TIMING
    TC_ADD(4, t_epi);
    if (warp == 2 && lane == 0) { tc_dbg[blockIdx.x][4] = dbg_local[4]; tc_dbg[blockIdx.x][5] = dbg_local[5]; }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (TWO) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (TWO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, d2, d3;
  uint32_t b0, b1, b2, b3, es;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && b0 == o.b0 && b1 == o.b1 &&
           b2 == o.b2 && b3 == o.b3 && es == o.es;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.d2); mix(k.d3); mix(k.b0); mix(k.b1); mix(k.b2); mix(k.b3); mix(k.es);
    return h;
  }
};

// rank-4 bf16 map (rank-3 tensors pass d3 = b3 = 1).  Cached: PyTorch's allocator hands the same
// pointers back every frame, so steady state does no driver calls.
int get_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b0,
            uint32_t b1, uint32_t b2, uint32_t b3, uint32_t estride, int rank) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, d0, d1, d2, d3, b0, b1, b2, b3, estride * 8u + (uint32_t)rank};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return HFAGP_OK;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(HFAGP_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
  cuuint32_t box[4] = {b0, b1, b2, b3};
  cuuint32_t es[4] = {1, estride, estride, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box,
                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HFAGP_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return HFAGP_OK;
}

}  // namespace hfagp

using namespace hfagp;

// ---------------------------------------------------------------- host: problem setup
namespace hfagp {

// Tap list of one desc -> grouped class.  patch mode (in_stride == 1): taps that share dx read one input patch.
static void build_class(const HfagpConvDesc& d, bool patch_mode, TcClass& c, int& max_span) {
  c.oh = d.oh; c.ow = d.ow; c.off_y = d.out_off_y; c.off_x = d.out_off_x;
  int order[HFAGP_MAX_TAPS];
  for (int t = 0; t < d.ntaps; ++t) order[t] = t;
  if (patch_mode) {   // sort by (dx, dy): insertion sort, <= 16 entries
    for (int i = 1; i < d.ntaps; ++i)
      for (int j = i; j > 0; --j) {
        const int a = order[j - 1], b = order[j];
        if (d.dx[a] > d.dx[b] || (d.dx[a] == d.dx[b] && d.dy[a] > d.dy[b])) { order[j - 1] = b; order[j] = a; } else break;
      }
  }
  c.ntaps = d.ntaps;
  c.ngroups = 0;
  c.splits = 1;
  int i = 0;
  while (i < d.ntaps) {
    int j = i;
    if (patch_mode) while (j + 1 < d.ntaps && d.dx[order[j + 1]] == d.dx[order[i]]) ++j;
    const int dy0 = d.dy[order[i]];
    const int span = d.dy[order[j]] - dy0;
    if (span > max_span) max_span = span;
    for (int k = i; k <= j; ++k) {
      const int t = order[k];
      c.dx[k] = (int8_t)d.dx[t];
      c.dy0[k] = (int8_t)dy0;
      c.row_off[k] = (int8_t)(d.dy[t] - dy0);
      c.first[k] = k == i;
      c.last[k] = k == j;
      c.wtap[k] = (int8_t)d.wtap[t];
    }
    c.gstart[c.ngroups++] = (int8_t)i;
    i = j + 1;
  }
  c.gstart[c.ngroups] = (int8_t)d.ntaps;
}

static int launch_tc(const HfagpConvDesc* descs, int ndesc, const uint16_t* x_hi, const uint16_t* x_lo,
                     const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, const float* dcoef,
                     const float* noise, const float* bias, const float* residual, const float* up_img, float* y,
                     uint16_t* y_hi, uint16_t* y_lo, int ksplit, void* stream, const char* who,
                     const float* rgb_w = nullptr, int rgb_k = 0, float* rgb_acc = nullptr) {
  HFAGP_CHECK_ARG(descs && ndesc >= 1 && ndesc <= TC_MAX_CLASSES, "%s: 1..%d descs", who, TC_MAX_CLASSES);
  HFAGP_CHECK_ARG(x_hi && x_lo && w_hi && w_lo, "%s: null pointer", who);
  HFAGP_CHECK_ARG((y != nullptr) != (y_hi != nullptr && y_lo != nullptr), "%s: give y or (y_hi, y_lo)", who);
  const HfagpConvDesc& d = descs[0];
  HFAGP_CHECK_ARG(d.batch > 0 && d.batch <= 65535 && d.cout > 0, "%s: bad dims", who);
  HFAGP_CHECK_ARG(d.cin % 8 == 0, "%s: cin must be a multiple of 8 (got %d)", who, d.cin);
  HFAGP_CHECK_ARG(d.in_stride == 1 || d.in_stride == 2, "%s: in_stride must be 1 or 2", who);
  HFAGP_CHECK_ARG(!up_img || (d.up_h * 2 == d.out_h && d.up_w * 2 == d.out_w), "%s: up_img must be out/2", who);
  HFAGP_CHECK_ARG(w_taps_total > 0 && w_taps_total <= 127, "%s: w_taps_total", who);
  for (int i = 0; i < ndesc; ++i) {
    const HfagpConvDesc& e = descs[i];
    HFAGP_CHECK_ARG(e.oh > 0 && e.ow > 0 && e.ntaps > 0 && e.ntaps <= HFAGP_MAX_TAPS, "%s: desc %d: empty window / taps", who, i);
    HFAGP_CHECK_ARG((e.oh - 1) * e.out_stride + e.out_off_y < e.out_h && (e.ow - 1) * e.out_stride + e.out_off_x < e.out_w,
                    "%s: desc %d: output window exceeds out_h/out_w", who, i);
    HFAGP_CHECK_ARG(e.batch == d.batch && e.in_h == d.in_h && e.in_w == d.in_w && e.cin == d.cin && e.cout == d.cout &&
                        e.in_stride == d.in_stride && e.out_h == d.out_h && e.out_w == d.out_w &&
                        e.out_stride == d.out_stride && e.w_batch_stride == d.w_batch_stride && e.act == d.act &&
                        e.act_gain == d.act_gain && e.clamp == d.clamp && e.noise_gain == d.noise_gain &&
                        e.residual_scale == d.residual_scale && e.up_h == d.up_h && e.up_w == d.up_w,
                    "%s: desc %d differs from desc 0 in more than its window and taps", who, i);
    for (int t = 0; t < e.ntaps; ++t)
      HFAGP_CHECK_ARG(e.dy[t] >= -64 && e.dy[t] <= 63 && e.dx[t] >= -64 && e.dx[t] <= 63 && e.wtap[t] >= 0 &&
                          e.wtap[t] < w_taps_total, "%s: desc %d: tap %d out of range", who, i, t);
  }

  TcParams p;
  p.cp = ConvParams{d, nullptr, nullptr, dcoef, noise, bias, residual, up_img, y};
  p.y_hi = reinterpret_cast<__nv_bfloat16*>(y_hi);
  p.y_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
  p.ncls = ndesc;
  p.accumulate = ksplit > 0;
  p.rgb_w = rgb_w;
  p.rgb_acc = rgb_acc;
  p.rgb_k = rgb_k;
  HFAGP_CHECK_ARG(!rgb_acc || (rgb_w && rgb_k >= 1 && rgb_k <= 4 && ksplit == 0 && !((residual || up_img) && (d.cout & 3)) &&
                               d.cout % 32 == 0 && d.out_stride == 1),
                  "%s: fused ToRGB needs rgb_w, 1..4 outputs, a dense non-split-K layer with cout %% 32 == 0", who);
  const bool patch_mode = d.in_stride == 1;
  int max_span = 0;
  for (int i = 0; i < ndesc; ++i) build_class(descs[i], patch_mode, p.cls[i], max_span);
  p.bn = d.cout >= 128 ? 128 : ((d.cout + 15) / 16) * 16;
  p.n_tiles = cdiv(d.cout, p.bn);
  // M tile shape BW x BH (= 128 pixels, BW % 8 == 0): least operand traffic = padded tiles x (patch rows + weight rows)
  double best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const int bh = 128 / bw;
    double cost = 0;
    for (int i = 0; i < ndesc; ++i) {
      const TcClass& c = p.cls[i];
      int groups = 0;
      for (int t = 0; t < c.ntaps; ++t) groups += c.first[t];
      const double tiles = (double)cdiv(c.ow, bw) * cdiv(c.oh, bh);
      cost += tiles * (groups * (double)(bh + max_span) * bw + (double)c.ntaps * p.bn);
    }
    if (best < 0 || cost < best) {
      best = cost;
      p.tile_w = bw;
      p.tile_h = bh;
    }
  }
  p.patch_rows = p.tile_h + max_span;
  // CTA-pair form (cta_group::2) for the big layers: at least one pair of M tiles per pair of SMs, no split-K
  int m_tiles = 0;
  for (int i = 0; i < ndesc; ++i) m_tiles += cdiv(p.cls[i].ow, p.tile_w) * cdiv(p.cls[i].oh, p.tile_h);
  static const bool pair_off = getenv("HFAGP_TC_NO_PAIR") != nullptr;      // debugging aid: force the single-CTA form
  const bool two = !pair_off && ksplit == 0 && (long long)m_tiles * d.batch * p.n_tiles >= 100 && p.bn % 16 == 0;
  int tiles = 0;
  const int kchunks_h = cdiv(d.cin, TC_BK);
  for (int i = 0; i < ndesc; ++i) {
    TcClass& c = p.cls[i];
    c.tiles_x = cdiv(c.ow, p.tile_w);
    if (two) c.tiles_x = cdiv(c.tiles_x, 2);       // pairs of horizontally adjacent tiles
    c.tiles_y = cdiv(c.oh, p.tile_h);
    c.tile_begin = tiles;
    const int units = kchunks_h * c.ngroups;
    c.splits = ksplit > 1 ? (ksplit < units ? ksplit : units) : 1;
    tiles += c.tiles_x * c.tiles_y * c.splits;
  }
  p.tiles_per_frame = tiles;
  p.total_tiles = tiles * p.n_tiles * d.batch;
  p.w_batched = d.w_batch_stride != 0;
  p.taps_per_frame = w_taps_total;

  // shared-memory rings: as deep as 225 KB allows
  const int a_stage = 2 * p.patch_rows * p.tile_w * TC_ROW;
  const int b_stage = two ? p.bn * TC_ROW * 3 / 2 : 2 * ((p.bn * TC_ROW + 1023) & ~1023);
  const int extra = 1024 /*alignment*/ + 512 /*barriers*/ + 6 * TC_BM * 4 /*epilogue vectors + ToRGB weights*/;
  const int budget = 227 * 1024 - extra;
  // the input patches of a tile are first-touch (HBM latency), the weights are L2 hits: favour patch depth
  static const int choices[][2] = {{3, 4}, {3, 3}, {2, 4}, {2, 3}, {2, 2}, {1, 2}, {1, 1}};
  p.a_stages = p.b_stages = 0;
  for (const auto& ch : choices)
    if (ch[0] * a_stage + ch[1] * b_stage <= budget) {
      p.a_stages = ch[0];
      p.b_stages = ch[1];
      break;
    }
  HFAGP_CHECK_ARG(p.a_stages > 0, "%s: tile does not fit shared memory", who);
  const size_t smem = (size_t)p.a_stages * a_stage + (size_t)p.b_stages * b_stage + extra;

  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mb_half;
  const uint32_t es = (uint32_t)d.in_stride;
  // box extents are given in un-strided coordinates: a stride-2 traversal of BW outputs spans 2*BW-1 inputs
  const uint32_t box_w = (uint32_t)(p.tile_w * d.in_stride - (d.in_stride - 1));
  const uint32_t box_h = (uint32_t)(p.patch_rows * d.in_stride - (d.in_stride - 1));
  HFAGP_CHECK_ARG(box_w <= 256 && box_h <= 256, "%s: TMA box too large", who);
  int rc;
  if ((rc = get_map(&ma_hi, x_hi, d.cin, d.in_w, d.in_h, d.batch, TC_BK, box_w, box_h, 1, es, 4))) return rc;
  if ((rc = get_map(&ma_lo, x_lo, d.cin, d.in_w, d.in_h, d.batch, TC_BK, box_w, box_h, 1, es, 4))) return rc;
  const uint64_t wz = (uint64_t)w_taps_total * (p.w_batched ? d.batch : 1);
  if ((rc = get_map(&mb_hi, w_hi, d.cin, d.cout, wz, 1, TC_BK, p.bn, 1, 1, 1, 3))) return rc;
  if ((rc = get_map(&mb_lo, w_lo, d.cin, d.cout, wz, 1, TC_BK, p.bn, 1, 1, 1, 3))) return rc;
  mb_half = mb_hi;
  if (two && (rc = get_map(&mb_half, w_hi, d.cin, d.cout, wz, 1, TC_BK, p.bn / 2, 1, 1, 1, 3))) return rc;

  static std::atomic<uint64_t> attr_done{0};
  const cudaError_t attr_rc = per_device_once(attr_done, [] {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e != cudaSuccess ? e : cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_rc != cudaSuccess) return fail(HFAGP_E_CUDA, "%s: cudaFuncSetAttribute(max dynamic smem) failed: %s", who, cudaGetErrorString(attr_rc));
  const int num_sms = device_sm_count();
  if (two) {
    const int pairs = p.total_tiles < num_sms / 2 ? p.total_tiles : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<true>, ma_hi, ma_lo, mb_hi, mb_lo, mb_half, p);
    if (e != cudaSuccess) return fail(HFAGP_E_CUDA, "%s: cluster launch failed: %s", who, cudaGetErrorString(e));
    return HFAGP_OK;
  }
  {
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<false>, ma_hi, ma_lo, mb_hi, mb_lo, mb_half, p);
    if (e != cudaSuccess) return fail(HFAGP_E_CUDA, "%s: launch failed: %s", who, cudaGetErrorString(e));
  }
  return HFAGP_OK;
}

}  // namespace hfagp

extern "C" int hfagp_conv2d_tc_fwd(const HfagpConvDesc* desc, const uint16_t* x_hi, const uint16_t* x_lo,
                                   const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, const float* dcoef,
                                   const float* noise, const float* bias, const float* residual, const float* up_img,
                                   float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  return launch_tc(desc, 1, x_hi, x_lo, w_hi, w_lo, w_taps_total, dcoef, noise, bias, residual, up_img, y, y_hi, y_lo,
                   0, stream, "conv2d_tc_fwd");
}

extern "C" int hfagp_conv2d_tc_rgb_fwd(const HfagpConvDesc* desc, const uint16_t* x_hi, const uint16_t* x_lo,
                                       const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, const float* dcoef,
                                       const float* noise, const float* bias, float* y, uint16_t* y_hi, uint16_t* y_lo,
                                       const float* rgb_w, int rgb_k, float* rgb_acc, void* stream) {
  HFAGP_CHECK_ARG(rgb_w && rgb_acc, "conv2d_tc_rgb_fwd: null ToRGB operands");
  return launch_tc(desc, 1, x_hi, x_lo, w_hi, w_lo, w_taps_total, dcoef, noise, bias, nullptr, nullptr, y, y_hi, y_lo, 0,
                   stream, "conv2d_tc_rgb_fwd", rgb_w, rgb_k, rgb_acc);
}

extern "C" int hfagp_conv2d_tc_multi_fwd(const HfagpConvDesc* descs, int ndesc, const uint16_t* x_hi,
                                         const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo,
                                         int w_taps_total, const float* dcoef, const float* noise, const float* bias,
                                         const float* residual, const float* up_img, float* y, uint16_t* y_hi,
                                         uint16_t* y_lo, void* stream) {
  return launch_tc(descs, ndesc, x_hi, x_lo, w_hi, w_lo, w_taps_total, dcoef, noise, bias, residual, up_img, y, y_hi,
                   y_lo, 0, stream, "conv2d_tc_multi_fwd");
}

extern "C" int hfagp_conv2d_tc_acc_fwd(const HfagpConvDesc* descs, int ndesc, const uint16_t* x_hi, const uint16_t* x_lo,
                                       const uint16_t* w_hi, const uint16_t* w_lo, int w_taps_total, int ksplit,
                                       float* acc, void* stream) {
  HFAGP_CHECK_ARG(ksplit >= 1 && acc, "conv2d_tc_acc_fwd: ksplit >= 1 and an accumulator are required");
  return launch_tc(descs, ndesc, x_hi, x_lo, w_hi, w_lo, w_taps_total, nullptr, nullptr, nullptr, nullptr, nullptr, acc,
                   nullptr, nullptr, ksplit, stream, "conv2d_tc_acc_fwd");
}

#ifdef HFAGP_TC_TIMING
extern "C" int hfagp_debug_tc_timing(unsigned long long* host_out /* [256][8] */) {
  return cudaMemcpyFromSymbol(host_out, hfagp::tc_dbg, sizeof(unsigned long long) * 256 * 8) == cudaSuccess ? 0 : -2;
}
#endif
