// Tri-plane volume renderer, forward — decoder MLP on tcgen05 (round 2).
//
// One persistent CTA per SM walks strips of 16 rays (16 worker warps, warp = ray for the per-ray stages, + 1 MMA warp).
// The decoder runs on 128-sample tiles = 8 consecutive samples of each of the 16 rays:
//
//   G(t)   every worker gathers its ray's 8 samples.  24 lanes first build one (sample, plane) tap record each: the
//          byte offset of the 2x2 texel block (clamped inside the plane, so its four taps are IMMEDIATE offsets of one
//          address when the plane width is known at compile time) and the four bilinear weights / 3, zero where
//          grid_sample pads.  Then lane = (sample, channel quad): 12 x 16 B loads, 48 FMAs, split into bf16 hi|lo and
//          store 8 rows of the layer-1 A operand (K-major, SWIZZLE_128B, one 128 B row = [hi 32 ch | lo 32 ch],
//          two tiles in flight)                                                             -> a1_full[t % 2] (16 arrivals)
//   MMA1   D1[128 x 64] (TMEM, fp32) = A1 . [W0hi|W0hi]^T + A1[:, :32] . W0lo^T   (6 x tcgen05.mma SS, K = 16 each)
//   E1(t)  the four warps with t % 4 == warp / 4 (one per TMEM lane quadrant), thread = row: tcgen05.ld 16 columns at
//          a time -> + bias -> softplus -> split -> tcgen05.st: the layer-2 A operand lives in TENSOR MEMORY
//          (hi 32 | lo 32 columns, two bf16 per column; no shared-memory round trip, no proxy fence)
//                                                                                          -> a2_full[t % 2] (4 arrivals)
//   MMA2   D2[128 x 48] = A2hi . W1hi^T + A2lo . W1hi^T + A2hi . W1lo^T          (12 x tcgen05.mma TS: A from TMEM)
//   E2(t)  same four warps: tcgen05.ld 8 colours at a time -> + bias -> sigmoid -> 16-bit fixed point -> the colour
//          rows of the row's ray; sigma -> its density row                                 -> d2_free[t % 2] (4 arrivals)
//
// software-pipelined per worker as G(t), E1(t-1), E2(t-2) so that an MMA (and its commit latency) always has a gather
// in front of its consumer; D1, D2 and A2 are double-buffered in TMEM (512 columns).  fp32-class accuracy comes from
// the same three-term split-bf16 scheme as the convolutions.  The per-ray stages (march, smoothed pdf / cdf /
// searchsorted, stable rank-sort merge, composite) are the ones of render.cu, one warp per ray between the passes.
//
// Replaces (forward): ImportanceRenderer.forward + OSGDecoder + MipRayMarcher2 of NVlabs/eg3d, reached through
// code/networks/headnerf.py:112.
#include <cuda_bf16.h>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include "common.cuh"
#include "render_common.cuh"
#include "tc_common.cuh"

namespace hfagp {

constexpr int RT_RAYS = 16;                         // rays per strip = worker warps
constexpr int RT_THREADS = (RT_RAYS + 1) * 32;      // + the MMA warp
constexpr int RT_TILE = 8;                          // samples of one ray per MMA tile
constexpr int RT_N2 = 48;                           // layer-2 N: 32 colours, sigma, 15 zero columns
constexpr int RT_B1 = 64 * 128;                     // bytes of one layer-1 weight tile (64 rows x 128 B)
constexpr int RT_B2 = RT_N2 * 128;
constexpr int RT_A = 128 * 128;                     // bytes of one A tile (128 rows x 128 B)
constexpr int RT_REC = 112;                         // tap record of one sample: 3 plane offsets (+pad), 3 x 4 bilinear weights, each
                                                    // stored TWICE (the (w, w) operand of a packed FFMA2);
                                                    // 112 B apart = conflict-free 16 B reads by 4 samples at once
constexpr int RT_TAPS = RT_TILE * RT_REC;           // tap records of one tile of one ray
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr uint32_t RT_TMEM_COLS = 512;              // D1[2] at 0 / 64, D2[2] at 128 / 192, A2[2] (hi 32 | lo 32 columns) at 256 / 320

// per-ray scratch: colq [T][16] u32 ; dep, sig, sdep, ssig [Tp] fp32 ; tap records (alias sdep.. when they fit) ;
// order [Tp] u8 ; cdf [S+2], zmid [S] fp32
struct RtLayout {
  int colq, dep, sig, sdep, ssig, taps, order, cdf, zmid, bytes;
};
__host__ __device__ inline RtLayout rt_layout(int S, int SF) {
  const int T = S + SF, Tp = (T + 3) & ~3;
  RtLayout l;
  l.colq = 0;
  l.dep = T * 64;
  l.sig = l.dep + Tp * 4;
  l.sdep = l.sig + Tp * 4;
  l.ssig = l.sdep + Tp * 4;
  const int after = l.ssig + Tp * 4;
  l.taps = l.sdep;                                  // gathers never overlap the sorted arrays' lifetime
  const int extra = RT_TAPS > 2 * Tp * 4 ? RT_TAPS - 2 * Tp * 4 : 0;
  l.order = after + extra;
  l.cdf = (l.order + Tp + 3) & ~3;
  l.zmid = l.cdf + (S + 2) * 4;
  l.bytes = (l.zmid + S * 4 + 15) & ~15;
  return l;
}
// weights, A1[na1] (layer-1 A tiles in flight), barriers + biases, per-ray scratch
__host__ __device__ inline size_t rt_smem_bytes(int S, int SF, int na1) {
  return 1024 + 2 * RT_B1 + 2 * RT_B2 + na1 * RT_A + 1024 + (size_t)RT_RAYS * rt_layout(S, SF).bytes;
}

// byte offset of element (row, k) of a K-major SWIZZLE_128B tile (64 bf16 per 128 B row, 16 B chunks XOR row % 8)
__device__ __forceinline__ uint32_t sw128(int row, int k) {
  return (uint32_t)(row * 128 + ((((k >> 3) ^ row) & 7) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// mbarrier wait with a hardware suspend hint (the thread may sleep inside try_wait instead of spinning through issue
// slots; the hint is only an upper bound on one suspension); still bounded: 2^24 failed polls trap instead of hanging
// the GPU
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, 0;\n"
      "RT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 10000000;\n\t"
      "@p bra RT_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 16777216;\n\t"
      "@p bra RT_WAIT;\n\t"
      "trap;\n"
      "RT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (128 rows x 16 bf16 = 8 columns per K step) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(RT_RAYS * 32) : "memory"); }

// The decoder's activations on packed pairs, in log2 units (the layer weights carry the log2(e) / ln 2 factors, see the
// weight set-up): h' = softplus(x) / ln 2 = max(x', 0) + log2(1 + 2^-|x'|) with x' = x log2(e)
__device__ __forceinline__ f32x2 softplus2_log2(f32x2 x) {
  float x0, x1, t0, t1;
  upk2(x, x0, x1);
  upk2(add2(pk2(ex2f(-fabsf(x0)), ex2f(-fabsf(x1))), pk2(1.f, 1.f)), t0, t1);
  return add2(pk2(lg2f(t0), lg2f(t1)), pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}
// two colours: e = -x log2(e) (already in that form) -> round(65535 / (1 + 2^e)) packed as two u16: the 1 / 65535 is
// folded into the denominator, the rounding done by adding 2^23 (round-to-nearest-even, as __float2uint_rn)
__device__ __forceinline__ uint32_t sigmoid2_q16(f32x2 e) {
  float e0, e1, d0, d1, y0, y1;
  upk2(e, e0, e1);
  constexpr float RQ = 1.f / QSCALE;
  upk2(fma2(pk2(ex2f(e0), ex2f(e1)), pk2(RQ, RQ), pk2(RQ, RQ)), d0, d1);
  upk2(add2(pk2(rcpf(d0), rcpf(d1)), pk2(8388608.f, 8388608.f)), y0, y1);
  return __byte_perm(__float_as_uint(y0), __float_as_uint(y1), 0x5410);
}

// PWC > 0: plane width known at compile time (the four taps of a plane are immediate offsets from one address)
template <int PWC, int RT_NA1, int ESPLIT>
__global__ void __launch_bounds__(RT_THREADS, 1) render_tc_kernel(const RenderParams p, const int tw_log2) {
  // SWIZZLE_128B operand tiles need 1024 B alignment (a profiler may put its own static shared memory in front of the
  // dynamic window).  The alignment is added as an integer OFFSET to the shared array — no pointer/integer round trip —
  // so every access below stays a shared-space (LDS/STS) access.
  extern __shared__ __align__(16) uint8_t rt_smem_raw[];
  uint8_t* smem = rt_smem_raw + ((1024u - (smem_u32(rt_smem_raw) & 1023u)) & 1023u);
  const HfagpRenderDesc& d = p.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = d.s_coarse, SF = d.s_fine, T = S + SF;

  uint8_t* b1a = smem;                    // [64][hi 32 | hi 32]
  uint8_t* b1b = b1a + RT_B1;             // [64][lo 32 | 0]
  uint8_t* b2h = b1b + RT_B1;             // [48][64]
  uint8_t* b2l = b2h + RT_B2;
  uint8_t* a1 = b2l + RT_B2;              // [128][hi 32 | lo 32]
  uint64_t* a1_full = reinterpret_cast<uint64_t*>(a1 + RT_NA1 * RT_A);     // [RT_NA1]
  uint64_t* mma1_done = a1_full + 2;      // [2]  (D1 buffer; also: A1 free)
  uint64_t* a2_full = mma1_done + 2;      // [2]  by tile parity: arrivals for tile g+1 may come before tile g's are complete
  uint64_t* mma2_done = a2_full + 2;      // [2]  (D2 buffer; also: A2 free)
  uint64_t* d2_free = mma2_done + 2;      // [2]  (ESPLIT only: the epilogue group has drained D2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_free + 2);
  float* b0s = reinterpret_cast<float*>(a1_full) + 32;     // [64]
  float* b1s = b0s + RH;                                    // [48] permuted: 0..31 colours, 32 sigma
  uint8_t* rays_base = reinterpret_cast<uint8_t*>(a1_full) + 1024;
  const RtLayout L = rt_layout(S, SF);

  // ---- one-time setup: decoder weights as split-bf16 UMMA B tiles, biases, barriers, TMEM
  {
    const float* W0 = p.mlp;
    const float* B0 = W0 + RH * RC;
    const float* W1 = B0 + RH;
    const float* B1 = W1 + RO * RH;
    for (int i = threadIdx.x; i < RH * RC; i += blockDim.x) {      // (n, k): layer 1, K = channel
      const int n = i >> 5, k = i & 31;
      const float v = __ldg(W0 + i) * LOG2E;         // hidden pre-activations in log2 units: x' = x log2(e)
      const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
      *reinterpret_cast<__nv_bfloat16*>(b1a + sw128(n, k)) = h;
      *reinterpret_cast<__nv_bfloat16*>(b1a + sw128(n, k + 32)) = h;
      *reinterpret_cast<__nv_bfloat16*>(b1b + sw128(n, k)) = l;
      *reinterpret_cast<__nv_bfloat16*>(b1b + sw128(n, k + 32)) = __float2bfloat16_rn(0.f);
    }
    for (int i = threadIdx.x; i < RT_N2 * RH; i += blockDim.x) {   // (n, k): layer 2, K = hidden unit
      const int n = i >> 6, k = i & 63;
      const int o = n < 32 ? n + 1 : (n == 32 ? 0 : -1);           // output permutation: colours first, then sigma
      // the hidden layer is kept as h' = softplus(x) / ln 2, so sigma's row carries the ln 2; a colour's row computes
      // -log2(e) x (the exponent its sigmoid needs): -log2(e) ln 2 = -1, i.e. the row is just negated
      const float v = o > 0 ? -__ldg(W1 + o * RH + k) : (o == 0 ? __ldg(W1 + k) * LN2 : 0.f);
      const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
      *reinterpret_cast<__nv_bfloat16*>(b2h + sw128(n, k)) = h;
      *reinterpret_cast<__nv_bfloat16*>(b2l + sw128(n, k)) = l;
    }
    for (int i = threadIdx.x; i < RH; i += blockDim.x) b0s[i] = __ldg(B0 + i) * LOG2E;
    for (int i = threadIdx.x; i < RT_N2; i += blockDim.x) b1s[i] = i < 32 ? -LOG2E * __ldg(B1 + i + 1) : (i == 32 ? __ldg(B1) : 0.f);
    if (threadIdx.x == 0) {
      for (int s = 0; s < RT_NA1; ++s) mbar_init(&a1_full[s], RT_RAYS);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a2_full[s], ESPLIT ? 4 : RT_RAYS);
        mbar_init(&mma1_done[s], 1);
        mbar_init(&mma2_done[s], 1);
        mbar_init(&d2_free[s], 4);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == RT_RAYS) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(RT_TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();                              // the weight tiles are read by the async proxy (tcgen05.mma)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tmem_base = *tmem_slot;

  const int res = d.res;
  // a strip = a (1 << tw_log2) x (16 >> tw_log2) block of pixels: neighbouring rays share texels in all three planes
  const int tile_w = 1 << tw_log2, tile_h = RT_RAYS >> tw_log2;
  const int xtiles = (res + tile_w - 1) / tile_w, ytiles = (res + tile_h - 1) / tile_h;
  const long long strips = (long long)d.batch * ytiles * xtiles;
  const int nt_c = (S + RT_TILE - 1) / RT_TILE, nt_f = (SF + RT_TILE - 1) / RT_TILE;

  if (warp == RT_RAYS) {
    // ===== MMA issuer: warp-uniform loop, the elected lane issues (operands stay in uniform registers)
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RH >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RT_N2 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da1_0 = umma_desc(smem_u32(a1));
    const uint64_t db1a = umma_desc(smem_u32(b1a)), db1b = umma_desc(smem_u32(b1b));
    const uint64_t db2h = umma_desc(smem_u32(b2h)), db2l = umma_desc(smem_u32(b2l));
    uint32_t g1 = 0, g2 = 0;                         // tiles issued to layer 1 / layer 2 so far
    for (long long strip = blockIdx.x; strip < strips; strip += gridDim.x) {
      for (int pass = 0; pass < 2; ++pass) {
        const int nt = pass == 0 ? nt_c : nt_f;
        for (int it = 0; it <= nt; ++it) {
          if (it < nt) {
            mbar_wait_sleep(&a1_full[g1 % RT_NA1], (g1 / RT_NA1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + (g1 & 1) * 64;
            const uint64_t da1 = da1_0 + (uint64_t)((g1 % RT_NA1) * (RT_A >> 4));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(acc, da1 + 2 * k, db1a + 2 * k, idesc1, k > 0 ? 1u : 0u);   // (f_hi + f_lo) . W0hi
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(acc, da1 + 2 * k, db1b + 2 * k, idesc1, 1u);                 // f_hi . W0lo
              umma_commit(&mma1_done[g1 & 1]);
            }
            __syncwarp();
            ++g1;
          }
          if (it >= 1) {
            mbar_wait_sleep(&a2_full[g2 & 1], (g2 >> 1) & 1);
            if (ESPLIT && g2 >= 2) mbar_wait_sleep(&d2_free[g2 & 1], ((g2 - 2) >> 1) & 1);      // E2(g2 - 2) has drained this D2 buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + 128 + (g2 & 1) * 64;
            const uint32_t a2 = tmem_base + 256 + (g2 & 1) * 64;       // hidden layer in TMEM: hi at +0, lo at +32 (8 columns per K step)
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_ts(acc, a2 + 8 * k, db2h + 2 * k, idesc2, k > 0 ? 1u : 0u);
                umma_bf16_ts(acc, a2 + 32 + 8 * k, db2h + 2 * k, idesc2, 1u);
                umma_bf16_ts(acc, a2 + 8 * k, db2l + 2 * k, idesc2, 1u);
              }
              umma_commit(&mma2_done[g2 & 1]);
            }
            __syncwarp();
            ++g2;
          }
        }
      }
    }
  } else {
    // ===== workers
    const int q4 = warp & 3, cs = warp >> 2;         // TMEM lane quadrant / column slice of the epilogues
    const int erow = q4 * 32 + lane;                 // accumulator row handled in E1 / E2
    const int eray = erow >> 3, esub = erow & 7;     // = (local ray, sample inside the tile)
    uint8_t* wbase = rays_base + (size_t)warp * L.bytes;
    uint32_t* colq = reinterpret_cast<uint32_t*>(wbase + L.colq);
    float* dep = reinterpret_cast<float*>(wbase + L.dep);
    float* sig = reinterpret_cast<float*>(wbase + L.sig);
    float* sdep = reinterpret_cast<float*>(wbase + L.sdep);
    float* ssig = reinterpret_cast<float*>(wbase + L.ssig);
    uint8_t* taps = wbase + L.taps;
    uint8_t* order = wbase + L.order;
    float* cdf = reinterpret_cast<float*>(wbase + L.cdf);
    float* zmid = reinterpret_cast<float*>(wbase + L.zmid);
    uint8_t* ebase = rays_base + (size_t)eray * L.bytes;          // scratch of the ray this thread's epilogue row belongs to
    uint32_t* ecolq = reinterpret_cast<uint32_t*>(ebase + L.colq);
    float* esig = reinterpret_cast<float*>(ebase + L.sig);

    const int PW = PWC > 0 ? PWC : d.plane_w, PH = d.plane_h;
    constexpr int TEXEL_B = 3 * RC * 4;              // bytes of one texel (96 fp32 channels)
    const int row_b = PW * TEXEL_B;                  // bytes of one plane row
    uint32_t gt = 0;                                 // global tile counter (same sequence as the MMA warp's)

    for (long long strip = blockIdx.x; strip < strips; strip += gridDim.x) {
      const int n = (int)(strip / ((long long)ytiles * xtiles));
      const int rem = (int)(strip - (long long)n * ytiles * xtiles);
      const int yt = rem / xtiles, xt = rem - yt * xtiles;
      const int px_ = xt * tile_w + (warp & (tile_w - 1)), py_ = yt * tile_h + (warp >> tw_log2);
      const bool rvalid = px_ < res && py_ < res;    // warp-uniform; an invalid ray still serves its epilogue rows
      const int px = min(px_, res - 1), py = min(py_, res - 1);
      const long long ray = ((long long)n * res + py) * res + px;
      const float* cam = p.cam + (size_t)n * 25;
      // this lane's channel quad of texel (0,0): the gather adds 32-bit texel offsets to it
      const char* lb = reinterpret_cast<const char*>(p.planes + (size_t)n * PH * PW * (3 * RC)) + (lane & 7) * 16;

      // ---- ray generation (uniform across the warp), then the plane-space affine maps of this lane's plane:
      // ix = ax + t * bx, iy = ay + t * by  (grid_sample unnormalisation folded in)
      float ax, bx, ay, by;
      {
        const float inv = 1.0f / res, half = 0.5f / res;
        const float xc = px * inv + half, yc = py * inv + half;
        const float fx = __ldg(cam + 16), sk = __ldg(cam + 17), cx = __ldg(cam + 18);
        const float fy = __ldg(cam + 20), cy = __ldg(cam + 21);
        const float rfx = 1.0f / fx, rfy = 1.0f / fy;
        const float xl = (xc - cx + (cy * sk) * rfy - (sk * yc) * rfy) * rfx;
        const float yl = (yc - cy) * rfy;
        float m[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = __ldg(cam + i);
        const float ox_ = m[3], oy_ = m[7], oz_ = m[11];
        const float vx = m[0] * xl + m[1] * yl + m[2];           // world point - camera origin
        const float vy = m[4] * xl + m[5] * yl + m[6];
        const float vz = m[8] * xl + m[9] * yl + m[10];
        const float rn = rsqrtf(fmaxf(vx * vx + vy * vy + vz * vz, 1e-24f));
        const float dx_ = vx * rn, dy_ = vy * rn, dz_ = vz * rn;
        const int pidx = min(lane >> 3, 2);          // lanes 0-7 plane 0, 8-15 plane 1, 16-31 plane 2
        const float gxo = pidx == 2 ? oz_ : ox_, gxd = pidx == 2 ? dz_ : dx_;
        const float gyo = pidx == 0 ? oy_ : (pidx == 1 ? oz_ : ox_), gyd = pidx == 0 ? dy_ : (pidx == 1 ? dz_ : dx_);
        const float hx = 0.5f * PW * d.box_scale, hy = 0.5f * PH * d.box_scale;
        ax = fmaf(gxo, hx, 0.5f * (PW - 1));
        bx = gxd * hx;
        ay = fmaf(gyo, hy, 0.5f * (PH - 1));
        by = gyd * hy;
      }

      // ---- coarse depths
      for (int s = lane; s < S; s += 32)
        dep[s] = __ldg(p.lin + s) + __ldg(p.jitter + (size_t)ray * S + s) * d.delta;
      __syncwarp();

      // ---- G: gather the 8 samples [s0, s0 + 8) of this ray into rows 8*warp.. of layer-1 A tile `abuf`
      auto gather = [&](int s0, int s_end, uint8_t* abuf) {
        if (lane < 24) {
          // one (sample, plane) per lane: byte offset of the 2x2 texel block (clamped inside the plane) and its four
          // bilinear weights / 3; taps outside the plane carry weight 0 (grid_sample padding_mode='zeros')
          const int sl = lane & 7, pidx = lane >> 3;
          const int s = s0 + sl;
          const bool sv = rvalid && s < s_end;
          const float t = dep[sv ? s : s0];
          float ix = fmaf(t, bx, ax), iy = fmaf(t, by, ay);
          ix = fminf(fmaxf(ix, -2.f), (float)PW + 1.f);          // far-away samples: still all four taps outside
          iy = fminf(fmaxf(iy, -2.f), (float)PH + 1.f);
          const float fx0 = floorf(ix), fy0 = floorf(iy);
          const int x0 = (int)fx0, y0 = (int)fy0;
          const float third = sv ? (1.f / 3.f) : 0.f;
          const float wr = ix - fx0, wb = (iy - fy0) * third;
          const float wl = 1.f - wr, wt = third - wb;
          // block origin clamped to [0, PW-2]: a tap that slid to the other slot keeps its weight, the slot that
          // now names an interior texel the sample does not touch gets 0
          const int xb = min(max(x0, 0), PW - 2), yb = min(max(y0, 0), PH - 2);
          const float wx0 = x0 == xb ? wl : (x0 + 1 == xb ? wr : 0.f);       // weight of texel xb
          const float wx1 = x0 == xb ? wr : (x0 == xb + 1 ? wl : 0.f);       // weight of texel xb + 1
          const float wy0 = y0 == yb ? wt : (y0 + 1 == yb ? wb : 0.f);
          const float wy1 = y0 == yb ? wb : (y0 == yb + 1 ? wt : 0.f);
          uint8_t* rec = taps + sl * RT_REC;
          *reinterpret_cast<uint32_t*>(rec + pidx * 4) = (uint32_t)(yb * row_b + xb * TEXEL_B + pidx * (RC * 4));
          const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
          *reinterpret_cast<float4*>(rec + 16 + pidx * 32) = make_float4(w00, w00, w10, w10);
          *reinterpret_cast<float4*>(rec + 32 + pidx * 32) = make_float4(w01, w01, w11, w11);
        }
        __syncwarp();
        const int sq = lane >> 3, cg = lane & 7;
#pragma unroll 1
        for (int q = 0; q < RT_TILE; q += 4) {
          const int sl = q + sq;
          const uint8_t* rec = taps + sl * RT_REC;
          const uint4 off = *reinterpret_cast<const uint4*>(rec);
          const uint32_t offs[3] = {off.x, off.y, off.z};
          ulonglong2 v[12];                          // 12 taps x 4 channels as fp32 pairs
#pragma unroll
          for (int pi = 0; pi < 3; ++pi) {
            const char* b0 = lb + offs[pi];
            v[4 * pi + 0] = __ldg(reinterpret_cast<const ulonglong2*>(b0));
            v[4 * pi + 1] = __ldg(reinterpret_cast<const ulonglong2*>(b0 + TEXEL_B));
            v[4 * pi + 2] = __ldg(reinterpret_cast<const ulonglong2*>(b0 + row_b));
            v[4 * pi + 3] = __ldg(reinterpret_cast<const ulonglong2*>(b0 + row_b + TEXEL_B));
          }
          f32x2 a01 = 0ull, a23 = 0ull;              // (+0, +0)
#pragma unroll
          for (int pi = 0; pi < 3; ++pi) {
            const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(rec + 16 + pi * 32);      // (w00, w00), (w10, w10)
            const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(rec + 32 + pi * 32);      // (w01, w01), (w11, w11)
            const f32x2 ww[4] = {wa.x, wa.y, wb.x, wb.y};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              a01 = fma2(ww[k], v[4 * pi + k].x, a01);
              a23 = fma2(ww[k], v[4 * pi + k].y, a23);
            }
          }
          uint2 hi, lo;
          split_pair2(a01, hi.x, lo.x);
          split_pair2(a23, hi.y, lo.y);
          const int row = warp * RT_TILE + sl;
          uint8_t* rp = abuf + row * 128 + (cg & 1) * 8;
          *reinterpret_cast<uint2*>(rp + ((((cg >> 1)) ^ (row & 7)) << 4)) = hi;       // channels 4cg.. of the hi half
          *reinterpret_cast<uint2*>(rp + (((4 + (cg >> 1)) ^ (row & 7)) << 4)) = lo;   // ... of the lo half
        }
        fence_async_smem();
        __syncwarp();
      };

      // ---- E1: hidden layer of tile g: D1 -> softplus -> split -> layer-2 A operand
      auto epi1 = [&](uint32_t g) {
        mbar_wait_sleep(&mma1_done[g & 1], (g >> 1) & 1);
        // A2 buffer g & 1 was last read by MMA2(g-2): by columns every worker has seen that in its own E2(g-2); by
        // tiles that E2 belonged to another group of warps
        if (ESPLIT && g >= 2) mbar_wait_sleep(&mma2_done[g & 1], ((g - 2) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
        for (int ch = ESPLIT ? 0 : cs; ch < (ESPLIT ? 4 : cs + 1); ++ch) {      // 16 hidden columns at a time
          float v[16];
          tmem_ld16(trow + (g & 1) * 64 + ch * 16, v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {              // biases: broadcast 16 B reads
            const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(b0s + 16 * ch + 4 * j);
            split_pair2(softplus2_log2(add2(pk2(v[4 * j], v[4 * j + 1]), b.x)), hi[2 * j], lo[2 * j]);
            split_pair2(softplus2_log2(add2(pk2(v[4 * j + 2], v[4 * j + 3]), b.y)), hi[2 * j + 1], lo[2 * j + 1]);
          }
          // layer-2 A operand straight into TMEM (row = lane, two bf16 per 32-bit column): no shared-memory round trip
          const uint32_t a2 = trow + 256 + (g & 1) * 64 + ch * 8;
          tmem_st8(a2, hi);
          tmem_st8(a2 + 32, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[g & 1]);
      };

      // ---- E2: outputs of tile g (first sample s0): colours -> 16-bit rows, sigma -> density row of the row's ray
      auto epi2 = [&](uint32_t g, int s0, int s_end) {
        mbar_wait_sleep(&mma2_done[g & 1], (g >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tcol = tmem_base + ((uint32_t)(q4 * 32) << 16) + 128 + (g & 1) * 64;
        const int s = s0 + esub;
#pragma unroll 1
        for (int ch = ESPLIT ? 0 : cs; ch < (ESPLIT ? 4 : cs + 1); ++ch) {      // 8 colours at a time
          float c[16];
          if (ch == 3) tmem_ld16(tcol + 24, c);      // colours 24..31, sigma (column 32), zero columns
          else tmem_ld8(tcol + ch * 8, c);
          if (s < s_end) {
            uint32_t qv[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(b1s + 8 * ch + 4 * j);
              qv[2 * j] = sigmoid2_q16(add2(pk2(c[4 * j], c[4 * j + 1]), b.x));
              qv[2 * j + 1] = sigmoid2_q16(add2(pk2(c[4 * j + 2], c[4 * j + 3]), b.y));
            }
            *reinterpret_cast<uint4*>(ecolq + s * 16 + colqx(s, 4 * ch)) = make_uint4(qv[0], qv[1], qv[2], qv[3]);
            if (ch == 3) esig[s] = c[8] + b1s[32];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (ESPLIT) {                                // by tiles: only this group read D2[g & 1]; hand it back explicitly
          __syncwarp();
          if (lane == 0) mbar_arrive(&d2_free[g & 1]);
        }
      };

      // ---- one pass over the samples [s_begin, s_end) of the strip's rays, software-pipelined: per iteration a
      // worker gathers tile it and, when the tile is its group's, finishes the hidden layer of tile it-1 and the outputs
      // of tile it-2, so an MMA (and its commit latency) always has a gather in front of its consumer.
      // (Measured on B200, profiles/r2_render_tc_notes.md: epilogues by tile 0.347 ms, by column slice 0.353 ms; letting
      // half of the warps run the epilogues before the gather 0.39-0.40 ms; one A1 buffer instead of two 0.36 ms.)
      auto run_pass = [&](int s_begin, int s_end) {
        const int nt = (s_end - s_begin + RT_TILE - 1) / RT_TILE;
        const uint32_t g0 = gt;
        auto do_g = [&](int it) {
          if (it < nt) {
            const uint32_t g = g0 + it;
            if (g >= RT_NA1) mbar_wait_sleep(&mma1_done[(g - RT_NA1) & 1], ((g - RT_NA1) >> 1) & 1);   // MMA1 has read this A1 buffer
            gather(s_begin + it * RT_TILE, s_end, a1 + (g % RT_NA1) * RT_A);
            if (lane == 0) mbar_arrive(&a1_full[g % RT_NA1]);
          }
        };
        // ESPLIT: the epilogues of tile g belong to the four warps with g % 4 == cs (all 64 / 33 columns of their rows);
        // else every warp takes its 16 / 8-column slice of every tile
        auto do_e1 = [&](int it) { if (it >= 1 && it <= nt && (!ESPLIT || ((g0 + it - 1) & 3) == (uint32_t)cs)) epi1(g0 + it - 1); };
        auto do_e2 = [&](int it) {
          if (it >= 2 && (!ESPLIT || ((g0 + it - 2) & 3) == (uint32_t)cs)) epi2(g0 + it - 2, s_begin + (it - 2) * RT_TILE, s_end);
        };
        for (int it = 0; it < nt + 2; ++it) { do_g(it); do_e1(it); do_e2(it); }
        gt = g0 + nt;
        worker_bar();                                // every ray's densities / colours of this pass are in place
      };

      run_pass(0, S);

      if (SF > 0) {
        if (rvalid) {
          // ---- coarse march (weights only) -> smoothed pdf -> inverse-CDF fine depths
          float* wts = sdep;                           // sdep is not live until the sort
          march_weights(S - 1, lane, wts, [&](int k) {
            float sm = softplus_t(0.5f * (sig[k] + sig[k + 1]) - 1.f);
            return 1.f - expf(-(sm * (dep[k + 1] - dep[k])));
          });
          __syncwarp();
          const int NB = S - 3;  // number of pdf bins actually used (upstream: weights[:, 1:-1])
          float psum = 0.f;
          for (int j = lane; j < NB; j += 32) {
            const int jj = j + 1;
            float m0 = fmaxf(wts[jj - 1], wts[jj]);
            float m1 = jj + 1 <= S - 2 ? fmaxf(wts[jj], wts[jj + 1]) : wts[jj];
            float a = 0.5f * (m0 + m1) + 0.01f + 1e-5f;
            cdf[1 + j] = a;
            psum += a;
          }
          for (int j = lane; j < S - 1; j += 32) zmid[j] = 0.5f * (dep[j] + dep[j + 1]);
          psum = warp_sum(psum);
          __syncwarp();
          {
            float carry = 0.f;
            for (int base = 0; base < NB; base += 32) {
              int j = base + lane;
              float v = j < NB ? cdf[1 + j] / psum : 0.f;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
              }
              if (j < NB) cdf[1 + j] = carry + v;
              carry += __shfl_sync(0xffffffffu, v, 31);
            }
            if (lane == 0) cdf[0] = 0.f;
          }
          __syncwarp();
          for (int k = lane; k < SF; k += 32) {
            const float u = __ldg(p.u_fine + (size_t)ray * SF + k);
            // searchsorted(cdf[0..NB], u, right=True) = #{cdf[i] <= u}
            const int ind = searchsorted_right(cdf, NB + 1, u);
            const int lo = max(ind - 1, 0), hi = min(ind, NB);
            const float c0 = cdf[lo], c1 = cdf[hi];
            float den = c1 - c0;
            if (den < 1e-5f) den = 1.f;
            const float z0 = zmid[lo], z1 = zmid[hi];
            dep[S + k] = z0 + (u - c0) / den * (z1 - z0);
            if (p.inds) {
              p.inds[(size_t)ray * SF + k] = ind;
              p.below[(size_t)ray * SF + k] = lo;
              p.above[(size_t)ray * SF + k] = hi;
            }
          }
        } else {
          for (int k = lane; k < SF; k += 32) dep[S + k] = dep[0];
        }
        __syncwarp();
        run_pass(S, T);
      }
      if (!rvalid) continue;                         // (warp-uniform; the barriers of this strip are behind us)

      if (SF > 0) {
        // ---- stable rank sort of the T depths (coarse first, as torch.cat + sort sees them)
        auto rank_sort = [&](auto ne_tag) {
          constexpr int NE = decltype(ne_tag)::value;       // elements per lane = ceil(T / 32)
          float de[NE];
          int rk[NE];
          stable_ranks<NE>(dep, T, lane, de, rk, S);
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int i = lane + 32 * e;
            if (i < T) {
              order[rk[e]] = (uint8_t)i;
              sdep[rk[e]] = de[e];
              ssig[rk[e]] = sig[i];
            }
          }
        };
        if (T <= 32) rank_sort(std::integral_constant<int, 1>{});
        else if (T <= 64) rank_sort(std::integral_constant<int, 2>{});
        else if (T <= 96) rank_sort(std::integral_constant<int, 3>{});
        else rank_sort(std::integral_constant<int, 4>{});
      } else {
        for (int i = lane; i < T; i += 32) {
          order[i] = (uint8_t)i;
          sdep[i] = dep[i];
          ssig[i] = sig[i];
        }
      }
      __syncwarp();

      // ---- final march, depth, composite (lane = channel)
      float* wts = dep;                                // unsorted depths are dead after the sort
      const float wtot = march_weights(T - 1, lane, wts, [&](int k) {
        float sm = softplus_t(0.5f * (ssig[k] + ssig[k + 1]) - 1.f);
        return 1.f - expf(-(sm * (sdep[k + 1] - sdep[k])));
      });
      __syncwarp();
      float dacc = 0.f;
      for (int k = lane; k < T - 1; k += 32) dacc = fmaf(wts[k], 0.5f * (sdep[k] + sdep[k + 1]), dacc);
      dacc = warp_sum(dacc);
      // sum_k w_k (c_k + c_{k+1})/2 over the sorted samples == sum_j omega_j c_j in storage order with
      // omega(order[k]) = (w_{k-1} + w_k)/2; omega overwrites sig[]
      float osum = 0.f;
      for (int k = lane; k < T; k += 32) {
        const float om = 0.5f * ((k > 0 ? wts[k - 1] : 0.f) + (k < T - 1 ? wts[k] : 0.f));
        sig[order[k]] = om;
        osum += om;
      }
      osum = warp_sum(osum);
      __syncwarp();
      float acc = 0.f;
      {
        const int cw = lane >> 1, sh = (lane & 1) * 16;
        auto qcol = [&](int crow) { return (float)((colq[crow * 16 + colqx(crow, cw)] >> sh) & 0xffffu); };
        int j = 0;
        for (; j + 4 <= T; j += 4) {
          const float4 om = *reinterpret_cast<const float4*>(sig + j);
          acc = fmaf(om.x, qcol(j), acc);
          acc = fmaf(om.y, qcol(j + 1), acc);
          acc = fmaf(om.z, qcol(j + 2), acc);
          acc = fmaf(om.w, qcol(j + 3), acc);
        }
        for (; j < T; ++j) acc = fmaf(sig[j], qcol(j), acc);
      }
      acc = fmaf(acc, QSTEP, -0.001f * osum);
      p.feat[(size_t)ray * RC + lane] = acc * 2.f - 1.f;
      if (lane == 0) {
        float dv = dacc / wtot;
        if (isnan(dv)) dv = INFINITY;
        dv = fminf(fmaxf(dv, __ldg(p.depth_range)), __ldg(p.depth_range + 1));
        p.depth[ray] = dv;
        p.wsum[ray] = wtot;
      }
      if (p.sort_idx)
        for (int i = lane; i < T; i += 32) p.sort_idx[(size_t)ray * T + i] = order[i];
      if (p.depths_sorted)
        for (int i = lane; i < T; i += 32) p.depths_sorted[(size_t)ray * T + i] = sdep[i];
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == RT_RAYS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(RT_TMEM_COLS) : "memory");
  }
}

// true when the tcgen05 renderer can take this problem (its shared-memory plan fits); else the caller uses render.cu
bool render_tc_supported(const HfagpRenderDesc& d) {
  return d.s_coarse >= 4 && d.plane_w >= 2 && d.plane_h >= 2 && rt_smem_bytes(d.s_coarse, d.s_fine, 2) <= 227 * 1024;
}

template <int PWC, int NA1, int ES>
static int rt_launch_variant(const RenderParams& p, int blocks, cudaStream_t stream) {
  const size_t smem = rt_smem_bytes(p.d.s_coarse, p.d.s_fine, NA1);
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(render_tc_kernel<PWC, NA1, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  if (smem > 227 * 1024) return fail(HFAGP_E_INVALID, "render_tc: shared memory plan does not fit");
  static const int tw_log2 = [] { const char* e = getenv("HFAGP_RT_TILE_W_LOG2"); const int v = e ? atoi(e) : 2; return v < 0 ? 0 : (v > 4 ? 4 : v); }();
  const int xt = (p.d.res + (1 << tw_log2) - 1) >> tw_log2, th = RT_RAYS >> tw_log2;
  const long long strips = (long long)p.d.batch * xt * ((p.d.res + th - 1) / th);
  if (strips < blocks) blocks = (int)strips;
  render_tc_kernel<PWC, NA1, ES><<<blocks, RT_THREADS, smem, stream>>>(p, tw_log2);
  HFAGP_CHECK_LAUNCH("render_tc_kernel");
  return HFAGP_OK;
}

int render_tc_launch(const RenderParams& p, int sms, cudaStream_t stream) {
  const HfagpRenderDesc& d = p.d;
  const int blocks = sms;                            // clipped to the number of strips by the variant
  if (d.plane_w == 256) return rt_launch_variant<256, 2, 1>(p, blocks, stream);
  return rt_launch_variant<0, 2, 1>(p, blocks, stream);
}

}  // namespace hfagp
