// Convolution weight gradient on tcgen05 (round 2):
//
//   dw[t][co][ci] += scale * sum_{n,my,mx} dz[n][my][mx][co] * x[n][my*s + dy_t][mx*s + dx_t][ci]
//
// as GEMMs D[M = co][N = ci] over K = output pixels.  Both operands are channels-last activations, i.e. their M / N index
// (the channel) is the contiguous one: they are fed to the tensor core as MN-MAJOR operands — the very TMA boxes the forward
// convolution loads ([64 channels] x [8 pixels] x [rows], SWIZZLE_128B: one pixel = one 128 B row, 8 pixels = one 1024 B
// swizzle atom) read with the MN-major canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16 B units — no transposed copy
// of any activation is ever made.  fp32-class accuracy by the usual three split-bf16 terms (hi*hi + lo*hi + hi*lo), fp32
// accumulators in TMEM, one accumulator per tap of the work item's tap group (at stride 1 the taps sharing dx read ONE input
// patch with a (rows + span) halo, exactly like the forward kernel).  K is split over CTAs; partial sums leave through
// 16 B red.global.add.  Replaces the fp32 SIMT wgrad_kernel (bwd.cu) whenever the shapes allow (cin, cout multiples of 64,
// output width >= 8, split-bf16 operands, no per-sample style factors).
//
// Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2-5 epilogue (one per TMEM lane quadrant).
// Reference use: autograd of F.conv2d w.r.t. its weight inside Trainer.gen_update (code/trainer_rgb.py:91).
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>
#include "common.cuh"
#include "tc_common.cuh"

namespace hfagp {

int get_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b0,
            uint32_t b1, uint32_t b2, uint32_t b3, uint32_t estride, int rank);

constexpr int WT_PX = 64;                 // pixels per K chunk: an 8 x 8 tile of the output grid
constexpr int WT_ROW = 128;               // bytes of one pixel row of a 64-channel box
constexpr int WT_ABLK = WT_PX * WT_ROW;   // one 64-channel dz box: 8 KB
constexpr int WT_THREADS = 192;
constexpr int WT_STAGES = 3;

struct WtParams {
  int batch, oh, ow, cin, cout, in_stride;
  int tiles_x, tiles_y;                   // 8 x 8 pixel tiles per sample
  int chunks;                             // batch * tiles_y * tiles_x
  int splits;                             // K splits per (group, co tile, ci tile)
  int co_tiles, ci_tiles, nblk;           // nblk = 64-channel blocks of one N tile (1 or 2)
  int ngroups;
  int gtaps[HFAGP_MAX_TAPS];              // taps of group g: [gstart[g], gstart[g+1])
  int gstart[HFAGP_MAX_TAPS + 1];
  int gdx[HFAGP_MAX_TAPS], gdy0[HFAGP_MAX_TAPS], gspan[HFAGP_MAX_TAPS];
  int trow[HFAGP_MAX_TAPS], twt[HFAGP_MAX_TAPS];   // per (sorted) tap: row offset inside the patch, weight-tap index
  int patch_rows;                         // 8 + max span
  float scale;
  float* dw;
  // modulated convolutions (hfagp_conv2d_wgrad_mod): per-sample style factors.  They commute with the pixel sum, so they
  // are applied to the finished per-sample accumulator: dw[t][co][ci] += scale * rowscale[n][co] * colscale[n][ci] * D_n
  // (a work item then never mixes samples: per_sample = 1 deals the K splits out inside one sample)
  const float* rowscale;                  // [batch][cout] or null  (the style of the dz-side operand)
  const float* colscale;                  // [batch][cin]  or null  (the style of the x-side operand)
  int per_sample;
};

__device__ __forceinline__ void tma_load_4d_wt(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// MN-major, SWIZZLE_128B shared-memory matrix descriptor: 64-channel blocks `lbo` bytes apart, 8-pixel groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32_wt(uint32_t taddr, float* v) {
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

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz_hi, const __grid_constant__ CUtensorMap map_dz_lo,
                const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ WtParams p) {
  extern __shared__ __align__(16) uint8_t wt_smem_raw[];
  uint8_t* smem = wt_smem_raw + ((1024u - (smem_u32(wt_smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bblk = p.patch_rows * 8 * WT_ROW;                  // one 64-channel x patch: (8 + span) rows of 8 pixels
  const int a_bytes = 2 * 2 * WT_ABLK;                         // dz: 2 channel blocks x (hi, lo)
  const int b_bytes = p.nblk * 2 * bblk;                       // x: nblk channel blocks x (hi, lo)
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WT_STAGES * stage_bytes);
  uint64_t* empty = full + WT_STAGES;
  uint64_t* acc_full = empty + WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // ---- work item: (tap group, co tile, ci tile, K split)
  int w = blockIdx.x;
  const int ks = w % p.splits; w /= p.splits;
  int sample = 0;
  if (p.per_sample) { sample = w % p.batch; w /= p.batch; }
  const int cit = w % p.ci_tiles; w /= p.ci_tiles;
  const int cot = w % p.co_tiles; w /= p.co_tiles;
  const int g = w;
  const int t0 = p.gstart[g], ntap = p.gstart[g + 1] - t0;
  const int span_chunks = p.per_sample ? p.tiles_x * p.tiles_y : p.chunks;      // chunks this item's splits divide
  const int c_base = p.per_sample ? sample * span_chunks : 0;
  const int c_begin = c_base + (int)((long long)span_chunks * ks / p.splits);
  const int c_end = c_base + (int)((long long)span_chunks * (ks + 1) / p.splits);
  const int N = p.nblk * 64;
  const uint32_t tmem_cols = 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer
    uint32_t it = 0;
    for (int c = c_begin; c < c_end; ++c, ++it) {
      const int s = it % WT_STAGES;
      mbar_wait(&empty[s], ((it / WT_STAGES) & 1) ^ 1);
      const int n = c / (p.tiles_y * p.tiles_x);
      const int r = c - n * p.tiles_y * p.tiles_x;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      uint8_t* sa = smem + s * stage_bytes;
      uint8_t* sb = sa + a_bytes;
      if (elect_one()) {
        mbar_expect_tx(&full[s], stage_bytes);
        for (int b = 0; b < 2; ++b) {
          tma_load_4d_wt(&map_dz_hi, sa + b * WT_ABLK, &full[s], cot * 128 + b * 64, tx * 8, ty * 8, n);
          tma_load_4d_wt(&map_dz_lo, sa + 2 * WT_ABLK + b * WT_ABLK, &full[s], cot * 128 + b * 64, tx * 8, ty * 8, n);
        }
        const int ax = tx * 8 * p.in_stride + p.gdx[g], ay = ty * 8 * p.in_stride + p.gdy0[g];
        for (int b = 0; b < p.nblk; ++b) {
          tma_load_4d_wt(&map_x_hi, sb + b * bblk, &full[s], cit * N + b * 64, ax, ay, n);
          tma_load_4d_wt(&map_x_lo, sb + p.nblk * bblk + b * bblk, &full[s], cit * N + b * 64, ax, ay, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D_tap[128 co x N ci] += dz^T . x_tap over the chunk's 64 pixels, three split-bf16 terms
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t it = 0;
    for (int c = c_begin; c < c_end; ++c, ++it) {
      const int s = it % WT_STAGES;
      mbar_wait(&full[s], (it / WT_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + s * stage_bytes), sb = sa + a_bytes;
      if (elect_one()) {
        for (int t = 0; t < ntap; ++t) {
          const uint32_t acc = tmem_base + t * N;
          const uint32_t boff = p.trow[t0 + t] * 8 * WT_ROW;            // the tap's first patch row
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t a0 = sa + (term == 1 ? 2 * WT_ABLK : 0);                 // hi, lo, hi
            const uint32_t b0 = sb + boff + (term == 2 ? p.nblk * bblk : 0);        // hi, hi, lo
#pragma unroll
            for (int k = 0; k < WT_PX / 16; ++k)                                   // 16 pixels = two 8-pixel groups per MMA
              umma_bf16(acc, umma_desc_mn(a0 + k * 2048, WT_ABLK), umma_desc_mn(b0 + k * 2048, bblk), idesc,
                        (it | term | k) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else {
    // ===== epilogue: accumulator row = output channel, columns = input channels of the N tile
    const int q = warp & 3;                                    // TMEM lane quadrant of this warp
    const int co = cot * 128 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (c_end > c_begin) {
      const float rs = p.scale * ((p.rowscale && co < p.cout) ? __ldg(p.rowscale + (size_t)sample * p.cout + co) : 1.f);
      const float* cs = p.colscale ? p.colscale + (size_t)sample * p.cin + cit * N : nullptr;
      for (int t = 0; t < ntap; ++t) {
        float* out = p.dw + ((size_t)p.twt[t0 + t] * p.cout + co) * p.cin + cit * N;
        for (int cb = 0; cb < N; cb += 32) {
          float v[32];
          tmem_ld32_wt(tmem_base + ((uint32_t)(q * 32) << 16) + t * N + cb, v);
          if (co < p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 f = make_float4(rs, rs, rs, rs);
              if (cs) {                                   // warp-uniform address: one broadcast 16 B load
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + cb + j));
                f.x *= c4.x; f.y *= c4.y; f.z *= c4.z; f.w *= c4.w;
              }
              atomicAdd(reinterpret_cast<float4*>(out + cb + j),
                        make_float4(v[j] * f.x, v[j + 1] * f.y, v[j + 2] * f.z, v[j + 3] * f.w));
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

bool wgrad_tc_supported(const HfagpConvDesc& d) {
  // cin in whole 64-channel blocks; cout only needs TMA's 16 B row rule (the last dz box is zero-filled past cout)
  if ((d.cin & 63) || (d.cout & 7) || d.cout < 32 || d.ow < 8 || d.oh < 8) return false;
  if ((long long)d.batch * d.oh * d.ow < 1024) return false;            // too little K to fill the pipe: SIMT kernel
  if (d.in_stride != 1 && d.in_stride != 2) return false;
  return true;
}

int wgrad_tc_launch(const HfagpConvDesc& d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dz_hi,
                    const uint16_t* dz_lo, const float* xscale, const float* dzscale, float scale, float* dw, int num_sms,
                    cudaStream_t stream) {
  WtParams p = {};
  p.batch = d.batch; p.oh = d.oh; p.ow = d.ow; p.cin = d.cin; p.cout = d.cout; p.in_stride = d.in_stride;
  p.tiles_x = cdiv(d.ow, 8); p.tiles_y = cdiv(d.oh, 8);
  p.chunks = d.batch * p.tiles_x * p.tiles_y;
  p.nblk = (d.cin % 128 == 0) ? 2 : 1;
  p.co_tiles = cdiv(d.cout, 128);
  p.ci_tiles = d.cin / (p.nblk * 64);
  p.scale = scale; p.dw = dw;
  p.colscale = xscale; p.rowscale = dzscale;
  p.per_sample = (xscale || dzscale) ? 1 : 0;
  // tap groups: at stride 1 the taps sharing dx read one patch (consecutive dy, at most 512 / N accumulators)
  int order[HFAGP_MAX_TAPS];
  for (int t = 0; t < d.ntaps; ++t) order[t] = t;
  for (int i = 1; i < d.ntaps; ++i)
    for (int j = i; j > 0; --j) {
      const int a = order[j - 1], b = order[j];
      if (d.dx[a] > d.dx[b] || (d.dx[a] == d.dx[b] && d.dy[a] > d.dy[b])) { order[j - 1] = b; order[j] = a; } else break;
    }
  const int max_acc = 512 / (p.nblk * 64);
  int max_span = 0, i = 0;
  p.ngroups = 0;
  while (i < d.ntaps) {
    int j = i;
    if (d.in_stride == 1)
      while (j + 1 < d.ntaps && d.dx[order[j + 1]] == d.dx[order[i]] && j + 1 - i < max_acc &&
             d.dy[order[j + 1]] - d.dy[order[i]] <= 4) ++j;
    const int g = p.ngroups++;
    p.gstart[g] = i;
    p.gdx[g] = d.dx[order[i]];
    p.gdy0[g] = d.dy[order[i]];
    p.gspan[g] = d.dy[order[j]] - d.dy[order[i]];
    if (p.gspan[g] > max_span) max_span = p.gspan[g];
    for (int k = i; k <= j; ++k) { p.trow[k] = d.dy[order[k]] - d.dy[order[i]]; p.twt[k] = d.wtap[order[k]]; }
    i = j + 1;
  }
  p.gstart[p.ngroups] = d.ntaps;
  p.patch_rows = 8 + max_span;
  const int items = p.ngroups * p.co_tiles * p.ci_tiles * (p.per_sample ? d.batch : 1);
  const int span_chunks = p.per_sample ? p.tiles_x * p.tiles_y : p.chunks;
  int splits = cdiv(2 * num_sms, items);
  if (splits > span_chunks / 4) splits = span_chunks / 4;
  if (splits < 1) splits = 1;
  p.splits = splits;
  const size_t stage = 2 * 2 * WT_ABLK + (size_t)p.nblk * 2 * p.patch_rows * 8 * WT_ROW;
  const size_t smem = 1024 + WT_STAGES * stage + 256;
  if (smem > 227 * 1024) return fail(HFAGP_E_INVALID, "wgrad_tc: shared-memory plan does not fit");
  CUtensorMap mz_hi, mz_lo, mx_hi, mx_lo;
  int rc;
  const uint32_t es = (uint32_t)d.in_stride;
  const uint32_t bw = 8 * es - (es - 1), bh = (uint32_t)p.patch_rows * es - (es - 1);
  if ((rc = get_map(&mz_hi, dz_hi, d.cout, d.ow, d.oh, d.batch, 64, 8, 8, 1, 1, 4))) return rc;
  if ((rc = get_map(&mz_lo, dz_lo, d.cout, d.ow, d.oh, d.batch, 64, 8, 8, 1, 1, 4))) return rc;
  if ((rc = get_map(&mx_hi, x_hi, d.cin, d.in_w, d.in_h, d.batch, 64, bw, bh, 1, es, 4))) return rc;
  if ((rc = get_map(&mx_lo, x_lo, d.cin, d.in_w, d.in_h, d.batch, 64, bw, bh, 1, es, 4))) return rc;
  static std::atomic<uint64_t> attr_done{0};
  HFAGP_CUDA(per_device_once(attr_done, [] { return cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  wgrad_tc_kernel<<<items * splits, WT_THREADS, smem, stream>>>(mz_hi, mz_lo, mx_hi, mx_lo, p);
  HFAGP_CHECK_LAUNCH("wgrad_tc_kernel");
  return HFAGP_OK;
}

}  // namespace hfagp
