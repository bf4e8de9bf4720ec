// Orthonormal basis of the latent subspace: the reduced QR factorisation of A = (bases + eps)^T  [M x K], M >> K
// (14*512 x 50), forward and backward, without a LAPACK call.
//
// Replaces: torch.qr(bases.T + 1e-8) in get_latent (code/networks/headnerf.py:92, :187, :247) and its autograd backward.
// cuSOLVER's geqrf runs the 50 Householder steps of this tall-skinny matrix as a 0.45 ms latency chain (0.77 ms with its
// set-up inside a CUDA graph) — on the critical path of every training step that has no encoder to hide it behind.
//
// Algorithm: CholeskyQR2 (two rounds of  G = A^T A,  R = chol(G),  A <- A R^-1; Gram matrices accumulated in fp64, the
// first round's K x K work in fp32 — it is a preconditioner — the second round's in fp64) followed by "Householder reconstruction" of the signs: LAPACK's R has diag(R)_j = -sign(x_j) |x_j| where x_j
// is the j-th pivot met by the Householder sweep; that sign sequence equals the one chosen by a sign-picking LU of the
// top K x K block of the orthonormal factor (Ballard et al., "Reconstructing Householder vectors from tall-skinny QR",
// 2014).  The result is LAPACK's (Q, R) up to rounding: measured 1e-7 max |dQ| against torch.linalg.qr on the CPU, and
// closer to the fp64 factor than LAPACK's fp32 result.  Needs cond(A) < ~1e3 (the bases are a randn matrix, cond ~ 1.2);
// a pivot below the floor raises the `info` flag.
//
// Backward (Q only; R is not used downstream):  gA = (gQ + Q Y) R^-T,  Y = X + X^T - diag(X),  X = triu(-Q^T gQ)
// (the m >= n case of torch's linalg_qr_backward), with R^-1 kept from the forward so the triangular solve is a product.
#include "common.cuh"

namespace hfagp {

constexpr int QR_MAXK = 64;
constexpr int QR_P = 68;        // shared-memory row pitch (floats): 16-byte aligned rows, 4 banks of skew per row
constexpr int QR_GROWS = 64;    // rows of A per chunk, Gram kernel
constexpr int QR_AROWS = 32;    // rows of A per chunk, apply kernel
constexpr int QR_CPC = 2;       // chunks per CTA (the Gram accumulators stay in registers across them: one partial tile per CTA)

// element (m, k) of a tall matrix stored with arbitrary strides (bases layout: sm = 1, sk = M; Q layout: sm = K, sk = 1)
struct TallMat {
  const float* p;
  long long sm, sk;
  float eps;
};

// rows [m0, m0 + rows) -> s[r][0..64): columns >= K and rows >= `rows` are zero
__device__ __forceinline__ void load_rows(const TallMat& a, int m0, int rows, int K, float (*s)[QR_P], int maxrows) {
  // (a block is latency bound: keep several independent loads in flight per thread)
  if (a.sm == 1) {          // consecutive threads along m
#pragma unroll 8
    for (int idx = threadIdx.x; idx < maxrows * QR_MAXK; idx += 256) {
      const int r = idx % maxrows, k = idx / maxrows;
      s[r][k] = (r < rows && k < K) ? __ldg(a.p + (long long)(m0 + r) + k * a.sk) + a.eps : 0.f;
    }
  } else {                  // consecutive threads along k
#pragma unroll 8
    for (int idx = threadIdx.x; idx < maxrows * QR_MAXK; idx += 256) {
      const int k = idx % QR_MAXK, r = idx / QR_MAXK;
      s[r][k] = (r < rows && k < K) ? __ldg(a.p + (long long)(m0 + r) * a.sm + k * a.sk) + a.eps : 0.f;
    }
  }
}

// acc[a][b] += sum_r sa[r][4 ti + a] sb[r][4 tj + b]   (a thread owns a 4 x 4 tile of the K x K product)
__device__ __forceinline__ void gram_tile(const float (*sa)[QR_P], const float (*sb)[QR_P], int nrows, int ti, int tj,
                                          float (&acc)[4][4]) {
#pragma unroll 4
  for (int r = 0; r < nrows; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(&sa[r][4 * ti]);
    const float4 b = *reinterpret_cast<const float4*>(&sb[r][4 * tj]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
// this CTA's partial product -> G (fp64 atomics: the sum's order is free, but its error is far below the fp32 result's
// last bit)
__device__ __forceinline__ void gram_flush(double* G, int K, int ti, int tj, const float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * ti + i < K && 4 * tj + j < K) atomicAdd(G + (4 * ti + i) * K + 4 * tj + j, (double)acc[i][j]);
}
// 1 / x in fp64 without the division routine: fp32 seed + two Newton steps (relative error ~1e-15)
__device__ __forceinline__ float rcp_nr(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) { return rsqrt(x); }
// relative pivot floor: below it the factorisation is declared ill conditioned (the fp32 first pass is only a
// preconditioner, but the second pass can only repair |Q1^T Q1 - I| < ~0.1, i.e. cond(A) up to ~1e3)
__device__ __forceinline__ double tiny_rel(double) { return 1e-9; }
__device__ __forceinline__ float tiny_rel(float) { return 1e-6f; }
__device__ __forceinline__ double rcp_nr(double x) {
  double r = (double)(1.0f / (float)x);
  r = r * (2.0 - x * r);
  return r * (2.0 - x * r);
}

// G[K][K] (fp64, zeroed before the launch) += X^T Y over this CTA's rows
__global__ void __launch_bounds__(256) qr_gram_kernel(int M, int K, int cpc, TallMat x, TallMat y, int same, double* __restrict__ G) {
  __shared__ __align__(16) float sx[QR_GROWS][QR_P], sy[QR_GROWS][QR_P];
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  float acc[4][4] = {};
  for (int c = 0; c < cpc; ++c) {
    const int m0 = (blockIdx.x * cpc + c) * QR_GROWS;
    if (m0 >= M) break;
    const int rows = min(QR_GROWS, M - m0);
    if (c) __syncthreads();
    load_rows(x, m0, rows, K, sx, QR_GROWS);
    if (!same) load_rows(y, m0, rows, K, sy, QR_GROWS);
    __syncthreads();
    gram_tile(sx, same ? sx : sy, QR_GROWS, ti, tj, acc);
  }
  gram_flush(G, K, ti, tj, acc);
}

// out = X M1 (+ Y M2) for this CTA's rows; optionally G (fp64) += out^T out
__global__ void __launch_bounds__(256) qr_apply_kernel(int M, int K, int cpc, TallMat x, const float* __restrict__ M1, TallMat y,
                                                       const float* __restrict__ M2, float* __restrict__ out, long long os_m,
                                                       long long os_k, double* __restrict__ G) {
  __shared__ __align__(16) float sx[QR_AROWS][QR_P], sm[QR_MAXK][QR_P], so[QR_AROWS][QR_P];
  const int tr = threadIdx.x >> 4, tj = threadIdx.x & 15;      // output rows 2 tr, 2 tr + 1; columns 4 tj .. 4 tj + 3
  float gacc[4][4] = {};
  for (int c = 0; c < cpc; ++c) {
    const int m0 = (blockIdx.x * cpc + c) * QR_AROWS;
    if (m0 >= M) break;
    const int rows = min(QR_AROWS, M - m0);
    float acc[2][4] = {};
    for (int term = 0; term < (M2 ? 2 : 1); ++term) {
      __syncthreads();
      load_rows(term ? y : x, m0, rows, K, sx, QR_AROWS);
      if (term || c == 0 || M2) {                   // the single-term form keeps its matrix across chunks
        const float* Mt = term ? M2 : M1;
#pragma unroll 8
        for (int idx = threadIdx.x; idx < QR_MAXK * QR_MAXK; idx += 256) {
          const int k = idx >> 6, j = idx & 63;
          sm[k][j] = (k < K && j < K) ? __ldg(Mt + k * K + j) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const float a0 = sx[2 * tr][k], a1 = sx[2 * tr + 1][k];
        const float4 b = *reinterpret_cast<const float4*>(&sm[k][4 * tj]);
        acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
        acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
        acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
        acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
      }
    }
    *reinterpret_cast<float4*>(&so[2 * tr][4 * tj]) = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
    *reinterpret_cast<float4*>(&so[2 * tr + 1][4 * tj]) = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
    __syncthreads();
    if (os_m == 1) {
      for (int idx = threadIdx.x; idx < QR_AROWS * K; idx += blockDim.x) {
        const int r = idx % QR_AROWS, k = idx / QR_AROWS;
        if (r < rows) out[(long long)(m0 + r) + k * os_k] = so[r][k];
      }
    } else {
      for (int idx = threadIdx.x; idx < QR_AROWS * QR_MAXK; idx += blockDim.x) {
        const int k = idx & 63, r = idx >> 6;
        if (r < rows && k < K) out[(long long)(m0 + r) * os_m + k * os_k] = so[r][k];
      }
    }
    if (G) gram_tile(so, so, QR_AROWS, threadIdx.x >> 4, tj, gacc);        // rows >= `rows` of so are zero
  }
  if (G) gram_flush(G, K, threadIdx.x >> 4, tj, gacc);
}

// ---- the K x K work, one CTA of 16 warps, fp64 in shared memory
// upper Cholesky factor in place (G = R^T R; the strict lower triangle is left untouched) then X = R^-1 (upper).
// One barrier per elimination step: row j stays unscaled while it is used (the update takes G[j][i] G[j][l] / G[j][j]),
// all rows are scaled by 1 / sqrt(pivot) at the end.
template <typename F>
__device__ void chol_inv(int K, F* G, F* X, int* info) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, t = threadIdx.x, nt = blockDim.x, nw = blockDim.x >> 5;
  __shared__ F diag0[QR_MAXK], rd[QR_MAXK];
  __syncthreads();
  for (int j = t; j < K; j += nt) diag0[j] = G[j * K + j];
  for (int j = 0; j < K; ++j) {
    __syncthreads();
    const F piv = G[j * K + j];
    // a pivot that lost 9+ digits against its column's squared norm: cond(A) beyond what CholeskyQR2 repairs
    if (!(piv > tiny_rel(F(0)) * diag0[j]) && t == 0) atomicOr(info, 1);
    const F pinv = rcp_nr(piv > F(0) ? piv : F(1e-30));
    for (int i = j + 1 + ty; i < K; i += nw) {
      const F f = G[j * K + i] * pinv;
      for (int l = i + tx; l < K; l += 32) G[i * K + l] -= f * G[j * K + l];
    }
  }
  __syncthreads();
  for (int j = ty; j < K; j += nw) {
    const F piv = G[j * K + j];
    const F d = rsqrt_t(piv > F(0) ? piv : F(1e-30));
    __syncwarp();
    for (int l = j + tx; l < K; l += 32) G[j * K + l] *= d;
    if (tx == 0) rd[j] = d;                          // = 1 / R[j][j]
  }
  __syncthreads();
  // back substitution R x = e_c: every column at once, 8 lanes per column share each step's dot product
  for (int c0 = 0; c0 < K; c0 += 4 * nw) {
    const int c = c0 + 4 * ty + (tx >> 3), lp = tx & 7;
    const bool on = c < K;
    if (on) {
      for (int i = c + 1 + lp; i < K; i += 8) X[i * K + c] = F(0);
      if (lp == 0) X[c * K + c] = rd[c];
    }
    __syncwarp();
    const int cmax = min(K - 1, c0 + 4 * ty + 3);     // the warp's longest column: uniform trip count
    for (int i = cmax - 1; i >= 0; --i) {
      F s = F(0);
      if (on && i < c)
        for (int l = i + 1 + lp; l <= c; l += 8) s += G[i * K + l] * X[l * K + c];
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (on && i < c && lp == 0) X[i * K + c] = -s * rd[i];
      __syncwarp();
    }
  }
  __syncthreads();
}

// mode 0: G1 -> Rinv1.  The first pass is a preconditioner (whatever upper-triangular Rinv1 it produces, the second pass
//         measures Q1 = A Rinv1 and corrects it), so its k x k work runs in fp32.
// mode 1: G2 = Q1^T Q1 = I + E, Rinv1, top K rows of Q1 -> Rinv2, signs; Rinv2s = Rinv2 S (fp32, apply pass 2);
//         rinv_total = Rinv1 Rinv2 S (fp32).  |E| is ~1e-6, so R2 = chol(I + E) = I + U + O(E^2) with U = triu(E, 1) + diag(E) / 2
//         and Rinv2 = I - U: no factorisation; only when max |E| > 1e-4 (an ill-conditioned basis) the fp64 Cholesky runs.
__global__ void __launch_bounds__(512) qr_small_fwd_kernel(int K, int mode, int square, const double* __restrict__ Gin, float* __restrict__ rinv_f32,
                                                           float* __restrict__ rinv1, const float* __restrict__ q1_top,
                                                           float* __restrict__ rinv_total, int* __restrict__ info) {
  extern __shared__ double qs[];
  __shared__ float sgn[QR_MAXK];
  const int t = threadIdx.x, nt = blockDim.x;
  if (mode == 0) {
    float* G = reinterpret_cast<float*>(qs);
    float* X = G + K * K;
    for (int o = t; o < K * K; o += nt) G[o] = (float)Gin[o];
    chol_inv<float>(K, G, X, info);
    for (int o = t; o < K * K; o += nt) { rinv_f32[o] = X[o]; rinv1[o] = X[o]; }
    return;
  }
  double* G = qs;
  double* X = G + K * K;
  float* T = reinterpret_cast<float*>(X + K * K);
  int big = 0;
  for (int o = t; o < K * K; o += nt) {
    const int i = o / K, j = o - i * K;
    const double g = Gin[o], e = g - (i == j ? 1.0 : 0.0);
    G[o] = g;
    big |= fabs(e) > 1e-4;
    X[o] = i == j ? 1.0 - 0.5 * e : (i < j ? -e : 0.0);
  }
  if (__syncthreads_or(big)) chol_inv<double>(K, G, X, info);
  // T = Q1[:K] Rinv2: the top block of the orthonormal factor
  for (int o = t; o < K * K; o += nt) {
    const int i = o / K, j = o - i * K;
    double s = 0.0;
    for (int l = 0; l <= j; ++l) s += (double)__ldg(q1_top + i * K + l) * X[l * K + j];
    T[o] = (float)s;
  }
  // sign-picking LU: S_jj = -sgn(pivot), pivot -= S_jj (|pivot| >= 1 afterwards: no pivoting needed).  One barrier per
  // step: row j and column j are only read at step j (the modified pivot stays in a register)
  for (int j = 0; j < K; ++j) {
    __syncthreads();
    const float piv = T[j * K + j];
    // (a square matrix: LAPACK applies no reflection to the last, one-element column and R keeps that element's sign)
    const float sg = (piv >= 0.f) != (square && j == K - 1) ? -1.f : 1.f;
    if (t == 0) sgn[j] = sg;
    const float pinv = __frcp_rn(piv - sg);
    for (int i = j + 1 + (t >> 5); i < K; i += (nt >> 5)) {
      const float lij = T[i * K + j] * pinv;
      for (int l = j + 1 + (t & 31); l < K; l += 32) T[i * K + l] -= lij * T[j * K + l];
    }
  }
  __syncthreads();
  for (int o = t; o < K * K; o += nt) {
    const int i = o / K, j = o - i * K;
    rinv_f32[o] = (float)(X[o] * sgn[j]);
    double s = 0.0;
    for (int l = i; l <= j; ++l) s += (double)rinv1[i * K + l] * X[l * K + j];        // both upper triangular
    rinv_total[o] = (float)(s * sgn[j]);
  }
}

// backward: P = Q^T gQ (fp64) and R^-1 -> M1 = R^-T, M2 = Y R^-T with Y = X + X^T - diag(X), X = triu(-P)
__global__ void __launch_bounds__(512) qr_small_bwd_kernel(int K, const double* __restrict__ P, const float* __restrict__ rinv,
                                                           float* __restrict__ M1, float* __restrict__ M2) {
  extern __shared__ double qs[];
  double* Y = qs;
  double* Ri = Y + K * K;
  const int t = threadIdx.x, nt = blockDim.x;
  for (int o = t; o < K * K; o += nt) {
    const int i = o / K, j = o - i * K;
    Y[o] = -(i <= j ? P[i * K + j] : P[j * K + i]);
    Ri[o] = (double)__ldg(rinv + o);
  }
  __syncthreads();
  for (int o = t; o < K * K; o += nt) {
    const int i = o / K, j = o - i * K;
    M1[o] = (float)Ri[j * K + i];
    double s = 0.0;
    for (int l = j; l < K; ++l) s += Y[i * K + l] * Ri[j * K + l];     // R^-1 upper: Rinv[j][l] = 0 for l < j
    M2[o] = (float)s;
  }
}

static size_t qr_align(size_t v) { return (v + 255) & ~(size_t)255; }
// workspace layout: G1, G2 / P [K*K] f64 | rinv1 [K*K] f64 | rinv_a, rinv_b [K*K] f32 | info | Q1 [M*K] f32
struct QrWs {
  double *g1, *g2;
  float *rinv1, *ra, *rb, *q1;
  int* info;
};
static QrWs qr_ws(void* ws, int K) {
  char* b = static_cast<char*>(ws);
  const size_t kk8 = qr_align((size_t)K * K * 8), kk4 = qr_align((size_t)K * K * 4);
  QrWs w;
  w.g1 = reinterpret_cast<double*>(b);
  w.g2 = reinterpret_cast<double*>(b + kk8);
  w.rinv1 = reinterpret_cast<float*>(b + 2 * kk8);
  w.ra = reinterpret_cast<float*>(b + 3 * kk8);
  w.rb = reinterpret_cast<float*>(b + 3 * kk8 + kk4);
  w.info = reinterpret_cast<int*>(b + 3 * kk8 + 2 * kk4);
  w.q1 = reinterpret_cast<float*>(b + 3 * kk8 + 2 * kk4 + 256);
  return w;
}

}  // namespace hfagp

using namespace hfagp;

extern "C" size_t hfagp_basis_qr_workspace_bytes(int k, int m) {
  if (k <= 0 || m <= 0) return 0;
  return 3 * qr_align((size_t)k * k * 8) + 2 * qr_align((size_t)k * k * 4) + 256 + qr_align((size_t)m * k * 4);
}

static int qr_small_attr() {
  static std::atomic<uint64_t> done_f{0}, done_b{0};
  HFAGP_CUDA(per_device_once(done_f, [] { return cudaFuncSetAttribute(qr_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * QR_MAXK * QR_MAXK * 8); }));
  HFAGP_CUDA(per_device_once(done_b, [] { return cudaFuncSetAttribute(qr_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * QR_MAXK * QR_MAXK * 8); }));
  return HFAGP_OK;
}

extern "C" int hfagp_basis_qr_fwd(int k, int m, const float* bases, float eps, float* q, float* rinv, void* workspace,
                                  int deterministic, void* stream) {
  HFAGP_CHECK_ARG(k >= 1 && k <= QR_MAXK && m >= k, "basis_qr: need 1 <= k <= 64 and m >= k (got k=%d m=%d)", k, m);
  HFAGP_CHECK_ARG(bases && q && rinv && workspace, "basis_qr: null pointer");
  if (int e = qr_small_attr()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const QrWs w = qr_ws(workspace, k);
  HFAGP_CUDA(cudaMemsetAsync(workspace, 0, (char*)w.q1 - (char*)workspace, st));
  const TallMat a{bases, 1, m, eps}, q1{w.q1, k, 1, 0.f};
  // deterministic: ONE CTA walks all rows wherever a Gram matrix is accumulated, so the fp64 sums have a fixed order
  // (~1 ms instead of ~40 us; the factor is cached at inference)
  const int cg = deterministic ? cdiv(m, QR_GROWS) : QR_CPC, ca = deterministic ? cdiv(m, QR_AROWS) : QR_CPC;
  const int ng = cdiv(m, QR_GROWS * cg), na = cdiv(m, QR_AROWS * ca);
  qr_gram_kernel<<<ng, 256, 0, st>>>(m, k, cg, a, a, 1, w.g1);
  qr_small_fwd_kernel<<<1, 512, 2 * k * k * 8, st>>>(k, 0, m == k, w.g1, w.ra, w.rinv1, nullptr, nullptr, w.info);
  qr_apply_kernel<<<na, 256, 0, st>>>(m, k, ca, a, w.ra, a, nullptr, w.q1, k, 1, w.g2);
  qr_small_fwd_kernel<<<1, 512, 3 * k * k * 8, st>>>(k, 1, m == k, w.g2, w.rb, w.rinv1, w.q1, rinv, w.info);
  qr_apply_kernel<<<cdiv(m, QR_AROWS * QR_CPC), 256, 0, st>>>(m, k, QR_CPC, q1, w.rb, q1, nullptr, q, k, 1, nullptr);
  HFAGP_CHECK_LAUNCH("basis_qr_fwd");
  return HFAGP_OK;
}

extern "C" int hfagp_basis_qr_info(const void* workspace, int k, int m, int* info_host, void* stream) {
  HFAGP_CHECK_ARG(workspace && info_host, "basis_qr_info: null pointer");
  (void)m;
  const QrWs w = qr_ws(const_cast<void*>(workspace), k);
  HFAGP_CUDA(cudaMemcpyAsync(info_host, w.info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HFAGP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return HFAGP_OK;
}

extern "C" int hfagp_basis_qr_bwd(int k, int m, const float* gq, const float* q, const float* rinv, float* gbases,
                                  void* workspace, int deterministic, void* stream) {
  HFAGP_CHECK_ARG(k >= 1 && k <= QR_MAXK && m >= k, "basis_qr_bwd: bad dims k=%d m=%d", k, m);
  HFAGP_CHECK_ARG(gq && q && rinv && gbases && workspace, "basis_qr_bwd: null pointer");
  if (int e = qr_small_attr()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const QrWs w = qr_ws(workspace, k);
  const TallMat Q{q, k, 1, 0.f}, GQ{gq, k, 1, 0.f};
  HFAGP_CUDA(cudaMemsetAsync(w.g1, 0, (size_t)k * k * 8, st));
  const int cg = deterministic ? cdiv(m, QR_GROWS) : QR_CPC;
  qr_gram_kernel<<<cdiv(m, QR_GROWS * cg), 256, 0, st>>>(m, k, cg, Q, GQ, 0, w.g1);
  qr_small_bwd_kernel<<<1, 512, 2 * k * k * 8, st>>>(k, w.g1, rinv, w.ra, w.rb);
  qr_apply_kernel<<<cdiv(m, QR_AROWS * QR_CPC), 256, 0, st>>>(m, k, QR_CPC, GQ, w.ra, Q, w.rb, gbases, 1, m, nullptr);
  HFAGP_CHECK_LAUNCH("basis_qr_bwd");
  return HFAGP_OK;
}
