"""TEST INFRASTRUCTURE (CPU oracle; never imported by the product path).

The latent basis' QR factorisation of ``get_latent`` — ``Q, _ = torch.qr(bases.T + 1e-8)``
(/root/reference/code/networks/headnerf.py:92, :187, :247) — restated the way ``hfa_gp_b200/csrc/qr.cu`` computes it, so
that the kernel's ALGORITHM (not only its output) is pinned on the CPU against the call the reference makes:

  * CholeskyQR2: two rounds of  G = A^T A,  R = chol(G),  A <- A R^-1  (tall-skinny A [M, K], M >> K);
  * Householder sign reconstruction (Ballard, Demmel, Grigori, Jacquelin, Nguyen, Solomonik: "Reconstructing Householder
    vectors from tall-skinny QR", IPDPS 2014): LAPACK's geqrf leaves diag(R)_j = -sign(x_j) |x_j| with x_j the j-th pivot
    of the Householder sweep; the same sign sequence is produced by an LU without pivoting of the top K x K block of the
    orthonormal factor that picks S_jj = -sgn(pivot_j) and subtracts it from the pivot.  (A square matrix gets no
    reflection for its last column: that sign is the pivot's own.)
  * backward of the reduced factorisation for a gradient arriving at Q only (the m >= n case of torch's
    linalg_qr_backward):  gA = (gQ + Q Y) R^-T,  Y = X + X^T - diag(X),  X = triu(-Q^T gQ).

Pinned by tests/test_oracle_reference.py::test_qr_restatement_equals_torch_qr (values, signs, gradient).
"""
import torch


def cholqr2_signed(a: torch.Tensor):
    """a [M, K] (M >= K, full column rank) -> (q [M, K], r [K, K]) equal to torch.linalg.qr(a, mode='reduced') up to
    rounding, column signs included."""
    m, k = a.shape
    wide = a.double()
    r_total = torch.eye(k, dtype=torch.float64)
    q = wide
    for _ in range(2):
        g = q.T @ q
        r = torch.linalg.cholesky(g, upper=True)
        q = q @ torch.linalg.inv(r)
        r_total = r @ r_total
    t = q[:k].clone()
    s = torch.ones(k, dtype=torch.float64)
    for j in range(k):
        piv = t[j, j]
        s[j] = -1.0 if piv >= 0 else 1.0
        if m == k and j == k - 1:
            s[j] = -s[j]
        t[j, j] = piv - s[j]
        t[j + 1:, j] /= t[j, j]
        t[j + 1:, j + 1:] -= torch.outer(t[j + 1:, j], t[j, j + 1:])
    return (q * s).to(a.dtype), (s[:, None] * r_total).to(a.dtype)


def qr_backward_q_only(gq: torch.Tensor, q: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """Gradient of sum(Q * gq) w.r.t. A for the reduced factorisation A = Q R (R is not used downstream)."""
    x = torch.triu(-(q.T @ gq))
    y = x + x.T
    y.diagonal().mul_(0.5)
    return torch.linalg.solve_triangular(r.T, gq + q @ y, upper=False, left=False)
