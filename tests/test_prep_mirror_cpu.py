"""Index-level NumPy emulation of the fp64 factorisation kernels of csrc/small_matrix.cu (k_prepA_ldl: square-root-free,
one barrier per pivot, transposed factor; k_prepA_c4: four pivots per panel on the combined factor/inverse array), checked
against numpy.linalg.  The emulation walks the same (panel, pivot, row, column) index sets and activity predicates as the
kernels' loops, so an off-by-one in a band / mask shows up here, on CPU, for every M including M % 4 != 0 and M < 4."""
import numpy as np
import pytest
from numpy.testing import assert_allclose


def spd(M, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(M, M))
    return A @ A.T + M * np.eye(M)


def emulate_ldl(K):
    """k_prepA_ldl: B[j][i] (i >= j) = A_j[i][j]; X = unit-lower inverse; outputs scaled by d^-1/2 at the end."""
    M = K.shape[0]
    B, X = np.triu(K).copy(), np.eye(M)
    for j in range(M):
        r = 1.0 / B[j, j]
        pr, xr = B[j].copy(), X[j].copy()
        for k in range(j + 1, M):                      # rows k, one warp each
            t = B[j, k] * r
            for i in range(max(k, j + 1), M):          # lanes: i = j+1+lane+32b, active iff i >= k
                B[k, i] -= t * pr[i]
            for c in range(j + 1):                     # lanes: c = lane+32b <= j
                X[k, c] -= t * xr[c]
    rsq = 1.0 / np.sqrt(np.diag(B))
    Lu, Linv = np.zeros((M, M)), np.zeros((M, M))
    for i in range(M):
        for j in range(i + 1):
            Lu[i, j] = 1.0 / rsq[j] if i == j else B[j, i] * rsq[j]
            Linv[i, j] = X[i, j] * rsq[i]
    return Lu, Linv


def emulate_c4(K):
    """k_prepA_c4: C[k][c >= k] = transposed factor, C[k][c < k] = unit-lower inverse (implicit ones on its diagonal)."""
    M = K.shape[0]
    C = np.triu(K).copy()
    for j0 in range(0, M, 4):
        w = min(4, M - j0)
        for p in range(3):                             # panel phase: warps 0..2 own rows j0+1..j0+3
            j = j0 + p
            for r in range(3):
                rrow = j0 + 1 + r
                if p < w - 1 and rrow < j0 + w and rrow > j:
                    t = C[j, rrow] / C[j, j]
                    for c in range(M):
                        pj = 1.0 if c == j else C[j, c]
                        if c <= j or c >= rrow:
                            C[rrow, c] -= t * pj
        kb = j0 + w
        if kb >= M:
            continue
        rq = [1.0 / C[j0 + q, j0 + q] for q in range(w)]
        pr = np.zeros((w, M))
        for q in range(w):                             # pivot rows as the lanes hold them
            jq = j0 + q
            for c in range(M):
                pr[q, c] = 1.0 if c == jq else (C[jq, c] if (c < jq or c >= kb) else 0.0)
        for k in range(kb, M):                         # trailing rows: rank-w update, active iff c < kb or c >= k
            t = [C[j0 + q, k] * rq[q] for q in range(w)]
            for c in range(M):
                if c < kb or c >= k:
                    C[k, c] -= sum(t[q] * pr[q, c] for q in range(w))
    rsq = 1.0 / np.sqrt(np.diag(C))
    Lu, Linv = np.zeros((M, M)), np.zeros((M, M))
    for i in range(M):
        for j in range(i + 1):
            Lu[i, j] = 1.0 / rsq[i] if i == j else C[j, i] * rsq[j]
            Linv[i, j] = rsq[i] if i == j else C[i, j] * rsq[i]
    return Lu, Linv


@pytest.mark.parametrize("emulate", [emulate_ldl, emulate_c4], ids=["ldl", "c4"])
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 6, 7, 9, 17, 37, 64, 100])
def test_factorisation_index_logic(M, emulate):
    K = spd(M, 10 + M)
    Lu, Linv = emulate(K)
    L = np.linalg.cholesky(K)
    assert_allclose(Lu, L, rtol=1e-11, atol=1e-12)
    assert_allclose(Linv, np.linalg.inv(L), rtol=1e-10, atol=1e-13)
    assert np.all(np.triu(Lu, 1) == 0) and np.all(np.triu(Linv, 1) == 0)


def test_non_positive_pivot_is_visible_on_the_diagonal():
    """The kernels report failure from the final pass over the stored pivots: a duplicated inducing point with zero
    jitter leaves an exactly-zero pivot on the diagonal (tests/test_gpu_parity.py::test_not_positive_definite_is_reported)."""
    K = spd(4, 3)
    K[1, :] = K[0, :]; K[:, 1] = K[:, 0]; K[1, 1] = K[0, 0]
    C = np.triu(K).copy()
    t = C[0, 1] / C[0, 0]
    C[1, 1:] -= t * C[0, 1:]
    assert not (C[1, 1] > 0.0)
