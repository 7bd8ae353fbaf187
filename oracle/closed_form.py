"""
ORACLE (test infrastructure only) -- INDEPENDENT closed forms used to pin
oracle/reference_dgp.py, standing in for the GPflow models the reference's
tests compare against (gpflow.models.SVGP / GPR: tests/test_dgp.py:69-78,
tests/test_collapsed.py:46-54).

Deliberately written with different linear algebra from the restatement
(explicit inverses / slogdet / solve in NumPy float64 instead of Cholesky +
triangular solves, kernels from explicit pairwise differences instead of the
-2XX^T expansion) so that agreement is evidence, not tautology.
"""
import math

import numpy as np


def pairwise_r2(X, X2, lengthscales):
    d = (X[:, None, :] - X2[None, :, :]) / lengthscales
    return np.sum(d * d, -1)


def k_rbf(X, X2, variance, lengthscales):
    return variance * np.exp(-0.5 * pairwise_r2(X, X2, lengthscales))


def k_matern52(X, X2, variance, lengthscales):
    r = np.sqrt(pairwise_r2(X, X2, lengthscales) + 1e-12)
    s5 = math.sqrt(5.0)
    return variance * (1 + s5 * r + 5.0 / 3.0 * r * r) * np.exp(-s5 * r)


KERNELS = {'rbf': k_rbf, 'matern52': k_matern52}


def svgp_predict_f(kern, variance, lengthscales, Z, q_mu, q_sqrt, Xs, white, jitter, full_cov=False):
    """q(f*) of a sparse variational GP, zero mean function.
    non-white: q(u)=N(m,S);  white: u = Lu v, q(v)=N(m,S)."""
    k = KERNELS[kern]
    M = Z.shape[0]
    Kuu = k(Z, Z, variance, lengthscales) + jitter * np.eye(M)
    Kuf = k(Z, Xs, variance, lengthscales)
    Kff = k(Xs, Xs, variance, lengthscales)
    Kinv = np.linalg.inv(Kuu)
    D = q_mu.shape[1]
    means, covs = [], []
    for d in range(D):
        S = np.tril(q_sqrt[d]) @ np.tril(q_sqrt[d]).T
        m = q_mu[:, d]
        if white:
            Lu = np.linalg.cholesky(Kuu)
            m_u = Lu @ m
            S_u = Lu @ S @ Lu.T
        else:
            m_u, S_u = m, S
        A = Kinv @ Kuf                                  # M,N
        means.append(A.T @ m_u)
        covs.append(Kff - Kuf.T @ Kinv @ Kuf + A.T @ S_u @ A)
    mean = np.stack(means, 1)
    if full_cov:
        return mean, np.stack(covs, -1)                 # N,N,D
    return mean, np.stack([np.diag(c) for c in covs], 1)


def svgp_kl(kern, variance, lengthscales, Z, q_mu, q_sqrt, white, jitter):
    k = KERNELS[kern]
    M = Z.shape[0]
    Kuu = k(Z, Z, variance, lengthscales) + jitter * np.eye(M)
    P = np.eye(M) if white else Kuu
    Pinv = np.linalg.inv(P)
    _, logdetP = np.linalg.slogdet(P)
    kl = 0.0
    for d in range(q_mu.shape[1]):
        Ld = np.tril(q_sqrt[d])
        S = Ld @ Ld.T
        logdetS = 2.0 * np.sum(np.log(np.abs(np.diag(Ld))))
        m = q_mu[:, d]
        kl += 0.5 * (np.trace(Pinv @ S) + m @ Pinv @ m - M + logdetP - logdetS)
    return kl


def svgp_elbo_gaussian(kern, variance, lengthscales, Z, q_mu, q_sqrt, X, Y, lik_var, white, jitter,
                       num_data=None):
    mean, var = svgp_predict_f(kern, variance, lengthscales, Z, q_mu, q_sqrt, X, white, jitter)
    ve = -0.5 * math.log(2 * math.pi) - 0.5 * math.log(lik_var) - 0.5 * ((Y - mean) ** 2 + var) / lik_var
    scale = (num_data or X.shape[0]) / X.shape[0]
    return scale * np.sum(ve) - svgp_kl(kern, variance, lengthscales, Z, q_mu, q_sqrt, white, jitter)


def gpr_log_marginal(kern, variance, lengthscales, X, Y, lik_var):
    k = KERNELS[kern]
    N = X.shape[0]
    K = k(X, X, variance, lengthscales) + lik_var * np.eye(N)
    _, logdet = np.linalg.slogdet(K)
    Kinv = np.linalg.inv(K)
    ll = 0.0
    for d in range(Y.shape[1]):
        ll += -0.5 * Y[:, d] @ Kinv @ Y[:, d] - 0.5 * logdet - 0.5 * N * math.log(2 * math.pi)
    return ll


def optimal_q_gaussian(kern, variance, lengthscales, Z, Xin, Y, lik_var, jitter, scale=1.0):
    """Optimal (non-white) q(u) of a sparse GP with Gaussian likelihood given inputs Xin (Titsias 2009):
       S = Kuu (Kuu + c Kuf Kfu / s2)^-1 Kuu,  m = c S Kuu^-1 Kuf y / s2   (c = num_data/minibatch).
    What one NatGrad step with gamma=1 lands on (tests/test_collapsed.py:99-104)."""
    k = KERNELS[kern]
    M = Z.shape[0]
    Kuu = k(Z, Z, variance, lengthscales) + jitter * np.eye(M)
    Kuf = k(Z, Xin, variance, lengthscales)
    Sig = np.linalg.inv(Kuu + scale * Kuf @ Kuf.T / lik_var)
    S = Kuu @ Sig @ Kuu
    m = scale * Kuu @ Sig @ Kuf @ Y / lik_var
    return m, S
