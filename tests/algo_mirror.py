"""
NumPy float64 mirror of the ALGORITHM the CUDA kernels implement (csrc/*.cu): the de-duplicated,
whitened-projection formulation with hand-derived adjoints (DESIGN.md "Math").  Test
infrastructure: it lets the derivation be checked against the oracle's autograd on CPU
(tests/test_algo_mirror.py) before/independently of the GPU run, and documents kernel-by-kernel
what each CUDA stage computes.

Per layer (reference layers.py:178-246, utils.py:41):
  prepA : K = k(Z,Z)+jit I ; Lu = chol K ; Linv = Lu^-1
  fwd   : k = k(Z,x) ; b = Linv k ; u = Linv^T b (non-white) | b (white)
          mean_d = u.q_mu[:,d] + mf(x)_d ; c_d = L_d^T u ; var_d = s2 - |b|^2 + |c_d|^2
          f = mean + z sqrt(var + jit)
  bwdA  : (mubar, vbar) -> ubar = sum_d mubar_d m_d + 2 vbar_d L_d c_d ; ...-> w, kbar, xbar, Zbar, lsbar, s2bar
  bwdB  : P_d = sum_r vbar_rd u_r u_r^T ; G = sum_r w_r u_r^T ; qmubar = sum_r u_r mubar_r^T
  fin   : q_sqrt grad, Kuu path (K bar -> Z, ls, s2), KL value and gradient
"""
import math

import numpy as np


def kern_eval(kind, r2, var):
    """returns k and dk/d(r2)."""
    if kind == 'rbf':
        k = var * np.exp(-0.5 * r2)
        return k, -0.5 * k
    r = np.sqrt(r2 + 1e-12)
    s5 = math.sqrt(5.0)
    e = np.exp(-s5 * r)
    k = var * (1 + s5 * r + 5.0 / 3.0 * r * r) * e
    # d/dr = var*e*(-5/3 r - 5 sqrt5/3 r^2) ; dr/dr2 = 1/(2r)
    kp = var * e * (-5.0 / 6.0) * (1 + s5 * r)
    return k, kp


def r2_mat(X, Z, ls):
    d = (X[:, None, :] - Z[None, :, :]) / ls
    return np.sum(d * d, -1)          # R,M


class LayerP:
    def __init__(self, kind, Z, q_mu, q_sqrt, ls, var, white, mean, W=None, bvec=None):
        self.kind, self.Z, self.q_mu, self.q_sqrt = kind, Z, q_mu, np.tril(q_sqrt)
        self.ls = np.broadcast_to(np.asarray(ls, dtype=np.float64), (Z.shape[1],)).copy()
        self.ard = np.ndim(ls) > 0 and np.size(ls) > 1
        self.var, self.white, self.mean, self.W, self.bvec = float(var), white, mean, W, bvec
        self.M, self.Din = Z.shape
        self.Dout = q_mu.shape[1]


def prepA(P, jitter):
    K, _ = kern_eval(P.kind, r2_mat(P.Z, P.Z, P.ls), P.var)
    K = K + jitter * np.eye(P.M)
    Lu = np.linalg.cholesky(K)
    Linv = np.linalg.inv(Lu)
    return K, Lu, Linv


def meanfn(P, X):
    if P.mean == 'zero':
        return 0.0
    if P.mean == 'identity':
        return X
    return X @ P.W + P.bvec


def layer_fwd(P, Linv, X, jitter):
    k, _ = kern_eval(P.kind, r2_mat(X, P.Z, P.ls), P.var)      # R,M
    b = k @ Linv.T
    u = b if P.white else b @ Linv
    mean = u @ P.q_mu + meanfn(P, X)
    c = np.einsum('dij,ri->rdj', P.q_sqrt, u)                   # R,D,M
    var = P.var - np.sum(b * b, 1)[:, None] + np.sum(c * c, 2)
    return mean, var, u


def layer_bwdA(P, Linv, X, u, mubar, vbar):
    """returns xbar, w, and the row-local parameter partials (Zbar, lsbar, s2bar)."""
    r2 = r2_mat(X, P.Z, P.ls)
    k, kp = kern_eval(P.kind, r2, P.var)
    c = np.einsum('dij,ri->rdj', P.q_sqrt, u)
    cbar = 2.0 * vbar[:, :, None] * c
    ubar = mubar @ P.q_mu.T + np.einsum('dij,rdj->ri', P.q_sqrt, cbar)
    vs = np.sum(vbar, 1)[:, None]
    if P.white:
        bbar = ubar - 2.0 * vs * u
        kbar = bbar @ Linv            # kbar_i = sum_j Linv[j,i] bbar_j
        w = kbar
    else:
        # var = s2 - k.u + ..., u = K^-1 k
        ubar = ubar - vs * k
        w = (ubar @ Linv.T) @ Linv    # K^-1 ubar
        kbar = w - vs * u
    s2bar = np.sum(vbar) + np.sum(kbar * k) / P.var
    g = 2.0 * kbar * kp                                   # R,M
    diff = X[:, None, :] - P.Z[None, :, :]                # R,M,Din
    xbar = np.einsum('rm,rmq->rq', g, diff) / P.ls ** 2
    Zbar = -np.einsum('rm,rmq->mq', g, diff) / P.ls ** 2
    lsbar = -np.einsum('rm,rmq->q', g, diff ** 2) / P.ls ** 3
    if P.mean == 'identity':
        xbar = xbar + mubar
    elif P.mean == 'linear':
        xbar = xbar + mubar @ P.W.T
    return xbar, w, Zbar, lsbar, s2bar


def layer_bwdB(u, w, mubar, vbar):
    Pd = np.einsum('rd,ri,rj->dij', vbar, u, u)
    G = w.T @ u                                            # G[i,j] = sum_r w_ri u_rj
    qmubar = u.T @ mubar
    return Pd, G, qmubar


def chol_bwd(Lu, Linv, Lbar):
    """Kbar (symmetric) from Lbar (lower)  -- Murray (2016)."""
    Phi = np.tril(Lu.T @ np.tril(Lbar))
    Phi[np.diag_indices_from(Phi)] *= 0.5
    Kb = Linv.T @ Phi @ Linv
    return 0.5 * (Kb + Kb.T)


def kuu_bwd(P, Kbar):
    """Kbar symmetric (dELBO/dK treating K's entries as independent) -> Z, ls, s2 grads."""
    r2 = r2_mat(P.Z, P.Z, P.ls)
    k, kp = kern_eval(P.kind, r2, P.var)
    g = Kbar * kp                                          # dE/dr2_ij
    diff = P.Z[:, None, :] - P.Z[None, :, :]
    Zbar = 2.0 * (np.einsum('ij,ijq->iq', g, diff) - np.einsum('ij,ijq->jq', g, diff)) / P.ls ** 2
    lsbar = -2.0 * np.einsum('ij,ijq->q', g, diff ** 2) / P.ls ** 3
    s2bar = np.sum(Kbar * k) / P.var
    return Zbar, lsbar, s2bar


def layer_fin(P, K, Lu, Linv, Pd, G, qmubar, klw=1.0):
    """Combine row-reduced accumulators with the KL term (weight klw = 1/world).
    Returns KL value and ELBO-gradients (q_mu, q_sqrt, Zbar, lsbar, s2bar) of the Kuu/KL part."""
    M, D = P.M, P.Dout
    Lq = P.q_sqrt
    dinv = np.zeros_like(Lq)
    for d in range(D):
        dinv[d][np.diag_indices(M)] = 1.0 / np.diag(Lq[d])
    logq = np.sum(np.log(np.stack([np.diag(Lq[d]) for d in range(D)]) ** 2))
    if P.white:
        KL = -0.5 * D * M - 0.5 * logq + 0.5 * np.sum(Lq ** 2) + 0.5 * np.sum(P.q_mu ** 2)
        gq_sqrt = np.stack([np.tril(2.0 * Pd[d] @ Lq[d] - klw * Lq[d]) for d in range(D)]) + klw * dinv
        gq_mu = qmubar - klw * P.q_mu
        Lbar = -np.tril(G)              # dE/dLu from b = Linv k
        Kbar = chol_bwd(Lu, Linv, Lbar)
    else:
        Kinv = Linv.T @ Linv
        Ssum = np.einsum('dij,dkj->ik', Lq, Lq) + P.q_mu @ P.q_mu.T
        KL = (-0.5 * D * M - 0.5 * logq + D * np.sum(np.log(np.diag(Lu)))
              + 0.5 * np.sum(Kinv * Ssum))
        gq_sqrt = np.stack([np.tril((2.0 * Pd[d] - klw * Kinv) @ Lq[d]) for d in range(D)]) + klw * dinv
        gq_mu = qmubar - klw * Kinv @ P.q_mu
        Kbar = -0.5 * (G + G.T)
        Kbar = Kbar - klw * (0.5 * D * Kinv - 0.5 * Kinv @ Ssum @ Kinv)
    Zbar, lsbar, s2bar = kuu_bwd(P, Kbar)
    return KL, gq_mu, gq_sqrt, Zbar, lsbar, s2bar


def gaussian_lik(mean, var, Y, lik_var, c):
    """c = num_data/(N*S).  Returns L contribution, mubar, vbar, lik_var grad."""
    ve = -0.5 * math.log(2 * math.pi) - 0.5 * math.log(lik_var) - 0.5 * ((Y - mean) ** 2 + var) / lik_var
    mubar = c * (Y - mean) / lik_var
    vbar = np.full_like(var, -0.5 * c / lik_var)
    lvbar = c * np.sum(-0.5 / lik_var + 0.5 * ((Y - mean) ** 2 + var) / lik_var ** 2)
    return c * np.sum(ve), mubar, vbar, lvbar


def elbo_and_grad(layers, X, Y, lik_var, S, zs, num_data, jitter, n_global=None, klw=1.0, aux=None):
    """Full step in the kernels' order.  X (N,D), zs[l] (S,N,Dout_l).  Layer 1 is evaluated on the N
    distinct rows only (the reference's tile, dgp.py:63, makes its conditional S-fold redundant).
    n_global / klw: the row-sharded data-parallel protocol of csrc/api.cu -- this rank holds N of n_global minibatch
    rows, the likelihood term is scaled by num_data / (n_global * S) and the KL term (value and gradient) is
    weighted by klw = 1/world, so that a plain SUM all-reduce over ranks gives the full ELBO and gradient."""
    N = X.shape[0]
    n_global = n_global or N
    L = len(layers)
    preps = [prepA(P, jitter) for P in layers]
    # ---------------- forward
    acts = []
    Xin = X
    for l, P in enumerate(layers):
        K, Lu, Linv = preps[l]
        mean, var, u = layer_fwd(P, Linv, Xin, jitter)
        sd = np.sqrt(var + jitter)
        acts.append((Xin, mean, var, u, sd))
        if l < L - 1:
            if l == 0:
                F = mean[None] + zs[0] * sd[None]                  # S,N,D
                Xin = F.reshape(S * N, -1)
            else:
                Xin = mean + zs[l].reshape(S * N, -1) * sd
    Xl, mean, var, u, sd = acts[-1]
    if L == 1:
        c = num_data / n_global
        Yr = Y
    else:
        c = num_data / (n_global * S)
        Yr = np.tile(Y, (S, 1))
    Lval, mubar, vbar, lvbar = gaussian_lik(mean, var, Yr, lik_var, c)
    # ---------------- backward
    grads = [None] * L
    KLs = 0.0
    for l in reversed(range(L)):
        P = layers[l]
        K, Lu, Linv = preps[l]
        Xl, mean, var, u, sd = acts[l]
        xbar, w, Zb, lsb, s2b = layer_bwdA(P, Linv, Xl, u, mubar, vbar)
        Pd, G, qmub = layer_bwdB(u, w, mubar, vbar)
        if aux is not None:                   # row-reduced accumulators, as csrc/natgrad.cu reads them
            aux[l] = dict(Pd=Pd, qmubar=qmub, Kinv=Linv.T @ Linv)
        KL, gq_mu, gq_sqrt, Zb2, lsb2, s2b2 = layer_fin(P, K, Lu, Linv, Pd, G, qmub, klw=klw)
        KLs += KL
        lsg = lsb + lsb2
        grads[l] = dict(Z=Zb + Zb2, q_mu=gq_mu, q_sqrt=gq_sqrt,
                        ls=lsg if P.ard else np.sum(lsg), var=s2b + s2b2)
        if l > 0:
            fbar = xbar                                            # (S*N, Dout_{l-1})
            Pm = layers[l - 1]
            _, pmean, pvar, pu, psd = acts[l - 1]
            if l - 1 == 0:
                fb = fbar.reshape(S, N, -1)
                mubar = fb.sum(0)
                vbar = (fb * zs[0]).sum(0) / (2.0 * psd)
            else:
                mubar = fbar
                vbar = fbar * zs[l - 1].reshape(S * N, -1) / (2.0 * psd)
    return Lval - klw * KLs, grads, lvbar


def natgrad_update(P, Kinv, Pd, qmubar, gamma):
    """csrc/natgrad.cu: natural-gradient step on (q_mu, q_sqrt) of one layer straight from the row-reduced
    accumulators of the backward pass (summed over ranks), SURVEY App. B last paragraph:
      dELBO/dS_d = P_d - 1/2 Prior^-1 + 1/2 S_d^-1 ,  dELBO/dm_d = qmubar_d - Prior^-1 m_d     (Prior = K, or I if white)
      => -2 theta2' = (1-gamma) S^-1 + gamma (Prior^-1 - 2 P_d)
            theta1' = (1-gamma) S^-1 m + gamma (qmubar_d - 2 P_d m)
      then GPflow's natural_to_meanvarsqrt: C = chol(-2 theta2'), V = C^-1, S' = V^T V, m' = S' theta1', q_sqrt' = chol S'.
    With gamma == 1 the old S drops out exactly (no S^-1 is formed)."""
    M, D = P.M, P.Dout
    prior_inv = np.eye(M) if P.white else Kinv
    new_mu = np.empty_like(P.q_mu)
    new_sqrt = np.empty_like(P.q_sqrt)
    for d in range(D):
        m = P.q_mu[:, d]
        P2 = Pd[d] + Pd[d].T                   # 2 P_d, symmetrised as the kernel does (atomics)
        if gamma == 1.0:
            prec = prior_inv - P2
            t1 = qmubar[:, d] - P2 @ m
        else:
            S = P.q_sqrt[d] @ P.q_sqrt[d].T
            Ls = np.linalg.cholesky(S)
            Li = np.linalg.inv(Ls)
            Sinv = Li.T @ Li
            prec = (1.0 - gamma) * Sinv + gamma * (prior_inv - P2)
            t1 = (1.0 - gamma) * (Sinv @ m) + gamma * (qmubar[:, d] - P2 @ m)
        C = np.linalg.cholesky(prec)
        V = np.linalg.inv(C)
        S_new = V.T @ V
        new_mu[:, d] = S_new @ t1
        new_sqrt[d] = np.linalg.cholesky(S_new)
    return new_mu, new_sqrt


def full_cov_propagate(layers, X, S, zs, jitter):
    """csrc/full_cov.cu, kernel by kernel (float64): per layer and sample, Kuf/Kff (k_fc_gram), A = Kuu^-1 Kuf or Lu^-1 Kuf
    (k_fc_A), SK_d (k_fc_SK), B_d = SK_d A (k_fc_B), var_d = Kff + A^T B_d (k_fc_cov), mean (k_fc_mean),
    f = mean + chol(var_d + jitter I) z (k_fc_chol_draw).  Returns lists Fs, Fmeans (S,N,D) and Fvars (S,N,N,D)."""
    N = X.shape[0]
    Fs, Fms, Fvs = [], [], []
    Xin = np.broadcast_to(X[None], (S,) + X.shape)
    for l, P in enumerate(layers):
        K, Lu, Linv = prepA(P, jitter)
        W = Linv if P.white else Linv.T @ Linv
        M, D = P.M, P.Dout
        F = np.empty((S, N, D)); Fm = np.empty((S, N, D)); Fv = np.empty((S, N, N, D))
        for s in range(S):
            Xs = Xin[s]
            Kuf, _ = kern_eval(P.kind, r2_mat(P.Z, Xs, P.ls), P.var)        # (M, N)
            Kff, _ = kern_eval(P.kind, r2_mat(Xs, Xs, P.ls), P.var)
            A = W @ Kuf
            Fm[s] = A.T @ P.q_mu + meanfn(P, Xs)
            for d in range(D):
                SK = P.q_sqrt[d] @ P.q_sqrt[d].T - (np.eye(M) if P.white else K)
                cov = Kff + A.T @ (SK @ A)
                Fv[s, :, :, d] = cov
                C = np.linalg.cholesky(cov + jitter * np.eye(N))
                F[s, :, d] = Fm[s, :, d] + C @ zs[l][s, :, d]
        Fs.append(F); Fms.append(Fm); Fvs.append(Fv)
        Xin = F
    return Fs, Fms, Fvs
