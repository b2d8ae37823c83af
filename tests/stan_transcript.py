"""Independent transcription of the three .stan programs into torch (fp64, autograd).

TEST INFRASTRUCTURE.  Written directly from /root/reference/stan/*.stan, statement by
statement, WITHOUT using the hand-derived gradient formulas of the oracle: gradients come
from torch's reverse-mode autodiff, like Stan Math's.  It pins the oracle's analytic
gradient and its constant-dropping conventions from a second, structurally different
implementation (SURVEY.md §8c "how the oracle earns trust").
"""
import numpy as np
import torch


def _layout(d, family):
    N = int(np.sum(d["N_spots"]))
    C = int(d["N_covariates"])
    K = int(d["N_celltypes"]) if family == "comp_splotch_stan_model" else 1
    L1, L2, L3 = int(d["N_level_1"]), int(d.get("N_level_2", 0)), int(d.get("N_level_3", 0))
    return N, C, K, L1, L2, L3


def stan_log_prob(d, family, theta, jacobian=True):
    """lp (propto=true) as a differentiable torch scalar of the unconstrained vector theta."""
    N, C, K, L1, L2, L3 = _layout(d, family)
    L = L1 + L2 + L3
    car, zi = int(d["car"]), int(d["zi"])
    comp = family == "comp_splotch_stan_model"
    pos = 0
    lp = torch.zeros((), dtype=torch.float64)

    def take(n):
        nonlocal pos
        v = theta[pos:pos + n]
        pos += n
        return v

    psi = take(N) if car else None
    if comp:
        beta_raw = take(L * C * K).reshape(L, C, K)           # arrays outer-to-inner, vector innermost
    else:
        beta_raw = take(L * C).reshape(C, L).T.unsqueeze(-1)  # matrix column-major
    if car:
        u = take(1)[0]; tau = torch.exp(u)
        if jacobian: lp = lp + u
        u = take(1)[0]; a = torch.sigmoid(u)
        if jacobian: lp = lp + torch.log(a) + torch.log1p(-a)
    u = take(1)[0]; sigma = torch.exp(u)
    if jacobian: lp = lp + u
    if L2:
        u = take(1)[0]; s2 = torch.exp(u)
        if jacobian: lp = lp + u
    if L3:
        u = take(1)[0]; s3 = torch.exp(u)
        if jacobian: lp = lp + u
    if zi:
        u = take(1)[0]; theta_zi = torch.sigmoid(u)
        if jacobian: lp = lp + torch.log(theta_zi) + torch.log1p(-theta_zi)
    nb = int(d.get("nb", 0)) if comp else 0
    if nb:
        u = take(1)[0]; phi = 1e-6 + torch.exp(u)
        if jacobian: lp = lp + u
    noise_raw = take(N)
    assert pos == theta.numel()

    # transformed parameters
    if comp:
        std = torch.tensor(np.asarray(d["beta_prior_std"], dtype=np.float64))
        mean = torch.tensor(np.asarray(d["beta_prior_mean"], dtype=np.float64))
        b1 = beta_raw[:L1] * std + mean
    else:
        b1 = 2.0 * beta_raw[:L1]
    levels = [b1]
    if L2:
        m2 = torch.tensor(np.asarray(d["level_2_mapping"]) - 1, dtype=torch.long)
        levels.append(b1[m2] + s2 * beta_raw[L1:L1 + L2])
    if L3:
        m3 = torch.tensor(np.asarray(d["level_3_mapping"]) - 1, dtype=torch.long)
        levels.append(levels[1][m3] + s3 * beta_raw[L1 + L2:])
    bottom = levels[-1]
    tm = np.asarray(d["tissue_mapping"]) - 1
    row = torch.tensor(np.repeat(tm, np.asarray(d["N_spots"])), dtype=torch.long)
    Dm1 = torch.tensor(np.asarray(d["D"]) - 1, dtype=torch.long)
    bj = bottom[row, Dm1]                                        # (N, K)
    if comp:
        E = torch.tensor(np.asarray(d["E"], dtype=np.float64).reshape(N, K))
        log_lambda = (bj * E).sum(1)
    else:
        log_lambda = bj[:, 0]
    if car:
        log_lambda = log_lambda + psi
    log_lambda = log_lambda + 0.3 * sigma * noise_raw

    # model
    if car:
        lp = lp + (-2.0 * torch.log(tau) - 1.0 / tau)               # inv_gamma(1,1), constants dropped
        Ds = torch.tensor(np.asarray(d["D_sparse"], dtype=np.float64))
        lam = torch.tensor(np.asarray(d["eig_values"], dtype=np.float64))
        if family == "splotch_stan_model_csr":
            U = np.asarray(d["U"]) - 1; V = np.asarray(d["V"]) - 1
            rows = np.repeat(np.arange(N), np.diff(U))
            Wpsi = torch.zeros(N, dtype=torch.float64).index_add(0, torch.tensor(rows), psi[torch.tensor(V)])
        else:
            W = np.asarray(d["W_sparse"]).reshape(-1, 2) - 1
            i1 = torch.tensor(W[:, 0]); i2 = torch.tensor(W[:, 1])
            Wpsi = torch.zeros(N, dtype=torch.float64).index_add(0, i1, psi[i2]).index_add(0, i2, psi[i1])
        ldet = torch.log1p(-a * lam).sum()
        lp = lp + 0.5 * (N * torch.log(tau) + ldet - tau * ((psi * Ds * psi).sum() - a * (Wpsi * psi).sum()))
    if zi:
        lp = lp + torch.log1p(-theta_zi)                             # beta(1,2)
    lp = lp - 0.5 * sigma ** 2 - 0.5 * (noise_raw ** 2).sum()
    if L2: lp = lp - 0.5 * s2 ** 2
    if L3: lp = lp - 0.5 * s3 ** 2
    lp = lp - 0.5 * (beta_raw ** 2).sum()

    y = torch.tensor(np.asarray(d["counts"], dtype=np.float64))
    eta = log_lambda + torch.log(torch.tensor(np.asarray(d["size_factors"], dtype=np.float64)))
    if nb:
        lp = lp + torch.log(phi) - phi                               # gamma(2,1), constants dropped
        # neg_binomial_2_log_lpmf as Stan writes it (binomial_coefficient_log kept whole)
        logphi = torch.log(phi)
        L = torch.nn.functional.softplus(eta - logphi)
        nbv = torch.lgamma(y + phi) - torch.lgamma(y + 1.0) - torch.lgamma(phi) + y * eta - phi * L - y * (logphi + L)
        if zi:
            mult = float(N) if nb == 1 else 1.0                     # phi_arr broadcast of the reference (comp_...stan:257,260)
            heads = torch.log(theta_zi); tails = torch.log1p(-theta_zi)
            zero = y == 0
            lp = lp + torch.logsumexp(torch.stack([heads.expand(int(zero.sum())), tails + mult * nbv[zero]]), 0).sum()
            lp = lp + (tails + mult * nbv[~zero]).sum()
        else:
            lp = lp + nbv.sum()
    elif zi:
        heads = torch.log(theta_zi); tails = torch.log1p(-theta_zi)
        zero = y == 0
        lp = lp + torch.logsumexp(torch.stack([heads.expand(int(zero.sum())), tails - torch.exp(eta[zero])]), 0).sum()
        nz = ~zero
        lp = lp + (tails + y[nz] * eta[nz] - torch.exp(eta[nz]) - torch.lgamma(y[nz] + 1.0)).sum()
    else:
        lp = lp + (y * eta - torch.exp(eta)).sum()
    return lp


def stan_log_prob_grad(d, family, theta, jacobian=True):
    th = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    lp = stan_log_prob(d, family, th, jacobian)
    (g,) = torch.autograd.grad(lp, th)
    return float(lp.detach()), g.numpy()
