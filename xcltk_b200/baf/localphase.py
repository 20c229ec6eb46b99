"""Local phasing of the SNPs inside one long region (host pre-step of `xcltk baf`, numpy).

What xcltk/baf/localphase.py computes (snp_local_phasing :14-132, Local_Phasing :138-243, the Gaussian
smoothing :247-275,:320-343), with its defaults: a two-component EM over SNPs ("flipped" / "not flipped", one allelic
ratio per cell), the SNP posteriors smoothed along the genome with a Gaussian kernel (b = 20 kb), cells whose
aggregated BAF lies in [0.45, 0.55] dropped, the SNPs flipped and the whole thing repeated until a round flips
nothing or everything.  The float expressions are kept operation for operation: which SNPs end up flipped
decides which haplotype a UMI counts for, and the goldens pin that.
"""

import numpy as np
from scipy.special import logsumexp


def smoothing_kernel(x, b=20000.0, a=0.0):
    """Row-normalised Gaussian kernel over the SNP positions: w_ij = exp(a - (x_j - x_i)^2 / b^2) / sum_j w_ij.
    The positions do not change during the EM, so the kernel is built once per region (the reference rebuilds
    row i inside every smoothing call, localphase.py:270-274; the values are the same, operation for operation)."""
    W = np.exp(a - (x[None, :] - x[:, None]) ** 2 / b ** 2)
    return W / np.sum(W, axis=1, keepdims=True)


def smooth_along_genome(v, W):
    """u[i] = sum_j v[j] W[i, j]  (gaussian_smoothing_1d, localphase.py:247-275)."""
    return np.sum(v[None, :] * W, axis=1)


def _loglik(AD, BD, thetas, eps=1e-6):
    th = thetas.copy()
    th[th <= 0] = eps
    th[th >= 1] = 1 - eps
    return AD @ np.log(th) + BD @ np.log(1 - th)


def phasing_em(AD, DP, W, min_iter=10, max_iter=1000, epsilon_conv=1e-3):
    """AD, DP: SNP x cell; W: smoothing kernel over the SNPs (or None).  Returns (Z, thetas, logLik):
    Z[:, 0] / Z[:, 1] = posterior of "as is" / "flipped" (Local_Phasing with init_mode 'warm', smooth_gk,
    localphase.py:138-243)."""
    N, M = AD.shape
    BD = DP - AD
    Z = np.zeros((N, 2))
    Z[:, 0] = (AD.sum(1) / DP.sum(1)).reshape(-1)
    Z[:, 1] = 1 - Z[:, 0]
    thetas = np.array((AD.T @ Z + BD.T @ (1 - Z)) / (DP.T.sum(1, keepdims=True)))
    ll_mat = _loglik(AD, BD, thetas)
    ll_new = np.sum(logsumexp(ll_mat, axis=1))
    for it in range(max_iter):
        ll_old = ll_new + 0.0
        E = np.exp(np.array(ll_mat) - np.max(ll_mat, axis=-1, keepdims=True))          # E step
        Z = E / np.sum(E, axis=-1, keepdims=True)
        if W is not None:
            Z[:, 0] = smooth_along_genome(Z[:, 0], W)
            Z[:, 1] = 1 - Z[:, 0]
        thetas = np.array((AD.T @ Z + BD.T @ (1 - Z)) / (DP.T.sum(1, keepdims=True)))    # M step
        ll_mat = _loglik(AD, BD, thetas)
        ll_new = np.sum(logsumexp(ll_mat, axis=1))
        if it >= min_iter and ll_new - ll_old < epsilon_conv:
            break
    return Z, thetas, ll_new


def snp_local_phasing(AD, DP, positions, min_iter=5, max_iter=50, cw_min_expr_snps=1, cw_low_baf=0.45,
                      cw_up_baf=0.55):
    """AD, DP: cell x SNP (AD already oriented by the reference phasing).  Returns the boolean flip vector
    over the SNPs, or None when no cell is left to phase with (snp_local_phasing, localphase.py:14-132)."""
    idx = DP.sum(axis=1) > 0
    AD, DP = AD[idx, :], DP[idx, :]
    if AD.shape[0] <= 0:
        return None
    idx = (DP > 0).sum(axis=1) >= cw_min_expr_snps
    AD, DP = AD[idx, :], DP[idx, :]
    if AD.shape[0] <= 0:
        return None
    flip_final = None
    W = smoothing_kernel(np.asarray(positions))
    for i in range(max_iter):
        baf = AD.sum(axis=1) / DP.sum(axis=1)
        idx = np.logical_or(baf < cw_low_baf, baf > cw_up_baf)
        AD, DP = AD[idx, :], DP[idx, :]
        if AD.shape[0] <= 0:
            return None
        Z, _thetas, _ll = phasing_em(AD.T, DP.T, W)
        flip = np.array(Z[:, 1] >= Z[:, 0])
        flip_final = flip if i == 0 else np.logical_xor(flip_final, flip)
        if i > 0 and i >= min_iter and (np.all(flip) or np.all(np.logical_not(flip))):
            return flip_final
        AD = AD * (1 - flip.T) + (DP - AD) * flip.T
    return flip_final
