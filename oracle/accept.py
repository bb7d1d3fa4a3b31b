"""Restatement of the accept test of ``read_concentric`` (emulator.py:737-746).  TEST INFRASTRUCTURE only."""
import numpy as np


def accept(dc, dc_jac, wr, wr_jac, wcr, wcr_jac, m_factor=20, seed=6, uniform=None):
    dc_score = np.exp(dc) * np.abs(dc_jac)            # :737
    wr_score = np.exp(wr) * np.abs(wr_jac)            # :738
    wcr_score = np.exp(wcr) * np.abs(wcr_jac)         # :739
    if uniform is None:
        uniform = np.random.RandomState(seed).uniform(0, 1, len(dc))   # :741-742
    p_proposal = m_factor * dc_score * wr_score       # :743
    p_ref = wcr_score                                 # :744
    return uniform < p_ref / p_proposal, uniform      # :746


def log_ratio(logps, log_abs_jacs, signs, log_m):
    """The same test in log space, as the fused GPU epilogue forms it:
    accept <=> log u < sum_k sign_k (log p_k + log|J_k|) - log m."""
    r = -float(log_m) * np.ones_like(np.asarray(logps[0], dtype=np.float64))
    for lp, lj, s in zip(logps, log_abs_jacs, signs):
        r = r + s * (np.asarray(lp, dtype=np.float64) + lj)
    return r
