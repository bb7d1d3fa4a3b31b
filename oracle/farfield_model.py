"""TEST INFRASTRUCTURE ONLY — numpy model of the far-field tile records the d <= 2 kernels use
(skysampler_b200/csrc/skb_fit.cu: skb_moment_kernel / skb_moment2_kernel; skb_score.cu: skb_score_d1_kernel, ff2_eval).

With prescaled coordinates a pair term is 2^-|q - x|^2.  About a centre c (s = q - c, t = x - c)

    sum_j w_j 2^-|s - t_j|^2  =  2^-|s|^2 * sum_j w_j 2^-|t_j|^2 e^(2 ln2 s.t_j)
                               ~  2^-|s|^2 * sum_{|a| <= P} B_a s^a,     B_a = (2 ln2)^|a| / a! * sum_j w_j t_j^a 2^-|t_j|^2

(the Taylor series of e^u cut after u^P).  With |s.t_j| <= sum_k |s_k| r_k the remainder is at most
z^(P+1)/(P+1)! e^z of sum_j w_j 2^-|t_j|^2 and the kept sum is at least e^-z of it, z = 2 ln2 sum_k |s_k| r_k: the relative
error of the record is bounded by  z^(P+1)/(P+1)! e^(2z)  — the number the kernels' admission tests are built on.
Nothing here is imported by the product; tests/test_oracle.py checks the bound and the fp32 evaluation error on the CPU.
"""
import math

import numpy as np

LN2 = math.log(2.0)


def remainder_bound(z, P):
    """Relative error bound of a degree-P record at admission value z."""
    return z ** (P + 1) / math.factorial(P + 1) * math.exp(2.0 * z)


def coefficients(t, w, P):
    """B_a for points t [n, d] (d = 1 or 2) relative to the centre, weights w; dict {multi-index tuple: float64}."""
    t = np.atleast_2d(np.asarray(t, dtype=np.float64))
    if t.shape[0] == 1 and t.shape[1] != 1 and np.ndim(w) and len(w) == t.shape[1]:
        t = t.T
    n, d = t.shape
    g = np.asarray(w, dtype=np.float64) * np.exp2(-(t * t).sum(1))
    out = {}
    if d == 1:
        for i in range(P + 1):
            out[(i,)] = float((g * t[:, 0] ** i).sum() * (2 * LN2) ** i / math.factorial(i))
    else:
        for i in range(P + 1):
            for j in range(P + 1 - i):
                out[(i, j)] = float((g * t[:, 0] ** i * t[:, 1] ** j).sum() * (2 * LN2) ** (i + j)
                                    / (math.factorial(i) * math.factorial(j)))
    return out


def evaluate(B, s, dtype=np.float64):
    """The record's value sum_j w_j 2^-|s - t_j|^2 at offset s, Horner form like the kernels, in `dtype` arithmetic."""
    s = np.atleast_1d(np.asarray(s, dtype=dtype))
    P = max(sum(a) for a in B)
    if len(s) == 1:
        p = dtype(0)
        for i in range(P, -1, -1):
            p = dtype(p * s[0] + dtype(B[(i,)]))
    else:
        p = dtype(0)
        for i in range(P, -1, -1):
            inner = dtype(0)
            for j in range(P - i, -1, -1):
                inner = dtype(inner * s[1] + dtype(B[(i, j)]))
            p = dtype(p * s[0] + inner)
    return dtype(np.exp2(-dtype((s * s).sum()))) * p


def exact(t, w, s):
    t = np.atleast_2d(np.asarray(t, dtype=np.float64))
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    return float((np.asarray(w, dtype=np.float64) * np.exp2(-((s[None, :] - t) ** 2).sum(1))).sum())
