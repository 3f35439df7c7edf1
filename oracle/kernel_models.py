"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  numpy models of the device algorithms.

Each function mirrors, step for step but serially, the *parallel formulation* one
CUDA kernel uses (blocked contraction, low-rate affine scan, incremental window
quantile, fix-point distance suppression ...).  They exist so the mathematics of
those formulations can be checked against ``oracle/ref_port.py`` on a CPU, where
no GPU is available; the CUDA kernels are then transcriptions of these.  Nothing
in the product imports this file.
"""
from __future__ import annotations

import numpy as np

from bpm_analysis_b200.design import PADLEN, BlockFilterDesign


def ext_sample(x: np.ndarray, n_dec: int, stride: int, e):
    """Odd-extended, strided view: ext[e] for e in [0, n_dec + 2*PADLEN)."""
    i = np.asarray(e, dtype=np.int64) - PADLEN
    lo = i < 0
    hi = i >= n_dec
    mid = np.clip(i, 0, n_dec - 1)
    v = x[mid * stride].astype(np.float64)
    if np.any(lo):
        v = np.where(lo, 2.0 * float(x[0]) - x[np.clip(-i, 0, n_dec - 1) * stride].astype(np.float64), v)
    if np.any(hi):
        v = np.where(hi, 2.0 * float(x[(n_dec - 1) * stride])
                     - x[np.clip(2 * (n_dec - 1) - i, 0, n_dec - 1) * stride].astype(np.float64), v)
    return v


def _df2t_step(sos, s, x):
    """One sample through the cascade, DF2T; s = (s1a, s2a, s1b, s2b)."""
    s = s.copy()
    for k in range(sos.shape[0]):
        b0, b1, b2, _, a1, a2 = sos[k]
        y = b0 * x + s[2 * k]
        s[2 * k] = b1 * x - a1 * y + s[2 * k + 1]
        s[2 * k + 1] = b2 * x - a2 * y
        x = y
    return s, x


def blocked_filtfilt(x: np.ndarray, d: BlockFilterDesign, stride: int = 1) -> np.ndarray:
    """Zero-phase band-pass, kept samples only: model of the K0/K1 device path.

    ``stride == 1, d.block == ds``: filter at the original rate, keep every ds-th output.
    ``stride == ds, d.block == 1``: the reference's order (decimate, then filter).
    """
    n_in = x.shape[0]
    n_dec = (n_in + stride - 1) // stride            # samples the filter sees
    blk = d.block
    m = (n_dec + blk - 1) // blk                     # outputs kept
    le = n_dec + 2 * PADLEN
    E = PADLEN + blk * np.arange(m, dtype=np.int64)
    # --- contraction (bulk kernel): Uf_j, Ub0_j for j < m-1
    uf = np.zeros((m, 4))
    ub0 = np.zeros((m, 4))
    for j in range(m - 1):
        seg = ext_sample(x, n_dec, stride, E[j] + np.arange(blk + 1))
        uf[j] = seg[:blk] @ d.wf
        ub0[j] = seg @ d.q
    # --- pre-block: 15 padded samples from the steady state
    s = d.zi * float(ext_sample(x, n_dec, stride, 0))
    for e in range(PADLEN):
        s, _ = _df2t_step(d.sos, s, float(ext_sample(x, n_dec, stride, e)))
    # --- forward low-rate scan
    sf = np.zeros((m, 4))
    sf[0] = s
    for j in range(m - 1):
        sf[j + 1] = d.Ad @ sf[j] + uf[j]
    # --- tail: serial forward from E[m-1] to the end, then backward down to E[m-1]
    lt = le - E[m - 1]
    yf_tail = np.zeros(lt)
    s = sf[m - 1].copy()
    for k in range(lt):
        s, yf_tail[k] = _df2t_step(d.sos, s, float(ext_sample(x, n_dec, stride, E[m - 1] + k)))
    sb = d.zi * yf_tail[-1]
    for k in range(lt - 1, 0, -1):
        sb, _ = _df2t_step(d.sos, sb, yf_tail[k])
    # --- backward low-rate scan
    sbv = np.zeros((m, 4))
    sbv[m - 1] = sb
    for j in range(m - 2, -1, -1):
        sbv[j] = d.Ad @ sbv[j + 1] + d.P @ sf[j] + ub0[j]
    xe = ext_sample(x, n_dec, stride, E)
    yf = sf @ d.C + d.D * xe
    return sbv @ d.C + d.D * yf
