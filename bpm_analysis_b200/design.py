"""Host-side filter design for the zero-phase band-pass (K1).

Parameter setup only -- a few hundred floats per (sample-rate, cut-off,
decimation) triple, cached -- never touches audio samples.

The reference designs its filter with ``scipy.signal.butter(2, [lo, hi], 'band')``
and runs it with ``filtfilt`` (``bpm_analysis.py:1044-1045``): odd extension by
15 samples, steady-state initial conditions.  The GPU
path evaluates the same forward-backward recurrence in *blocked* form so that
it parallelises and so that, when the filter runs at the original sample rate
and only every ``ds``-th output is kept, the per-sample work is a short dot
product instead of a serial recurrence:

  state-space of the SOS cascade (direct-form II transposed coordinates)
        s[e+1] = A s[e] + B x[e],      y[e] = C s[e] + D x[e]
  forward, one block of ``ds`` samples starting at extended index E_j
        s_f[E_{j+1}] = A^ds s_f[E_j] + sum_{l<ds}  wf[l] x[E_j+l]
  backward (state before consuming y_f[e], running downward), same block
        s_b[E_j] = A^ds s_b[E_{j+1}] + P s_f[E_j] + sum_{l<=ds} q[l] x[E_j+l]
  output at the kept sample
        y[E_j] = C s_b[E_j] + D (C s_f[E_j] + D x[E_j])

``wf``/``q``/``P`` are impulse responses of the cascade and are tabulated here
(in extended precision, rounded once to float64).  With ``ds == 1`` (the
reference's decimate-then-filter order) the same formulas reduce to the plain
per-sample recurrence.  The low-rate recurrences over ``j`` are then scanned in
parallel on the device; for that the tables also carry the powers of ``A^ds``
the scan combines partial states with.
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass

import numpy as np

PADLEN = 15          # 3 * max(len(a), len(b)) for a 4th-order ba filter; same for its SOS form
SCAN_CHUNK = 8       # low-rate steps each thread runs serially   (must match csrc/sosfilt.cu)
SCAN_THREADS = 256   # threads per scan tile                       (must match csrc/sosfilt.cu)
SCAN_TILE = SCAN_CHUNK * SCAN_THREADS
N_POW = 16           # powers of A^ds tabulated: (A^ds)^(CHUNK * 2^k), k = 0..N_POW-1


def butter_bandpass_sos(low: float, high: float) -> np.ndarray:
    """2nd-order Butterworth band-pass as two biquads, ``[[b0,b1,b2,1,a1,a2]] * 2``.

    ``low``/``high`` are fractions of Nyquist, as in ``butter(2, [low, high], 'band')``
    (``bpm_analysis.py:1038-1044``).  Bilinear transform with pre-warping; the four
    zeros sit at z=+1 (x2) and z=-1 (x2); each conjugate pole pair becomes one section,
    pole pair nearer the unit circle last (the usual SOS ordering).
    """
    if not (0.0 < low < high < 1.0):
        raise ValueError("band edges must satisfy 0 < low < high < 1 (fractions of Nyquist)")
    fs = 2.0
    w1 = 2.0 * fs * math.tan(math.pi * low / fs)
    w2 = 2.0 * fs * math.tan(math.pi * high / fs)
    bw, w0 = w2 - w1, math.sqrt(w1 * w2)
    proto = np.array([np.exp(1j * math.pi * (2 * k + 3) / 4.0) for k in range(2)])   # N=2, LHP
    half = proto * bw / 2.0
    disc = np.sqrt(half * half - w0 * w0)
    p_analog = np.concatenate([half + disc, half - disc])
    k_analog = bw ** 2
    fs2 = 2.0 * fs
    p_z = (fs2 + p_analog) / (fs2 - p_analog)
    # two zeros at s=0 -> z=+1, two at infinity -> z=-1
    k_z = k_analog * np.real(fs2 ** 2 / np.prod(fs2 - p_analog))
    upper = sorted([p for p in p_z if p.imag > 0], key=lambda p: abs(p))
    if len(upper) != 2:
        raise ValueError("degenerate band-pass design (real poles); widen the band")
    secs = []
    # pair each pole pair with one zero at +1 and one at -1: numerator (1 - z^-2)
    for idx, p in enumerate(upper):
        a1, a2 = -2.0 * p.real, abs(p) ** 2
        g = k_z if idx == 0 else 1.0
        secs.append([g, 0.0, -g, 1.0, a1, a2])
    return np.asarray(secs, dtype=np.float64)


def sos_to_ba(sos: np.ndarray):
    b, a = np.array([1.0]), np.array([1.0])
    for s in sos:
        b, a = np.convolve(b, s[:3]), np.convolve(a, s[3:])
    return b, a


def cascade_state_space(sos: np.ndarray, dtype=np.longdouble):
    """(A, B, C, D) of the cascade in DF2T state coordinates (s1a, s2a, s1b, s2b)."""
    ns = sos.shape[0]
    n = 2 * ns
    A = np.zeros((n, n), dtype=dtype)
    B = np.zeros(n, dtype=dtype)
    C = np.zeros(n, dtype=dtype)
    D = dtype(1.0)
    for k, (b0, b1, b2, _, a1, a2) in enumerate(np.asarray(sos, dtype=dtype)):
        r = 2 * k
        Ak = np.array([[-a1, 1.0], [-a2, 0.0]], dtype=dtype)
        Bk = np.array([b1 - a1 * b0, b2 - a2 * b0], dtype=dtype)
        # input of this section = C s + D x of everything before it
        A[r:r + 2, :] = np.outer(Bk, C)
        A[r:r + 2, r:r + 2] = Ak
        B[r:r + 2] = Bk * D
        C = C * b0
        C[r] += 1.0
        D = D * b0
    return A, B, C, D


@dataclass(frozen=True)
class BlockFilterDesign:
    """Everything the device needs for one (filter, block length) pair, float64."""
    block: int                 # ds (1 in the reference's decimate-first order)
    sos: np.ndarray            # (2, 6)
    zi: np.ndarray             # (4,)   steady state for a unit step
    C: np.ndarray              # (4,)
    D: float
    Ad: np.ndarray             # (4, 4) A^block
    P: np.ndarray              # (4, 4)
    wf: np.ndarray             # (block, 4)
    q: np.ndarray              # (block + 1, 4)
    pow_chunk: np.ndarray      # (N_POW, 4, 4)  Ad^(SCAN_CHUNK * 2^k)
    pow_lane: np.ndarray       # (32, 4, 4)     Ad^(SCAN_CHUNK * l), l = 0..31 (start state of lane l's chunk)
    lookback_tiles: int        # how many preceding scan tiles still matter at 1e-22
    spectral_radius: float     # of A (per input sample)

    def packed(self) -> np.ndarray:
        """Flat float64 image in the layout ``BpmFilterDesign`` (include/bpm_b200.h) reads."""
        head = np.array([float(self.block), float(self.lookback_tiles), self.D, 0.0])
        # wq8[l] = (wf[l] | q[l]) interleaved, wf padded with a zero row: the image the
        # full-rate contraction kernel copies into constant memory
        wq8 = np.concatenate([np.vstack([self.wf, np.zeros((1, 4))]), self.q], axis=1)
        return np.concatenate([head, self.sos.ravel(), self.zi, self.C, self.Ad.ravel(),
                               self.P.ravel(), self.pow_chunk.ravel(),
                               self.wf.ravel(), self.q.ravel(), wq8.ravel(),
                               self.pow_lane.ravel()]).astype(np.float64)


def _matpow(M, e: int):
    R = np.eye(M.shape[0], dtype=M.dtype)
    Bm = M.copy()
    while e:
        if e & 1:
            R = R @ Bm
        Bm = Bm @ Bm
        e >>= 1
    return R


@functools.lru_cache(maxsize=256)
def design_block_filter(low: float, high: float, block: int) -> BlockFilterDesign:
    """Tabulate the blocked forward-backward band-pass (see module docstring)."""
    if block < 1:
        raise ValueError("block must be >= 1")
    ld = np.longdouble
    sos = butter_bandpass_sos(low, high)
    A, B, C, D = cascade_state_space(sos, ld)
    n = A.shape[0]
    zi = np.linalg.solve((np.eye(n) - A.astype(np.float64)), B.astype(np.float64))
    # refine zi in extended precision (one Newton step on the linear system)
    r = B - (np.eye(n, dtype=ld) - A) @ zi.astype(ld)
    zi = (zi.astype(ld) + np.linalg.solve((np.eye(n) - A.astype(np.float64)), r.astype(np.float64)).astype(ld))
    # powers: AkB[k] = A^k B, CAk[k] = C A^k
    AkB = np.zeros((block + 1, n), dtype=ld)
    CAk = np.zeros((block + 1, n), dtype=ld)
    v, c = B.copy(), C.copy()
    for k in range(block + 1):
        AkB[k], CAk[k] = v, c
        v, c = A @ v, c @ A
    Ad = _matpow(A, block)
    wf = AkB[:block][::-1].copy()                       # wf[l] = A^(block-1-l) B
    h = np.array([C @ AkB[k] for k in range(block)], dtype=ld)   # h[k] = C A^k B
    P = np.zeros((n, n), dtype=ld)
    q = np.zeros((block + 1, n), dtype=ld)
    for m in range(1, block + 1):
        g = AkB[m - 1]                                  # A^(m-1) B  multiplies y_f[E+m]
        P += np.outer(g, CAk[m])
        q[m] += g * D
        for l in range(0, m):
            q[l] += g * h[m - 1 - l]
    exps = [SCAN_CHUNK * (1 << k) for k in range(N_POW)]
    pows = np.stack([_matpow(Ad, e) for e in exps])
    lane_pows = np.stack([_matpow(Ad, SCAN_CHUNK * l) for l in range(32)])
    rho = float(np.max(np.abs(np.linalg.eigvals(A.astype(np.float64)))))
    per_tile = rho ** (block * SCAN_TILE)
    if per_tile <= 0.0:
        look = 1
    else:
        look = max(1, int(math.ceil(math.log(1e-22) / math.log(per_tile)))) if per_tile < 1.0 else 1 << 30
    f = lambda x: np.asarray(x, dtype=np.float64)
    return BlockFilterDesign(block=int(block), sos=sos, zi=f(zi), C=f(C), D=float(D), Ad=f(Ad), P=f(P),
                             wf=f(wf), q=f(q), pow_chunk=f(pows), pow_lane=f(lane_pows), lookback_tiles=int(min(look, 1 << 30)),
                             spectral_radius=rho)
