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


# ----------------------------------------------------------------------------- csrc/sosfilt.cu
SS_CHUNK, SS_THREADS, SS_HALO = 8, 256, 64
SS_TILE = SS_CHUNK * SS_THREADS


def _matpow(M, e):
    R = np.eye(4)
    Bm = M.copy()
    while e:
        if e & 1:
            R = R @ Bm
        Bm = Bm @ Bm
        e >>= 1
    return R


def _sos_scan_pass(u: np.ndarray, n_scan: int, d: BlockFilterDesign, part: int, ppart: np.ndarray, lookback: int,
                   s_init: np.ndarray):
    """One direction of k_sos_scan over the inputs ``u[0 .. n_scan)`` (scan order), tile by tile, the
    way the CTAs do it: chunk zero-state responses, inclusive combination inside the tile with the
    tabulated powers, aggregate after ``part`` samples, look-back over ``lookback`` aggregates.
    Returns one array of tile outputs per tile (SS_TILE scan positions each, zeros past n_scan)."""
    pw = d.pow_chunk                                  # A^(8 2^k)
    tiles = (n_scan + part - 1) // part
    aggs = np.zeros((tiles, 4))
    outs = []
    for b in range(tiles):
        t0 = b * part
        x = np.zeros(SS_TILE)
        nv = min(SS_TILE, n_scan - t0)
        x[:nv] = u[t0:t0 + nv]
        # sweep 1: zero-state response of every chunk
        z = np.zeros((SS_THREADS, 4))
        for t in range(SS_THREADS):
            s = np.zeros(4)
            for c in range(SS_CHUNK):
                s, _ = _df2t_step(d.sos, s, x[t * SS_CHUNK + c])
            z[t] = s
        # inclusive scan inside each warp (shuffle steps with A^(8 2^k))
        zi = z.copy()
        for k in range(5):
            o = 1 << k
            prev = zi.copy()
            for t in range(SS_THREADS):
                if (t & 31) >= o:
                    zi[t] = pw[k] @ prev[t - o] + prev[t]
        # warp totals -> zero-start state at the end / start of every warp
        T = zi[31::32].copy()
        for k in range(3):
            o = 1 << k
            prev = T.copy()
            for w in range(8):
                if w >= o:
                    T[w] = pw[5 + k] @ prev[w - o] + prev[w]
        pre0 = np.vstack([np.zeros((1, 4)), T[:-1]])
        pc = part // SS_CHUNK
        if part == SS_TILE:
            ag = T[7]
        else:
            wq, nq = (pc - 1) >> 5, ((pc - 1) & 31) + 1
            v = pre0[wq].copy()
            for k in range(6):
                if (nq >> k) & 1:
                    v = pw[k] @ v
            ag = v + zi[pc - 1]
        aggs[b] = ag
        k0 = b - lookback if b > lookback else 0
        start = s_init.copy() if k0 == 0 else np.zeros(4)
        for t in range(k0, b):
            start = ppart @ start + aggs[t]
        # sweep 2
        y = np.zeros(SS_TILE)
        for w in range(8):
            sp = _matpow(pw[5], w) @ start + pre0[w]
            for l in range(32):
                t = 32 * w + l
                st = d.pow_lane[l] @ sp + (zi[t - 1] if l else 0.0)
                for c in range(SS_CHUNK):
                    st, y[t * SS_CHUNK + c] = _df2t_step(d.sos, st, x[t * SS_CHUNK + c])
        outs.append(y)
    return outs


def sos_filtfilt_envelope(x_kept: np.ndarray, d: BlockFilterDesign, env_window: int):
    """Model of sosfilt_run (decimate-first order, ``d.block == 1``) on the already decimated
    samples: returns (filtered, envelope, absmax) as the two k_sos_scan launches produce them."""
    assert d.block == 1 and env_window - 1 <= SS_HALO
    m = len(x_kept)
    n_ext = m + 2 * PADLEN
    xe = ext_sample(np.asarray(x_kept), m, 1, np.arange(n_ext))
    K = d.lookback_tiles
    # forward: part = tile
    outs = _sos_scan_pass(xe, n_ext, d, SS_TILE, d.pow_chunk[8], K, d.zi * xe[0])
    yf = np.concatenate(outs)[:n_ext]
    # backward over y_f reversed, down to the first real sample, aggregates SS_TILE - SS_HALO apart
    part = SS_TILE - SS_HALO
    ppart = d.pow_chunk[7]
    for k in (6, 5, 4, 3):
        ppart = ppart @ d.pow_chunk[k]
    n_scan = m + PADLEN
    Kb = K if K > 900000000 else K + K // 16 + 1
    outs = _sos_scan_pass(yf[::-1], n_scan, d, part, ppart, Kb, d.zi * yf[-1])
    w = env_window
    off = (w - 1) // 2
    left = w - 1 - off
    y = np.full(m, np.nan)
    env = np.full(m, np.nan)
    for b, yt in enumerate(outs):
        t0 = b * part
        jhi = m + (PADLEN - 1) - t0
        jlo = jhi - (SS_TILE - 1)
        tile = np.zeros(SS_TILE)                      # ascending tile-local index la; j = jlo + la
        for sl in range(SS_TILE):
            j = jhi - sl
            tile[SS_TILE - 1 - sl] = yt[sl] if 0 <= j < m else 0.0
        for sl in range(part):
            j = jhi - sl
            if 0 <= j < m:
                assert np.isnan(y[j])
                y[j] = tile[SS_TILE - 1 - sl]
        # envelope of the outputs this tile owns: per thread a sliding sum over the staged |y| (zeros
        # outside the recording and beyond the tile), restarted every SS_CHUNK outputs
        la_min, la_max = max(0, -jlo), min(SS_TILE - 1, m - 1 - jlo)
        av = np.concatenate([np.abs(tile), np.zeros(SS_HALO + 1)])
        q0 = SS_TILE - part - off
        own_lo = max(la_min, q0)
        own_hi = la_max if b == 0 else min(la_max, SS_TILE - 1 - off)
        for t in range(SS_THREADS):
            qb = q0 + t * SS_CHUNK
            if qb + SS_CHUNK - 1 < own_lo or qb > own_hi:
                continue
            lb = qb - left
            assert lb >= 0
            s = 0.0
            for q in range(w):
                s += av[lb + q]
            for k in range(SS_CHUNK):
                la = qb + k
                cnt = min(la_max, la + off) - max(la_min, la - left) + 1
                if own_lo <= la <= own_hi:
                    j = jlo + la
                    assert np.isnan(env[j])
                    env[j] = s * (1.0 / w) if cnt == w else s / cnt
                s = (s + av[lb + k + w]) - av[lb + k]
    assert not np.any(np.isnan(y)) and not np.any(np.isnan(env))
    return y, env, float(np.max(np.abs(y)))


# ----------------------------------------------------------------------------- csrc/peaks.cu: k_distance_tiles
def distance_tiles_model(pos: np.ndarray, val: np.ndarray, d: int, own: int = 1024, halo: int = 256):
    """scipy's distance rule (highest priority first removes neighbours closer than d) the way
    k_distance_tiles resolves it: per tile of `own` candidates staged with `halo` neighbours on both
    sides, fix-point of  kept(k) <=> no kept higher-priority neighbour within d  where a candidate
    may only be KEPT if its whole neighbourhood is staged; leftovers (state 3) are finished by the
    global fix-point of the last CTA.  Equal values: the later index has the higher priority.
    Returns (kept mask, number of candidates the tiles left pending)."""
    nc = len(pos)
    state = np.full(nc, 3, dtype=np.int8)                      # 0 removed, 1 kept, 3 pending
    if d <= 1:
        return np.ones(nc, dtype=bool), 0
    for k0 in range(0, nc, own):
        k1 = min(nc, k0 + own)
        s0, s1 = max(0, k0 - halo), min(nc, k1 + halo)
        p, v = pos[s0:s1], val[s0:s1]
        L = s1 - s0
        st = np.zeros(L, dtype=np.int8)                        # 0 open, 1 kept, 2 removed
        open_left, open_right = s0 > 0, s1 < nc
        while True:
            changed = False
            prev = st.copy()                                   # rounds are lock-step on the device only between barriers;
            for k in range(L):                                 # reading fresher states than `prev` is harmless (monotone)
                if prev[k] != 0:
                    continue
                any_keep = any_open = False
                k2 = k - 1
                while k2 >= 0 and p[k] - p[k2] < d:
                    if v[k2] > v[k]:
                        any_keep |= prev[k2] == 1
                        any_open |= prev[k2] == 0
                    k2 -= 1
                k2 = k + 1
                while k2 < L and p[k2] - p[k] < d:
                    if v[k2] >= v[k]:
                        any_keep |= prev[k2] == 1
                        any_open |= prev[k2] == 0
                    k2 += 1
                full = (not open_left or p[k] - p[0] >= d) and (not open_right or p[-1] - p[k] >= d)
                if any_keep:
                    st[k], changed = 2, True
                elif not any_open and full:
                    st[k], changed = 1, True
            if not changed:
                break
        for k in range(k0, k1):
            s = st[k - s0]
            state[k] = 1 if s == 1 else (0 if s == 2 else 3)
    n_pending = int(np.count_nonzero(state == 3))
    while np.any(state == 3):
        prev = state.copy()
        for k in np.flatnonzero(prev == 3):
            any_keep = any_open = False
            for rng in (range(k - 1, -1, -1), range(k + 1, nc)):
                for k2 in rng:
                    if abs(pos[k2] - pos[k]) >= d:
                        break
                    if val[k2] > val[k] or (val[k2] == val[k] and k2 > k):
                        any_keep |= prev[k2] == 1
                        any_open |= prev[k2] == 3
            if any_keep:
                state[k] = 0
            elif not any_open:
                state[k] = 1
    return state == 1, n_pending


def warp_first_knot_ge_model(t: np.ndarray, bound: int) -> int:
    """Serial model of ``warp_first_knot_ge`` (csrc/floor.cu): first index r in [0, T] with t[r] >= bound,
    narrowed 32 ways per step the way the warp does it (lane l probes lo + l * step; lanes past `hi`
    count as "reached").  Returns (index, dependent probe rounds)."""
    T = len(t)
    lo, hi, rounds = 0, T, 0
    while hi - lo > 32:
        rounds += 1
        step = (hi - lo + 31) // 32
        ge = [(lo + l * step >= hi) or (int(t[lo + l * step]) >= bound) for l in range(32)]
        f = ge.index(True) if any(ge) else 32
        if f == 0:
            return lo, rounds
        lo, hi = lo + (f - 1) * step + 1, min(hi, lo + f * step)
    rounds += 1
    ge = [(lo + l >= hi) or (int(t[lo + l]) >= bound) for l in range(32)]
    return (lo + ge.index(True)) if any(ge) else hi, rounds


def select_groups_model(prefixes, active, first: bool):
    """Model of ``sel_groups`` (csrc/select.cu): levels with equal resolved prefixes share one histogram.
    -> (representative level of every group, group of every level or -1)."""
    lvl, of = [], []
    for l, (p, a) in enumerate(zip(prefixes, active)):
        if not a:
            of.append(-1)
            continue
        g = next((h for h, r in enumerate(lvl) if first or prefixes[r] == p), None)
        if g is None:
            lvl.append(l)
            g = len(lvl) - 1
        of.append(g)
    return lvl, of
