"""CPU engine for bpm_analysis_b200.stream built on the oracle (oracle/ref_port.py and the
numpy / scipy / pandas calls the reference makes).  Test infrastructure only."""
import numpy as np
import pandas as pd
import torch
from scipy.signal import find_peaks

from oracle import ref_port


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


class OracleEngine:
    def __init__(self, sample_rate, params):
        self.sample_rate, self.params = sample_rate, params

    def tensor(self, a):
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a if a.flags.writeable else a.copy())

    def full(self, n, value):
        return torch.full((n,), float(value), dtype=torch.float64)

    def frontend(self, pcm, n_in, plan, channels, np_dtype, want_filtered=True):
        x = _np(pcm)
        x = x.reshape(n_in, channels) if channels > 1 else x.reshape(n_in)
        env, _, filt = ref_port.preprocess_pcm(x, self.sample_rate, self.params)
        return self.tensor(filt), self.tensor(env)

    def quantile(self, x, q):
        return torch.tensor([np.quantile(_np(x), q)], dtype=torch.float64)

    def find_peaks(self, x, sign, height, prominence, distance):
        v = _np(x) if sign > 0 else -_np(x)
        idx, _ = find_peaks(v, height=None if height is None else _np(height),
                            prominence=None if prominence is None else float(_np(prominence)[0]),
                            distance=distance)
        return self.tensor(idx.astype(np.int64))

    def rolling_floor(self, env, knots, window, q):
        s = ref_port._interp_troughs(_np(env), _np(knots))
        f = s.rolling(window=window, min_periods=3, center=True).quantile(q).bfill().ffill()
        return self.tensor(f.values)

    def sanitize(self, env, draft, troughs, mult):
        e, d = _np(env), _np(draft)
        kept = [int(t) for t in _np(troughs) if not np.isnan(d[t]) and e[t] <= mult * d[t]]
        return self.tensor(np.asarray(kept, dtype=np.int64))

    def peak_metrics(self, env, floor, peaks, factor):
        p = dict(self.params)
        p["deviation_smoothing_factor"] = factor
        if len(peaks) < 2:
            z = torch.zeros(0, dtype=torch.float64)
            st = _np(env)[_np(peaks)] - _np(floor)[_np(peaks)]
            return self.tensor(np.maximum(st, 0)), z, z
        m = ref_port.peak_metrics(_np(env), self.sample_rate, p, pd.Series(_np(floor)), _np(peaks))
        return self.tensor(m["strength"]), self.tensor(m["deviation"]), self.tensor(m["smoothed_dev_series"].values)


# ------------------------------------------------------------------------------------------------
# Chunk mode (stream.ShardedFrontEnd): numpy / scipy restatements of the chunk operators of
# include/bpm_b200.h (bpm_key_*, bpm_noise_floor_chunk, bpm_find_peaks_chunk, bpm_chunk_proof), so
# that the planner, the proofs and the exchanges run under gloo without a GPU.
def _f64_keys(x):
    b = np.ascontiguousarray(_np(x), dtype=np.float64).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def _key_f64(k):
    k = np.uint64(k)
    b = (k & np.uint64(0x7FFFFFFFFFFFFFFF)) if (k >> np.uint64(63)) else ~k
    return float(np.array([b], dtype=np.uint64).view(np.float64)[0])


def _lerp(a, b, t):
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    d = b - a
    return float(b - d * (np.float64(1.0) - t)) if t >= 0.5 else float(a + d * t)


def _local_candidates(v, height):
    """scipy's local maxima (plateau midpoints), height condition applied"""
    idx, _ = find_peaks(v, height=height)
    return idx


def _anchors(v, cand, d, lo, hi):
    """max position <= lo / min position >= hi - 1 of a candidate that outranks every candidate within d"""
    left, right = -1, np.iinfo(np.int64).max
    for k, p in enumerate(cand):
        if not (p <= lo or p >= hi - 1):
            continue
        top = True
        j = k - 1
        while j >= 0 and p - cand[j] < d:
            if v[cand[j]] > v[p]:
                top = False
                break
            j -= 1
        j = k + 1
        while top and j < len(cand) and cand[j] - p < d:
            if v[cand[j]] >= v[p]:
                top = False
                break
            j += 1
        if top:
            if p <= lo:
                left = max(left, int(p))
            if p >= hi - 1:
                right = min(right, int(p))
    return left, right


def _edge_hits(v, survivors, thr, lo, hi, open_left, open_right):
    """Survivors of the distance rule inside [lo, hi) whose prominence test FAILED because the walk ran off
    an open end: every sample towards that end is <= the peak and less than `thr` below it (in the whole
    stream the walk would have gone on)."""
    hits = 0
    pre_max, pre_min = np.maximum.accumulate(v), np.minimum.accumulate(v)
    suf_max, suf_min = np.maximum.accumulate(v[::-1])[::-1], np.minimum.accumulate(v[::-1])[::-1]
    for p in survivors:
        if lo <= p < hi and thr > 0:
            if open_left and (p == 0 or (pre_max[p - 1] <= v[p] and v[p] - pre_min[p - 1] < thr)):
                hits += 1
            elif open_right and (p == len(v) - 1 or (suf_max[p + 1] <= v[p] and v[p] - suf_min[p + 1] < thr)):
                hits += 1
    return hits


def _proven_range(knots, lo, hi, g):
    off = (g.window - 1) // 2
    left = g.window - 1 - off
    k = knots[(knots >= lo) & (knots < hi)]
    if len(k) == 0:
        return (0, g.n) if (g.at_start and g.at_end) else (0, 0)
    return (0 if g.at_start else int(k[0]) + left), (g.n if g.at_end else int(k[-1]) - off + 1)


class OracleChunkEngine(OracleEngine):
    def key_state(self, n_total, k):
        return torch.tensor([0, int(k), int(n_total)], dtype=torch.int64)

    def key_histogram(self, x, shift, bits, state, hist):
        keys = _f64_keys(x)
        up = shift + bits
        if up < 64:
            keys = keys[(keys >> np.uint64(up)) == np.uint64(int(state[0]))]
        bins = ((keys >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
        hist += torch.from_numpy(np.bincount(bins, minlength=1 << bits).astype(np.int64))

    def key_pick(self, hist, bits, state):
        h = hist.numpy()
        c = np.cumsum(h)
        rank = int(state[1])
        b = int(np.searchsorted(c, rank, side="right"))
        state[0] = (int(state[0]) << bits) | b
        state[1] = rank - (int(c[b - 1]) if b else 0)
        state[2] = int(h[b])

    def key_collect(self, x, up_shift, state, cap):
        keys = _f64_keys(x)
        top = keys >> np.uint64(up_shift)
        pref = np.uint64(int(state[0]))
        inside, above = keys[top == pref], keys[top > pref]
        row = np.zeros(cap + 4, dtype=np.uint64)
        row[0] = len(inside)
        row[1] = above.min() if len(above) else np.uint64(0xFFFFFFFFFFFFFFFF)
        row[2] = inside.min() if len(inside) else np.uint64(0xFFFFFFFFFFFFFFFF)
        row[3] = inside.max() if len(inside) else 0
        row[4:4 + min(len(inside), cap)] = inside[:cap]
        return torch.from_numpy(row.view(np.int64).copy())

    def key_finish(self, rows, cap, state, gamma, out, status):
        r = rows.numpy().view(np.uint64)
        total, rank = int(r[:, 0].sum()), int(state[1])
        above = r[:, 1].min()
        if total == 0 or total != int(state[2]) or any(int(c) > cap for c in r[:, 0]):
            status[0] = 1
            return
        keys = np.sort(np.concatenate([row[4:4 + int(row[0])] for row in r]))
        ka = keys[rank]
        kb = keys[rank + 1] if rank + 1 < total else (above if above != np.uint64(0xFFFFFFFFFFFFFFFF) else ka)
        out[0] = _lerp(_key_f64(ka), _key_f64(kb), gamma)
        status[0] = 0

    def chunk_chain(self, env, thr, qstat, g, params):
        e = _np(env)
        n, d = len(e), g.distance
        q_t, q_p = float(thr[0]), float(thr[1])
        open_l, open_r = not g.at_start, not g.at_end
        every, _ = find_peaks(-e, prominence=q_t, distance=d)
        hits = _edge_hits(-e, find_peaks(-e, distance=d)[0], q_t, g.t_lo, g.t_hi, open_l, open_r)
        ta = _anchors(-e, _local_candidates(-e, None), d, g.t_lo, g.t_hi)
        fq = float(params["noise_floor_quantile"])

        def rolling(knots):
            if len(knots) == 0:
                return np.full(n, np.nan)
            s = ref_port._interp_troughs(e, knots)
            return s.rolling(window=g.window, min_periods=3, center=True).quantile(fq).bfill().ffill().values

        draft = rolling(every)
        mult = float(params.get("trough_rejection_multiplier", 4.0))
        kept = np.asarray([t for t in every if not np.isnan(draft[t]) and e[t] <= mult * draft[t]], dtype=np.int64)
        floor = rolling(kept)
        height = np.where(np.isnan(floor), np.inf, floor)
        peaks, _ = find_peaks(e, height=height, prominence=q_p, distance=d)
        hits += _edge_hits(e, find_peaks(e, height=height, distance=d)[0], q_p, g.core_lo, g.core_hi, open_l, open_r)
        pa = _anchors(e, _local_candidates(e, height), d, g.core_lo, g.core_hi)
        strength = np.zeros(n)
        strength[:len(peaks)] = np.maximum(e[peaks] - floor[peaks], 0.0)
        # the proof (k_chunk_proof)
        ok = hits == 0 and int(qstat[0]) == 0 and int(qstat[1]) == 0
        if d > 1:
            ok &= g.at_start or ta[0] - d >= g.filter_halo
            ok &= g.at_end or ta[1] + d <= n - g.filter_halo
        x2 = _proven_range(every, g.t_lo, g.t_hi, g)
        x3 = _proven_range(kept, max(x2[0], g.t_lo), min(x2[1], g.t_hi), g)
        ok &= x3[0] <= g.core_lo and x3[1] >= g.core_hi
        if d > 1:
            ok &= g.at_start or (pa[0] >= 0 and pa[0] - d >= x3[0])
            ok &= g.at_end or (pa[1] < n and pa[1] + d < x3[1])
        core = [g.core_lo, g.core_hi]
        (la, ha), (lk, hk), (lp, hp) = (np.searchsorted(a, core) for a in (every, kept, peaks))
        proof = torch.tensor([0 if ok else 1, ha - la, hk - lk, hp - lp, lk, lp, x3[0], x3[1]], dtype=torch.int64)

        def buf(a):
            b = np.zeros(n, dtype=np.int64)
            b[:len(a)] = a
            return torch.from_numpy(b)

        return {"floor": self.tensor(floor), "kept": buf(kept), "every": buf(every), "peaks": buf(peaks),
                "strength": self.tensor(strength), "proof": proof}

    def chunk_pack(self, c, origin, cap_t, cap_p):
        nk, npk, lk, lp = (int(v) for v in c["proof"][2:6])
        out = torch.zeros(cap_t + 2 * cap_p, dtype=torch.int64)
        out[:nk] = c["kept"][lk:lk + nk] + origin
        out[cap_t:cap_t + npk] = c["peaks"][lp:lp + npk] + origin
        out[cap_t + cap_p:cap_t + cap_p + npk] = c["strength"][lp:lp + npk].view(torch.int64)
        return out

    def chunk_unpack(self, rows, table, cap_t, cap_p, n_t, n_p):
        w = rows.shape[0]
        tr = torch.cat([rows[r, :int(table[r, 2])] for r in range(w)])
        pk = torch.cat([rows[r, cap_t:cap_t + int(table[r, 3])] for r in range(w)])
        st = torch.cat([rows[r, cap_t + cap_p:cap_t + cap_p + int(table[r, 3])] for r in range(w)]).view(torch.float64)
        assert len(tr) == n_t and len(pk) == n_p
        return tr, pk, st

    def deviation_series(self, strength, factor):
        s = _np(strength)
        if len(s) < 2:
            z = torch.zeros(0, dtype=torch.float64)
            return z, z
        dev = np.abs(np.diff(s)) / (np.maximum(s[:-1], s[1:]) + 1e-9)
        win = max(5, int(len(dev) * factor))
        sm = pd.Series(dev).rolling(window=win, min_periods=1, center=True).mean().values
        return self.tensor(dev), self.tensor(sm)
