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

    def frontend(self, pcm, n_in, plan, channels, np_dtype):
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
