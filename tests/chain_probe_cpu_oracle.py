#!/usr/bin/env python
"""tools/chain_probe.py with the CPU ORACLE as front end (test infrastructure: the "before" column
of the whole-chain timing, and a dry run of the probe on a machine without a GPU).

    python tests/chain_probe_cpu_oracle.py [duration_sec] [sample_rate]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))

import chain_probe                      # noqa: E402
from oracle import ref_port            # noqa: E402


class OracleFrontEnd:
    label = "cpu oracle"

    def __init__(self, sr, params):
        self.sr, self.params = sr, params
        self.metrics = [ref_port.find_major_hr_inclines, ref_port.find_major_hr_declines,
                        ref_port.find_peak_recovery_rate, ref_port.find_peak_exertion_rate]

    def preprocess(self, x):
        env, rate, _ = ref_port.preprocess_pcm(x, self.sr, self.params)
        return env, rate

    def noise_floor(self, env, rate):
        return ref_port.calculate_dynamic_noise_floor(env, rate, self.params)

    @staticmethod
    def init_state(self, hint, floor, troughs):
        peaks = ref_port.find_raw_peaks(self.audio_envelope, self.sample_rate, self.params, floor.values)
        met = ref_port.peak_metrics(self.audio_envelope, self.sample_rate, self.params, floor, peaks)
        return {"analysis_data": {}, "dynamic_noise_floor": floor, "trough_indices": troughs, "all_peaks": peaks,
                "smoothed_dev_series": met["smoothed_dev_series"], "long_term_bpm": float(hint) if hint else 80.0,
                "candidate_beats": [], "beat_debug_info": {}, "long_term_bpm_history": [],
                "consecutive_rr_rejections": 0, "loop_idx": 0}

    def bpm_series(self, beats, rate):
        return ref_port.calculate_bpm_series(beats, rate, self.params)

    def hrv(self, beats, rate):
        return ref_port.calculate_windowed_hrv(beats, rate, self.params)


if __name__ == "__main__":
    chain_probe.run(OracleFrontEnd)
