"""Variants of the full-rate contraction kernel (k_contract_i16) side by side on the C2 recording:
average launch time from the per-kernel event pass and the band-passed output against variant 0.

  python tools/fullrate_probe.py 0 2 3 4 ...      (BPM_CONTRACT_VARIANT values, one subprocess each)
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(variant: int) -> None:
    import numpy as np
    import torch
    from bpm_analysis_b200 import synth
    from bpm_analysis_b200.params import default_params
    from bpm_analysis_b200.runtime import StageARunner, profile_kernels
    p = default_params()
    p["save_filtered_wav"] = False
    p["filter_mode"] = "fullrate"
    pcm, sr, _ = synth.config_c2(seed=0, duration_sec=3600.0, sample_rate=48000)
    A = StageARunner([len(pcm)], sr, p, want_filtered=True)
    A.upload([pcm])
    for _ in range(3):
        A.launch()
    torch.cuda.synchronize()
    prof = profile_kernels(lambda: [A.launch() for _ in range(10)])
    torch.cuda.synchronize()
    cnt, ms = prof["k_contract_i16"]
    tot = sum(v[1] for v in prof.values()) / 10
    y = A.out["filtered"].cpu().numpy()
    path = "/tmp/fullrate_probe_v0.npy"
    if variant == 0:
        np.save(path, y)
        err = 0.0
    elif os.path.exists(path):
        ref = np.load(path)
        err = float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))
    else:
        err = float("nan")
    nbytes = len(pcm) * 2 + 8 * A.total_m
    us = ms * 1e3 / cnt
    print(f"variant {variant}: k_contract_i16 {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s   step(kernels) {tot:.3f} ms   "
          f"filtered vs variant 0: {err:.2e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(int(sys.argv[2]))
    else:
        for v in [int(a) for a in sys.argv[1:]] or [0]:
            env = dict(os.environ, BPM_CONTRACT_VARIANT=str(v))
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(v)], env=env, check=True)
