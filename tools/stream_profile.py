"""Where a step of stream.ShardedFrontEnd goes (one rank, the whole recording as its chunk):
host wall-clock per phase with a device wait after each (so the phases add up to more than a
real step, which waits once), the step's real time, and the kernel table of one step.

  python tools/stream_profile.py [hours]            (C4 generator, 4 kHz)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from bpm_analysis_b200 import stream, synth
from bpm_analysis_b200.params import default_params
from bpm_analysis_b200.runtime import profile_kernels


def main():
    hours = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    params = default_params()
    params["save_filtered_wav"] = False
    pcm, sr, _ = synth.config_c4(seed=4, duration_sec=hours * 3600.0)
    comm, eng = stream.DistComm(), stream.DeviceEngine()
    fe = stream.ShardedFrontEnd(len(pcm), sr, params, comm, eng)
    dev = eng.tensor(pcm)
    for _ in range(3):
        out = fe.run(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = fe.run(dev)
    torch.cuda.synchronize()
    print(f"step {1e3 * (time.perf_counter() - t0) / 10:.3f} ms   sharded={out['sharded']}  m={fe.chunks.m}")

    def timed(label, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        t_host = time.perf_counter() - t
        torch.cuda.synchronize()
        print(f"  {label:28s} host {1e3 * t_host:7.3f} ms   with device {1e3 * (time.perf_counter() - t):7.3f} ms")
        return r

    g = fe.geometry()
    P = params
    for _ in range(2):
        filt, env = timed("frontend", lambda: eng.frontend(dev, len(pcm), fe.plan, 1, pcm.dtype))
        core = env[g.core_lo:g.core_hi]
        one, qstat = timed("stream_quantiles", lambda: stream.stream_quantiles(eng, comm, core, fe.chunks.m,
                                                                            [float(P["trough_prominence_quantile"])]))
        thr, qs = one.repeat(2), qstat.repeat(2)
        c = timed("chunk_chain", lambda: eng.chunk_chain(env, thr, qs, g, P))
        table = timed("table", lambda: comm.all_gather_rows(c["proof"]).cpu().numpy())
        timed("deviation", lambda: eng.deviation_series(c["strength"][:int(table[0, 3])], 0.05))
    prof = profile_kernels(lambda: fe.run(dev))
    torch.cuda.synchronize()
    tot = sum(v[1] for v in prof.values())
    print(f"kernels of one step: {tot:.3f} ms")
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:28s} x{cnt:<3d} {1e3 * ms / cnt:9.1f} us   {100 * ms / tot:5.1f} %")


if __name__ == "__main__":
    main()
