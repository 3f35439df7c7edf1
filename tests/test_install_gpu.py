"""`frontend.install(ref)` under the reference's OWN orchestrator, on the GPU.

The unmodified reference travels to the GPU box in ``baseline/_ref`` (``baseline/install_ref.sh``).
One copy of the module stays untouched and runs the reference's stages on the host; a second,
independent copy gets every hot-path name rebound by ``install`` (GPU front end, compiled
classifier and correction passes) and runs the SAME reference functions --
``preprocess_audio`` .. ``_run_preliminary_pass`` .. ``PeakClassifier`` ..
``_refine_and_correct_peaks`` .. ``_calculate_final_metrics`` (bpm_analysis.py:1731-1757).
Every output is compared: float signals to 1e-9 of their maximum, index lists, beat lists, debug
strings and reductions exactly (the a2..a8 stages are fed the installed copy's own envelope, so
the comparison also covers "GPU envelope -> GPU peaks" end to end)."""
import inspect

import numpy as np
import pandas as pd
import pytest
from scipy.io import wavfile

from baseline import ref_loader
from conftest import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="run baseline/install_ref.sh")]

TOL = 1e-9


def _pipeline(mod, path, params, out_dir, env_override=None):
    """analyze_wav_file's stages 1-6 (bpm_analysis.py:1731-1757) on module ``mod``."""
    env, rate = mod.preprocess_audio(path, params, out_dir)
    used = env if env_override is None else env_override
    floor, troughs = mod._calculate_dynamic_noise_floor(used, rate, params)
    start_bpm, peak_t, rec_t = mod._run_preliminary_pass(used, rate, params, floor, troughs, None)
    clf = mod.PeakClassifier(used, rate, params, start_bpm, floor, troughs, peak_t, rec_t)
    s1, raw, data = clf.classify_peaks()
    final, data = mod._refine_and_correct_peaks(s1, raw, data, used, rate, params)
    metrics = mod._calculate_final_metrics(final, rate, params) if len(final) >= 2 else None
    return dict(env=env, rate=rate, floor=floor, troughs=troughs, start_bpm=start_bpm, peak_t=peak_t, rec_t=rec_t,
                s1=s1, raw=raw, final=final, data=data, metrics=metrics, smoothed_dev=clf.state["smoothed_dev_series"])


def _same_records(a, b):
    assert (a is None) == (b is None)
    if a is None:
        return
    recs_a, recs_b = (a, b) if isinstance(a, list) else ([a], [b])
    assert len(recs_a) == len(recs_b)
    for ra, rb in zip(recs_a, recs_b):
        assert list(ra) == list(rb)
        for k in ra:
            va, vb = ra[k], rb[k]
            if isinstance(va, (float, np.floating)):
                assert va == pytest.approx(vb, rel=1e-12, abs=1e-12), k
            else:
                assert va == vb, k


@pytest.mark.parametrize("case", ["ramp_8k", "c1_44k"])
def test_reference_orchestrator_with_gpu_front_end(case, tmp_path):
    import torch
    assert torch.cuda.is_available()
    from bpm_analysis_b200 import dropin, frontend, synth
    if case == "ramp_8k":
        pcm, sr, _ = synth.pcg_recording(150.0, 8000, lambda t: 65.0 + 70.0 * np.exp(-((t - 60.0) / 25.0) ** 2), 77,
                                         noise_sigma=0.12)
    else:
        pcm, sr, _ = synth.config_c1(seed=5, duration_sec=90.0)
    path = str(tmp_path / "rec.wav")
    wavfile.write(path, sr, pcm)
    out_ref, out_gpu = tmp_path / "ref", tmp_path / "gpu"
    out_ref.mkdir(), out_gpu.mkdir()
    params = ref_loader.default_params()
    params["save_filtered_wav"] = False

    ref = ref_loader.load(fresh=True)
    want = _pipeline(ref, path, params, str(out_ref))

    mod = ref_loader.load(fresh=True)
    sigs = {n: inspect.signature(getattr(mod, n)) for n in ("preprocess_audio", "_calculate_dynamic_noise_floor",
                                                            "calculate_hrr", "find_recovery_phase")}
    frontend.install(mod)
    for n, sg in sigs.items():
        assert getattr(mod, n) is getattr(frontend, n)
        assert list(inspect.signature(getattr(mod, n)).parameters) == list(sg.parameters), n
    d = dropin.dropin()
    before = dict(d.stats)
    # (1) the installed module on the reference's OWN envelope: indices / beats / strings exact
    got = _pipeline(mod, path, params, str(out_gpu), env_override=want["env"])
    assert got["rate"] == want["rate"]
    assert rel_err(got["env"], want["env"]) < TOL
    assert rel_err(got["floor"].values, want["floor"].values) < TOL
    assert isinstance(got["floor"], pd.Series) and got["floor"].index.equals(want["floor"].index)
    assert np.array_equal(got["troughs"], want["troughs"]) and got["troughs"].dtype == want["troughs"].dtype
    assert np.array_equal(got["raw"], want["raw"])
    assert (got["start_bpm"], got["peak_t"], got["rec_t"]) == (want["start_bpm"], want["peak_t"], want["rec_t"])
    assert np.array_equal(got["s1"], want["s1"]) and np.array_equal(got["final"], want["final"])
    assert len(want["final"]) > 50
    assert list(got["data"]["beat_debug_info"].items()) == list(want["data"]["beat_debug_info"].items())
    a, b = got["smoothed_dev"], want["smoothed_dev"]
    assert np.array_equal(a.index.values, b.index.values) and rel_err(a.values, b.values) < TOL
    gm, wm = got["metrics"], want["metrics"]
    assert gm["smoothed_bpm"].index.equals(wm["smoothed_bpm"].index)
    assert rel_err(gm["smoothed_bpm"].values, wm["smoothed_bpm"].values) < TOL
    assert np.array_equal(gm["bpm_times"], wm["bpm_times"])
    for k in ("major_inclines", "major_declines", "hrr_stats", "peak_recovery_stats", "peak_exertion_stats"):
        _same_records(gm[k], wm[k])
    assert list(gm["windowed_hrv_df"].columns) == list(wm["windowed_hrv_df"].columns)
    assert rel_err(gm["windowed_hrv_df"].values, wm["windowed_hrv_df"].values) < TOL
    assert gm["hrv_summary"].keys() == wm["hrv_summary"].keys()
    for k in wm["hrv_summary"]:
        assert gm["hrv_summary"][k] == pytest.approx(wm["hrv_summary"][k], rel=1e-9)

    # (1b) the on-disk outputs: the installed module's ReportGenerator is reports.py's;
    #      its files equal the untouched module's, the time-stamp lines aside
    import re
    from bpm_analysis_b200 import reports
    assert mod.ReportGenerator is reports.ReportGenerator and ref.ReportGenerator is not reports.ReportGenerator
    stamp = re.compile(rb"(Generated on: |Analysis performed on: )[0-9: -]+")
    for m_, out_dir, res in ((ref, out_ref, want), (mod, out_gpu, got)):
        rg = m_.ReportGenerator(path, str(out_dir))
        rg.save_analysis_summary(res["metrics"])
        rg.create_chronological_log(want["env"], res["rate"], res["raw"], res["data"], res["metrics"])
        rg.save_analysis_settings(None)
    for suffix in ("_Analysis_Summary.md", "_Debug_Log.md", "_Analysis_Settings.json"):
        a, b = (open(str(d_ / ("rec" + suffix)), "rb").read() for d_ in (out_ref, out_gpu))
        assert stamp.sub(b"", a) == stamp.sub(b"", b), suffix
    assert reports.bpm_plot_csv_bytes(got["metrics"]) == reports.bpm_plot_csv_bytes(want["metrics"])

    # (2) end to end on the GPU envelope (preprocess -> session -> every later call answered from it):
    #     the same lists come out, and the chain cost ONE stage-A call
    mid = dict(d.stats)
    e2e = _pipeline(mod, path, params, str(out_gpu))
    after = dict(d.stats)
    assert after["stage_a_calls"] - mid["stage_a_calls"] == 1
    assert after["session_misses"] == mid["session_misses"], "a later call re-uploaded the envelope"
    assert np.array_equal(e2e["troughs"], want["troughs"]) and np.array_equal(e2e["raw"], want["raw"])
    assert np.array_equal(e2e["final"], want["final"])
    assert rel_err(e2e["floor"].values, want["floor"].values) < TOL
    assert before["stage_a_calls"] <= mid["stage_a_calls"]
