"""CPU-side checks: the C-ABI library loads and exports what include/bpm_b200.h declares,
the filter design matches scipy, the blocked formulation (numpy model of the kernels) matches
scipy's filtfilt / sosfiltfilt, and host-side planning mirrors the reference's arithmetic."""
import os
import re

import numpy as np
import pytest
from scipy.signal import butter, filtfilt, sosfiltfilt

from bpm_analysis_b200 import _native, design, params as P, synth
from oracle import kernel_models

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bpm_analysis_b200.build import build_native
    build_native()
    return _native.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(REPO, "include", "bpm_b200.h")).read()
    declared = set(re.findall(r"\b(bpm_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bpm_b200.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS)
    assert lib.bpm_abi_version() == _native.ABI_VERSION
    assert lib.bpm_error_string(-2) == b"workspace too small"


def test_workspace_queries_are_monotone(lib):
    a = lib.bpm_stage_a_workspace_bytes(100_000, 1)
    b = lib.bpm_stage_a_workspace_bytes(1_000_000, 4)
    assert 0 < a < b


def test_missing_library_is_loud(tmp_path, monkeypatch):
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(_native.NativeLibraryError):
        _native.load_library(str(tmp_path / "nope.so"))


@pytest.mark.parametrize("sr,ds", [(44100, 146), (48000, 159), (4000, 12)])
def test_butterworth_matches_scipy(sr, ds):
    for rate in (sr // ds, sr):
        lo, hi = 20 / (rate / 2), 150 / (rate / 2)
        b, a = design.sos_to_ba(design.butter_bandpass_sos(lo, hi))
        b0, a0 = butter(2, [lo, hi], btype="band")
        assert np.max(np.abs(b - b0)) < 1e-12 * np.max(np.abs(b0)) + 1e-16
        assert np.max(np.abs(a - a0)) < 1e-12


def test_blocked_model_parity_mode_matches_filtfilt():
    pcm, sr, _ = synth.config_c1(seed=3, duration_sec=6.0)
    ds, rate, _ = P.effective_decimation(sr, P.default_params())
    assert (ds, rate) == (146, 302)
    lo, hi = 20 / (rate / 2), 150 / (rate / 2)
    d = design.design_block_filter(lo, hi, 1)
    y = kernel_models.blocked_filtfilt(pcm, d, stride=ds)
    b, a = butter(2, [lo, hi], btype="band")
    ref = filtfilt(b, a, pcm[::ds])
    assert y.shape == ref.shape
    assert np.max(np.abs(y - ref)) < 1e-11 * np.max(np.abs(ref))


def test_blocked_model_fullrate_matches_sosfiltfilt():
    pcm, sr, _ = synth.config_c1(seed=4, duration_sec=3.0)
    ds = 146
    lo, hi = 20 / (sr / 2), 150 / (sr / 2)
    d = design.design_block_filter(lo, hi, ds)
    y = kernel_models.blocked_filtfilt(pcm, d, stride=1)
    sos = butter(2, [lo, hi], btype="band", output="sos")
    ref = sosfiltfilt(sos, pcm.astype(np.float64), padlen=15)[::ds]
    assert y.shape == ref.shape
    assert np.max(np.abs(y - ref)) < 1e-9 * np.max(np.abs(ref))


def test_design_image_layout():
    d = design.design_block_filter(0.1, 0.9, 7)
    img = d.packed()
    assert img.shape[0] == _native.DESIGN_HEADER_WORDS + 16 * 7 + 12 + 512
    assert img[0] == 7 and img[2] == d.D
    assert np.array_equal(img[24:40].reshape(4, 4), d.Ad)
    assert np.array_equal(img[312:312 + 28].reshape(7, 4), d.wf)
    wq8 = img[312 + 4 * 15:312 + 4 * 15 + 64].reshape(8, 8)
    assert np.array_equal(wq8[:7, :4], d.wf) and np.all(wq8[7, :4] == 0) and np.array_equal(wq8[:, 4:], d.q)
    lane = img[-512:].reshape(32, 4, 4)
    assert np.array_equal(lane[0], np.eye(4)) and np.array_equal(lane, d.pow_lane)
    assert np.allclose(lane[4], d.pow_chunk[2], rtol=1e-13, atol=0)        # Ad^(8*4) == Ad^(8*2^2)


@pytest.mark.parametrize("dur,sr,w", [(12.0, 44100, None), (60.0, 4000, None), (0.2, 44100, None), (14.0, 48000, 65)])
def test_sos_scan_model_matches_filtfilt_and_rolling_mean(dur, sr, w):
    """numpy model of csrc/sosfilt.cu (tiles, partition / halo, look-back, envelope ownership) against
    scipy.signal.filtfilt + pandas' centred rolling mean (bpm_analysis.py:1044-1054)."""
    import pandas as pd
    pcm, sr, _ = synth.pcg_recording(dur, sr, lambda t: 75.0, 11)
    ds, rate, _ = P.effective_decimation(sr, P.default_params())
    lo, hi = 20 / (rate / 2), 150 / (rate / 2)
    d = design.design_block_filter(lo, hi, 1)
    x = pcm[::ds]
    w = w or rate // 10
    y, env, amax = kernel_models.sos_filtfilt_envelope(x, d, w)
    b, a = butter(2, [lo, hi], btype="band")
    ref = filtfilt(b, a, x)
    renv = pd.Series(np.abs(ref)).rolling(window=w, min_periods=1, center=True).mean().values
    assert np.max(np.abs(y - ref)) < 1e-11 * np.max(np.abs(ref))
    assert np.max(np.abs(env - renv)) < 1e-11 * np.max(renv)
    assert amax == np.max(np.abs(y))


def test_decimation_plan_matches_reference_arithmetic():
    p = P.default_params()
    assert P.effective_decimation(44100, p) == (146, 302, True)
    assert P.effective_decimation(48000, p) == (159, 301, True)
    assert P.effective_decimation(4000, p) == (12, 333, True)
    p2 = dict(p, downsample_factor=10)
    assert P.effective_decimation(44100, p2) == (10, 4410, False)
    p3 = dict(p, downsample_factor=1)
    assert P.effective_decimation(44100, p3) == (1, 44100, False)


def test_items_layout():
    pytest.importorskip("torch")
    from bpm_analysis_b200.runtime import make_items
    it = make_items([100, 50, 70], [10, 5, 7])
    assert it["in_off"].tolist() == [0, 100, 150] and it["m_off"].tolist() == [0, 10, 15]
    assert it.view(np.int64).reshape(-1, 4)[1].tolist() == [100, 50, 10, 5]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs on the CPU (the unmodified reference from baseline/_ref when it is
    installed there, else the oracle port, on the host cores) and must print
    exactly one JSON line with the contract's keys; stdout carries nothing else."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--duration-sec", "60"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio_hours_per_sec" and d["unit"] == "audio-hours/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    from baseline import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_copy_frames_rejects_bad_arguments_without_touching_the_device(lib):
    """bpm_copy_frames validates before it enqueues anything: these calls never reach CUDA."""
    import ctypes
    from bpm_analysis_b200.runtime import make_items
    items = make_items([1000], [7])                      # m must be ceil(n_in / stride) = 7 for stride 159
    buf = ctypes.create_string_buffer(64)
    ok_items = items.ctypes.data
    assert lib.bpm_copy_frames(None, 0, 1, ok_items, 1, 159, ctypes.addressof(buf), None) == -1       # null pcm
    assert lib.bpm_copy_frames(ctypes.addressof(buf), 9, 1, ok_items, 1, 159, ctypes.addressof(buf), None) == -1   # dtype
    assert lib.bpm_copy_frames(ctypes.addressof(buf), 0, 0, ok_items, 1, 159, ctypes.addressof(buf), None) == -1   # channels
    assert lib.bpm_copy_frames(ctypes.addressof(buf), 0, 1, ok_items, 1, 0, ctypes.addressof(buf), None) == -1     # stride
    bad = make_items([1000], [8])
    assert lib.bpm_copy_frames(ctypes.addressof(buf), 0, 1, bad.ctypes.data, 1, 159, ctypes.addressof(buf), None) == -1


def test_design_cache_is_keyed_by_value_and_bounded(monkeypatch):
    """More designs than design_block_filter's lru_cache holds (256): evicted design objects are
    collected and CPython reuses their id(); the device cache must still hand every plan ITS image."""
    import torch
    from bpm_analysis_b200 import runtime
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(runtime, "to_device", lambda a: torch.from_numpy(np.array(a, copy=True)))
    runtime._design_cache.clear()
    p = P.default_params()
    n = 0
    for k in range(320):
        params = dict(p, lowcut_hz=15.0 + 0.01 * k, highcut_hz=120.0 + 0.02 * k)
        plan = runtime.plan_filter(48000, params)
        dev, host = runtime.design_images(plan)
        assert np.array_equal(dev.numpy(), plan.design.packed()) and np.array_equal(host, plan.design.packed())
        n += 1
    assert n == 320 and len(runtime._design_cache) <= runtime.DESIGN_CACHE_SIZE
    # and a repeat of an early (long evicted) design is rebuilt, not confused with a newer one
    plan = runtime.plan_filter(48000, dict(p, lowcut_hz=15.0, highcut_hz=120.0))
    dev, _ = runtime.design_images(plan)
    assert np.array_equal(dev.numpy(), plan.design.packed())
    runtime._design_cache.clear()


def test_host_gather_frames_matches_numpy_slicing():
    """bpm_host_gather_frames == audio_data[::stride] (bpm_analysis.py:1033) for every frame size."""
    import ctypes as C
    from bpm_analysis_b200 import classifier
    from bpm_analysis_b200.build import build_host
    build_host()
    lib = classifier.load_host_library()
    assert lib.bpm_host_threads() >= 1
    rng = np.random.default_rng(5)
    for dtype, ch, n, stride in ((np.int16, 1, 1000003, 159), (np.int16, 2, 300001, 146), (np.uint8, 1, 70001, 12),
                                 (np.float32, 1, 99999, 7), (np.float64, 3, 4100, 3), (np.int32, 1, 17, 1),
                                 (np.int16, 1, 5, 300)):
        x = rng.integers(-100, 100, size=(n, ch)).astype(dtype)
        if ch == 1:
            x = x[:, 0].copy()
        want = x[::stride]
        out = np.empty_like(want)
        for threads in (0, 1, 3):
            out[...] = 0
            rc = lib.bpm_host_gather_frames(C.c_void_p(x.ctypes.data), x.dtype.itemsize * ch, n, stride,
                                            C.c_void_p(out.ctypes.data), threads)
            assert rc == 0 and np.array_equal(out, want)
    assert lib.bpm_host_gather_frames(None, 2, 10, 1, None, 0) != 0


def test_distance_tiles_model_matches_scipy():
    """numpy model of csrc/peaks.cu::k_distance_tiles (tile + halo fix-point, leftovers finished by the
    global fix-point) against scipy's own _select_by_peak_distance, including chains of rising
    candidates that escape every halo."""
    from scipy.signal._peak_finding_utils import _select_by_peak_distance
    rng = np.random.default_rng(1)
    pending = 0
    for trial in range(24):
        n = int(rng.integers(50, 400))
        gaps = rng.integers(1, 8, size=n) if trial % 3 else rng.integers(1, 40, size=n)
        pos = np.cumsum(gaps).astype(np.intp)
        if trial % 4 == 0:
            val = np.arange(n, dtype=float) + rng.random(n) * 0.5
        elif trial % 4 == 1:
            val = -np.arange(n, dtype=float) + rng.random(n) * 0.5
        else:
            val = rng.random(n)
        d = int(rng.integers(2, 20))
        want = _select_by_peak_distance(pos, val.copy(), float(d))
        for own, halo in ((16, 4), (1024, 256)):
            got, npend = kernel_models.distance_tiles_model(pos, val, d, own, halo)
            pending += npend
            assert np.array_equal(got, want), (trial, own, halo)
    assert pending > 0


def _write_wav24(path, samples, rate, extensible=False, junk_before_data=b""):
    """24-bit little-endian PCM WAV from int32 `samples` ([n] or [n, ch], values within 24 bits)."""
    import struct
    x = np.asarray(samples, dtype=np.int64)
    ch = 1 if x.ndim == 1 else x.shape[1]
    b = (x.reshape(-1) & 0xFFFFFF).astype("<u4").view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
    if extensible:
        guid = struct.pack("<H", 1) + bytes.fromhex("000000001000800000AA00389B71")
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, ch, rate, rate * ch * 3, ch * 3, 24, 22, 24, (1 << ch) - 1) + guid
    else:
        fmt = struct.pack("<HHIIHH", 1, ch, rate, rate * ch * 3, ch * 3, 24)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if junk_before_data:
        body += b"LIST" + struct.pack("<I", len(junk_before_data)) + junk_before_data + (b"\0" if len(junk_before_data) & 1 else b"")
    body += b"data" + struct.pack("<I", len(b)) + b + (b"\0" if len(b) & 1 else b"")
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def test_24_bit_wav_is_mapped_and_gathered_like_scipy_reads_it(tmp_path):
    """bpm_analysis.py:1014 reads a 24-bit file through scipy (int32, the 24 bits in the upper bytes);
    wav24.map_s24 + bpm_host_gather_s24 give the same values for the whole file and for every
    decimation, without expanding the file."""
    import warnings
    from scipy.io import wavfile
    from bpm_analysis_b200 import frontend, wav24
    rng = np.random.default_rng(24)
    cases = [("mono", rng.integers(-(1 << 23), 1 << 23, 50001), dict()),
             ("stereo", rng.integers(-(1 << 23), 1 << 23, (20000, 2)), dict()),
             ("ext3", rng.integers(-(1 << 23), 1 << 23, (7001, 3)), dict(extensible=True)),
             ("junk", rng.integers(-(1 << 23), 1 << 23, 12345), dict(junk_before_data=b"INFOsoftware"[:11])),
             ("edges", np.array([0, 1, -1, (1 << 23) - 1, -(1 << 23), 255, 256, -256] * 40), dict())]
    for name, x, kw in cases:
        path = str(tmp_path / f"{name}.wav")
        _write_wav24(path, x, 8000, **kw)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rate, want = wavfile.read(path)
        assert want.dtype == np.int32 and np.array_equal(want >> 8, np.asarray(x).reshape(want.shape)), name
        got = wav24.map_s24(path)
        assert got is not None, name
        assert got[0] == rate and got[1].shape == want.shape and got[1].dtype == want.dtype and len(got[1]) == len(want)
        assert np.array_equal(np.asarray(got[1]), want), name
        for stride in (1, 2, 7, 159, len(want) - 1, len(want), len(want) + 5):
            assert np.array_equal(got[1].decimated(stride), want[::stride]), (name, stride)
        r2, rec = frontend.read_wav(path)                   # the reference-facing reader takes the mapped route
        assert r2 == rate and isinstance(rec, wav24.S24Recording)
    # what the parser does not recognise is scipy's business
    p16 = str(tmp_path / "p16.wav")
    wavfile.write(p16, 8000, rng.integers(-30000, 30000, 4000).astype(np.int16))
    assert wav24.map_s24(p16) is None
    r16, a16 = frontend.read_wav(p16)
    assert r16 == 8000 and isinstance(a16, np.ndarray) and a16.dtype == np.int16
    cut = str(tmp_path / "cut.wav")
    raw = open(str(tmp_path / "mono.wav"), "rb").read()
    open(cut, "wb").write(raw[:len(raw) // 2])
    assert wav24.map_s24(cut) is None
    assert wav24.map_s24(str(tmp_path / "missing.wav")) is None
    from bpm_analysis_b200 import classifier
    assert classifier.load_host_library().bpm_host_gather_s24(None, 1, 10, 1, None, 0) != 0


def test_warp_knot_search_model_matches_searchsorted():
    """The 32-ary search the rolling-floor CTAs start with (csrc/floor.cu, warp_first_knot_ge): same
    index as np.searchsorted(side='left') for every table size and bound, in at most
    ceil(log32(T)) + 1 dependent rounds (a 13 k-knot table: 3, the bisection it replaced: 14)."""
    rng = np.random.default_rng(11)
    for T in (1, 2, 31, 32, 33, 34, 63, 64, 65, 100, 1023, 1024, 1025, 1057, 12856, 40000):
        t = np.sort(rng.choice(np.arange(0, 40 * T + 50), size=T, replace=False)).astype(np.int32)
        bounds = list(rng.integers(-5, int(t[-1]) + 10, 60)) + [int(t[0]), int(t[0]) - 1, int(t[-1]), int(t[-1]) + 1]
        bounds += [int(v) for v in t[rng.integers(0, T, 20)]] + [int(v) + 1 for v in t[rng.integers(0, T, 20)]]
        worst = 0
        for b in bounds:
            got, rounds = kernel_models.warp_first_knot_ge_model(t, int(b))
            assert got == int(np.searchsorted(t, b, side="left")), (T, b)
            worst = max(worst, rounds)
        limit = 1
        while 32 ** limit < T:
            limit += 1
        assert worst <= limit + 1, (T, worst)
    assert kernel_models.warp_first_knot_ge_model(np.arange(12856, dtype=np.int32) * 85, 500000)[1] <= 3


def test_select_level_groups_model():
    """Levels with the same resolved prefix are served by one histogram (csrc/select.cu, sel_groups)."""
    g = kernel_models.select_groups_model
    assert g([0, 0, 0], [True, True, True], first=True) == ([0], [0, 0, 0])
    assert g([7, 7, 9], [True, True, True], first=False) == ([0, 2], [0, 0, 1])
    assert g([7, 8, 7], [True, False, True], first=False) == ([0], [0, -1, 0])
    assert g([1, 2, 3], [True, True, True], first=False) == ([0, 1, 2], [0, 1, 2])
    assert g([5, 5, 5], [False, False, False], first=False) == ([], [-1, -1, -1])
    assert g([4, 6, 6], [False, True, True], first=False) == ([1], [-1, 0, 0])
