"""Long-stream chunking on the GPU: ranks as threads of one process sharing cuda:0
(stream.ThreadComm), every numeric step through libbpm_b200 (stream.DeviceEngine)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def engine():
    import torch
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    from bpm_analysis_b200 import _native, stream
    _native.load_library()
    return stream.DeviceEngine()


def _run(world, engine, pcm, sr, params):
    import torch
    from bpm_analysis_b200 import stream

    def body(comm):
        fe = stream.ChunkedFrontEnd(len(pcm), sr, params, comm, engine)
        f0, f1 = fe.frames()
        out = fe.run(engine.tensor(pcm[f0:f1]))
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    return stream.run_thread_world(world, body)


@pytest.mark.parametrize("mode", ["parity", "fullrate"])
@pytest.mark.parametrize("world", [2, 3])
def test_chunked_stream_matches_oracle(world, mode, engine, ref_params):
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    params = dict(ref_params, filter_mode=mode)
    pcm, sr, _ = synth.config_c2(seed=11, duration_sec=420.0, sample_rate=48000)
    ref = ref_port.front_end(pcm, sr, params)
    for got in _run(world, engine, pcm, sr, params):
        assert rel_err(got["envelope"], ref["envelope"]) < TOL
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
        assert rel_err(got["smoothed_dev"], np.asarray(ref["smoothed_dev_series"])) < TOL


@pytest.mark.parametrize("name", ["vulpine", "synth_c2_240s", "synth_holter_180s"])
@pytest.mark.parametrize("world", [2, 5])
def test_chunked_analysis_bit_exact_on_reference_envelope(world, name, engine, ref_params):
    """Fed the reference's own envelope, the chunked floor / troughs / peaks equal the
    UNCHUNKED GPU result bit for bit, and the reference's indices exactly."""
    import torch
    from bpm_analysis_b200 import runtime, stream
    g = load_golden(name)
    env, rate = g["envelope"], int(g["rate"])
    one_floor, one_troughs = runtime.ops().noise_floor(env, rate, ref_params)

    def body(comm):
        fe = stream.ChunkedFrontEnd.for_envelope(len(env), rate, ref_params, comm, engine)
        out = fe.analyse(engine.tensor(env))
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    for got in stream.run_thread_world(world, body):
        assert np.array_equal(got["troughs"], g["troughs"]) and np.array_equal(got["troughs"], one_troughs)
        assert np.array_equal(got["peaks"], g["raw_peaks"])
        assert np.array_equal(got["floor"], one_floor)
        assert rel_err(got["floor"], g["floor"]) < TOL


def test_chunked_stream_dropouts_and_few_troughs(engine, ref_params):
    """Holter-style stream with bursts / dropouts, and a stream too quiet to have 5 troughs
    (bpm_analysis.py:1073-1077: constant floor, all troughs returned)."""
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    pcm, sr, _ = synth.config_c4(seed=4, duration_sec=1500.0)
    ref = ref_port.front_end(pcm, sr, ref_params)
    for got in _run(4, engine, pcm, sr, ref_params):
        assert rel_err(got["envelope"], ref["envelope"]) < TOL
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
    rng = np.random.default_rng(3)
    t = np.arange(8 * 4000) / 4000.0
    quiet = (2000.0 * np.exp(-0.5 * ((t - 4.0) / 1.5) ** 2) * np.sin(2 * np.pi * 40 * t)).astype(np.int16)
    quiet[::997] += rng.integers(-1, 2, size=len(quiet[::997])).astype(np.int16)
    ref = ref_port.front_end(quiet, 4000, ref_params)
    for got in _run(2, engine, quiet, 4000, ref_params):
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL
