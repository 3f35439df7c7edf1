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


# ----------------------------------------------------------------------------- every stage sharded
def _run_sharded(world, engine, pcm, sr, params, **kw):
    import torch
    from bpm_analysis_b200 import stream

    def body(comm):
        fe = stream.ShardedFrontEnd(len(pcm), sr, params, comm, engine, **kw)
        f0, f1 = fe.frames()
        out = fe.run(engine.tensor(pcm[f0:f1]), gather_series=True)
        torch.cuda.synchronize()
        res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
        res["proof"] = fe.last_proof
        return res

    return stream.run_thread_world(world, body)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_stream_quantile_is_np_quantile(world, engine):
    """The radix descent over per-rank histograms returns np.quantile bit for bit (duplicates, a
    bucket with thousands of equal values, negative values, q = 0 and 1)."""
    import torch
    from bpm_analysis_b200 import stream
    from bpm_analysis_b200.dist import shard_range
    rng = np.random.default_rng(3)
    series = [np.abs(rng.standard_normal(300001)) * 1e-3,
              np.concatenate([np.full(9000, 0.25), rng.standard_normal(5000)]),
              np.round(rng.standard_normal(50000), 2),
              np.array([3.0]), np.array([1.0, 2.0])]
    for x in series:
        for q in (0.0, 0.1, 0.5, 0.9371, 1.0):
            want = float(np.quantile(x, q))

            def body(comm):
                lo, hi = shard_range(len(x), comm.world, comm.rank)
                mine = engine.tensor(x[lo:hi]) if hi > lo else torch.zeros(0, dtype=torch.float64, device="cuda")
                val, status = stream.stream_quantiles(engine, comm, mine, len(x), [q, 0.5])
                return val.cpu().numpy(), status.cpu().numpy()

            for val, status in stream.run_thread_world(world, body):
                assert status[0] == 0 and val[0] == want, (len(x), q, val, want)
                assert status[1] == 0 and val[1] == float(np.quantile(x, 0.5))


@pytest.mark.parametrize("mode", ["parity", "fullrate"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stream_matches_oracle(world, mode, engine, ref_params):
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    params = dict(ref_params, filter_mode=mode)
    pcm, sr, _ = synth.config_c2(seed=11, duration_sec=420.0, sample_rate=48000)
    ref = ref_port.front_end(pcm, sr, params)
    for got in _run_sharded(world, engine, pcm, sr, params):
        assert got["sharded"], got["proof"]
        assert rel_err(got["envelope"], ref["envelope"]) < TOL
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
        assert rel_err(got["smoothed_dev"], np.asarray(ref["smoothed_dev_series"])) < TOL


@pytest.mark.parametrize("name", ["vulpine", "synth_c2_240s", "synth_holter_180s"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_analysis_bit_exact_on_reference_envelope(world, name, engine, ref_params):
    """Fed the reference's own envelope, every rank's sharded result equals the UNCHUNKED GPU result
    bit for bit -- whether the chunk proofs held (sharded) or the ranks fell back."""
    import torch
    from bpm_analysis_b200 import runtime, stream
    g = load_golden(name)
    env, rate = g["envelope"], int(g["rate"])
    one_floor, one_troughs = runtime.ops().noise_floor(env, rate, ref_params)
    took = []

    def body(comm):
        fe = stream.ShardedFrontEnd.for_envelope(len(env), rate, ref_params, comm, engine)
        out = fe.analyse(engine.tensor(env))
        torch.cuda.synchronize()
        return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}

    for got in stream.run_thread_world(world, body):
        took.append(got["sharded"])
        assert np.array_equal(got["troughs"], g["troughs"]) and np.array_equal(got["troughs"], one_troughs)
        assert np.array_equal(got["peaks"], g["raw_peaks"])
        assert np.array_equal(got["floor"], one_floor)
    assert len(set(took)) == 1                      # the ranks agree on the path they took


def test_sharded_falls_back_when_the_halo_is_too_short(engine, ref_params):
    """An analysis halo shorter than the rolling window cannot be proven: every rank must notice,
    and the replicated evaluation must still give the reference's lists."""
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    pcm, sr, _ = synth.config_c2(seed=5, duration_sec=300.0, sample_rate=48000)
    ref = ref_port.front_end(pcm, sr, ref_params)
    for got in _run_sharded(2, engine, pcm, sr, ref_params, analysis_halo=200):
        assert not got["sharded"]
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL


def test_sharded_long_holter_stream(engine, ref_params):
    """25 min of the C4 generator (bursts, dropouts) over 4 chunks."""
    from bpm_analysis_b200 import synth
    from oracle import ref_port
    pcm, sr, _ = synth.config_c4(seed=4, duration_sec=1500.0)
    ref = ref_port.front_end(pcm, sr, ref_params)
    for got in _run_sharded(4, engine, pcm, sr, ref_params):
        assert got["sharded"], got["proof"]
        assert np.array_equal(got["troughs"], ref["troughs"])
        assert np.array_equal(got["peaks"], ref["peaks"])
        assert rel_err(got["floor"], np.asarray(ref["floor"])) < TOL
        assert rel_err(got["strength"], np.asarray(ref["strength"])) < TOL
        assert rel_err(got["smoothed_dev"], np.asarray(ref["smoothed_dev_series"])) < TOL
