"""Long-stream chunking (bpm_analysis_b200/stream.py) with the oracle as the engine:
planner geometry, thread world (2, 3 ranks) and a real world-size-2 gloo run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from bpm_analysis_b200 import stream as bstream
from bpm_analysis_b200 import synth
from bpm_analysis_b200.params import default_params, effective_decimation


def _params():
    p = default_params()
    p["save_filtered_wav"] = False
    return p


def _oracle_plan(sample_rate, params):
    """FilterPlan stand-in for the CPU tests (parity mode geometry; no CUDA needed)."""
    from bpm_analysis_b200.runtime import plan_filter
    return plan_filter(sample_rate, params)


def _reference(pcm, sr, params):
    from oracle import ref_port
    return ref_port.front_end(pcm, sr, params)


def _chunked(comm, pcm, sr, params):
    from _oracle_engine import OracleEngine
    eng = OracleEngine(sr, params)
    fe = bstream.ChunkedFrontEnd(len(pcm), sr, params, comm, eng, plan=_oracle_plan(sr, params))
    f0, f1 = fe.frames()
    out = fe.run(torch.from_numpy(pcm[f0:f1].copy()))
    return {k: v.numpy() for k, v in out.items()}


def _check(got, ref):
    assert np.array_equal(got["troughs"], ref["troughs"])
    assert np.array_equal(got["peaks"], ref["peaks"])
    scale = np.max(np.abs(ref["envelope"]))
    assert np.max(np.abs(got["envelope"] - ref["envelope"])) <= 1e-9 * scale
    fl = ref["floor"].values if hasattr(ref["floor"], "values") else ref["floor"]
    assert np.max(np.abs(got["floor"] - fl)) <= 1e-9 * np.max(np.abs(fl))


def test_floor_item_range_geometry():
    knots = np.array([10, 500, 2000, 4000, 6000, 9000])
    # window 1000 -> off 499, left 500
    a0, a1 = bstream.floor_item_range(knots, 3000, 5000, 1000, 10000)
    assert a0 == 2000 and a1 == 6001          # last knot <= 2499, first knot >= 5499
    a0, a1 = bstream.floor_item_range(knots, 0, 400, 1000, 10000)
    assert a0 == 0 and a1 == 2001
    a0, a1 = bstream.floor_item_range(knots, 8800, 10000, 1000, 10000)
    assert a0 == 6000 and a1 == 10000
    assert bstream.floor_item_range(np.array([], dtype=np.int64), 5, 9, 11, 100) == (0, 100)


def test_chunk_plan_covers_stream():
    plan = bstream.ChunkPlan(n_frames=1000003, m=6850, frames_per_sample=146, halo=700, world=3)
    cores = [plan.core(r) for r in range(3)]
    assert cores[0][0] == 0 and cores[-1][1] == 6850
    assert all(cores[i][1] == cores[i + 1][0] for i in range(2))
    for r in range(3):
        f0, f1 = plan.frames(r)
        e0, e1 = plan.ext(r)
        assert f0 == e0 * 146 and f1 <= 1000003
        assert -(-(f1 - f0) // 146) == e1 - e0          # the slice decimates to exactly the extended chunk


@pytest.mark.parametrize("world", [2, 3])
def test_thread_world_matches_unchunked_oracle(world):
    params = _params()
    pcm, sr, _ = synth.config_c2(seed=11, duration_sec=420.0, sample_rate=48000)
    ref = _reference(pcm, sr, params)
    res = bstream.run_thread_world(world, lambda comm: _chunked(comm, pcm, sr, params))
    for got in res:
        _check(got, ref)


def test_exact_floor_on_a_given_envelope():
    """Same envelope in, chunked floor / troughs / peaks must be bit-identical."""
    from oracle import ref_port
    from _oracle_engine import OracleEngine
    params = _params()
    pcm, sr, _ = synth.config_c4(seed=4, duration_sec=1500.0)       # bursts and dropouts
    env, rate, _ = ref_port.preprocess_pcm(pcm, sr, params)
    floor, troughs = ref_port.calculate_dynamic_noise_floor(env, rate, params)
    peaks = ref_port.find_raw_peaks(env, rate, params, floor.values)

    def body(comm):
        fe = bstream.ChunkedFrontEnd.for_envelope(len(env), rate, params, comm, OracleEngine(sr, params))
        return {k: v.numpy() for k, v in fe.analyse(torch.from_numpy(env.copy())).items()}

    for got in bstream.run_thread_world(4, body):
        assert np.array_equal(got["troughs"], troughs)
        assert np.array_equal(got["peaks"], peaks)
        assert np.array_equal(got["floor"], floor.values)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _params()
        pcm, sr, _ = synth.config_c1(seed=5, duration_sec=240.0)
        got = _chunked(bstream.DistComm(), pcm, sr, params)
        q.put((rank, {k: got[k] for k in ("troughs", "peaks", "envelope", "floor")}))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_matches_unchunked_oracle():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params = _params()
    pcm, sr, _ = synth.config_c1(seed=5, duration_sec=240.0)
    ref = _reference(pcm, sr, params)
    for r in range(world):
        _check(got[r], ref)


# ----------------------------------------------------------------------------- every stage sharded
def _sharded(comm, pcm, sr, params, **kw):
    from _oracle_engine import OracleChunkEngine
    eng = OracleChunkEngine(sr, params)
    fe = bstream.ShardedFrontEnd(len(pcm), sr, params, comm, eng, plan=_oracle_plan(sr, params), **kw)
    f0, f1 = fe.frames()
    out = fe.run(torch.from_numpy(pcm[f0:f1].copy()), gather_series=True)
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


@pytest.mark.parametrize("world", [1, 3])
def test_stream_quantiles_descent_is_np_quantile(world):
    """The descent / collect / finish protocol (numpy restatement of the bpm_key_* operators)."""
    from _oracle_engine import OracleChunkEngine
    from bpm_analysis_b200.dist import shard_range
    eng = OracleChunkEngine(8000, _params())
    rng = np.random.default_rng(1)
    for x in (np.abs(rng.standard_normal(40001)) * 1e-3, np.round(rng.standard_normal(5000), 1), np.array([2.0, 1.0])):
        for q in (0.0, 0.1, 0.77, 1.0):
            def body(comm):
                lo, hi = shard_range(len(x), comm.world, comm.rank)
                v, st = bstream.stream_quantiles(eng, comm, torch.from_numpy(x[lo:hi].copy()), len(x), [q, 0.5])
                return v.numpy(), st.numpy()
            for v, st in bstream.run_thread_world(world, body):
                assert st[0] == 0 and v[0] == float(np.quantile(x, q))
                assert st[1] == 0 and v[1] == float(np.quantile(x, 0.5))


@pytest.mark.parametrize("world", [3])
def test_sharded_thread_world_matches_unchunked_oracle(world):
    params = _params()
    pcm, sr, _ = synth.config_c1(seed=3, duration_sec=240.0)
    ref = _reference(pcm, sr, params)
    for got in bstream.run_thread_world(world, lambda comm: _sharded(comm, pcm, sr, params)):
        assert got["sharded"]
        _check(got, ref)
        assert np.allclose(got["smoothed_dev"], np.asarray(ref["smoothed_dev_series"]), rtol=1e-9, atol=0)


def test_sharded_short_halo_falls_back():
    params = _params()
    pcm, sr, _ = synth.config_c1(seed=3, duration_sec=120.0)
    ref = _reference(pcm, sr, params)
    for got in bstream.run_thread_world(2, lambda comm: _sharded(comm, pcm, sr, params, analysis_halo=150)):
        assert not got["sharded"]
        _check(got, ref)


def _gloo_sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = _params()
        pcm, sr, _ = synth.config_c1(seed=5, duration_sec=200.0)
        got = _sharded(bstream.DistComm(), pcm, sr, params)
        q.put((rank, {k: got[k] for k in ("troughs", "peaks", "envelope", "floor", "sharded")}))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_matches_unchunked_oracle():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params = _params()
    pcm, sr, _ = synth.config_c1(seed=5, duration_sec=200.0)
    ref = _reference(pcm, sr, params)
    for r in range(world):
        assert got[r]["sharded"]
        _check(got[r], ref)
