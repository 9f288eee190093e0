"""Multi-GPU host logic on CPU (gloo, world_size 2 and 3): the MDX chunk list is partitioned across ranks with no overlap /
no gap, every rank owns ONE contiguous sample span of the stem, the all-gather of those spans equals the single-rank
result; RVC segments converted round-robin and all-gathered reassemble the utterance."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aicovergen_b200.mdx import allgather_spans, chunk_descriptors, shard_range, shard_spans


def fake_process(n, n_fft, chunk, rank, world):
    """Stand-in for STFT -> net -> iSTFT: each kept output sample gets the value f(song index); exactly the
    addressing of b200vc_mdx_ola_store."""
    src, lo, hi, dst, klo, khi = chunk_descriptors(n, n_fft, chunk, 2)
    gen = chunk - n_fft
    out = np.zeros(n, dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int64)
    a, b = shard_range(len(src), rank, world)
    for i in range(a, b):
        k = np.arange(gen)
        d = dst[i] + k
        ok = (d >= klo[i]) & (d < khi[i]) & (d >= 0) & (d < n)
        out[d[ok]] += np.sin(d[ok] * 0.001) + 2.0
        cnt[d[ok]] += 1
    return out, cnt


def test_descriptors_cover_every_sample_exactly_once():
    for n in (44100 * 7 + 13, 10584000, 253440 * 4, 100000):
        for (n_fft, dim_t) in ((7680, 256), (5120, 256), (6144, 512)):
            chunk = 1024 * (dim_t - 1)
            out, cnt = fake_process(n, n_fft, chunk, 0, 1)
            assert (cnt == 1).all(), (n, n_fft, cnt.min(), cnt.max())
    # reference geometry: 4-min song, Kim_Vocal_2 class -> 22 + 22 chunk inferences per sweep
    assert len(chunk_descriptors(10584000, 7680, 261120, 2)[0]) == 44


def test_shard_range_partitions():
    for n in (1, 7, 44, 88):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, cnt = fake_process(n, 7680, 261120, rank, world)
    spans = shard_spans(n, 7680, 261120, 2, world)
    lo, hi = spans[rank]
    assert (cnt[lo:hi] == 1).all() and cnt.sum() == hi - lo, "a rank writes exactly its own contiguous span"
    t = torch.from_numpy(np.stack([out, -out]).astype(np.float32))          # [2, n] like a stereo stem
    allgather_spans(t, spans, dist.group.WORLD)
    if rank == world - 1:
        q.put(t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sum_equals_single_rank(world):
    n = 44100 * 40 + 321
    ref, cnt = fake_process(n, 7680, 261120, 0, 1)
    assert (cnt == 1).all()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], ref.astype(np.float32)) and np.array_equal(got[1], -ref.astype(np.float32))


def _seg_worker(rank, world, port, seg_samples, q):
    import types

    from aicovergen_b200.vc_infer_pipeline import VC
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vc = VC(40000, types.SimpleNamespace(device="cpu", is_half=True, x_pad=1, x_query=1, x_center=2, x_max=3))
    vc.group = dist.group.WORLD
    net_g = types.SimpleNamespace(upp=400)
    lens = [vc._segment_frames(n) * 400 - 2 * vc.t_pad_tgt for n in seg_samples]
    outs = [(torch.arange(L, dtype=torch.float32) + 1000.0 * i) if i % world == rank else None for i, L in enumerate(lens)]
    vc._gather_segments(outs, seg_samples, net_g, world, rank)
    if rank == world - 1:
        q.put([o.numpy() for o in outs])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rvc_segments_round_robin_all_gather(world):
    """4 segments of different lengths (a 4-min song, vc_infer_pipeline.py:516-603) over 2 / 3 ranks: every rank ends up
    with every segment; lengths follow p_len = min(n // 160, 2 * T_hubert)."""
    from aicovergen_b200.hubert import conv_out_len
    seg_samples = [16000 * 4 + 160 * 3, 16000 * 3 + 160 * 11, 16000 * 5, 16000 * 2 + 160 * 7 + 80]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_seg_worker, args=(r, world, port, seg_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (n, o) in enumerate(zip(seg_samples, got)):
        L = min(n // 160, 2 * conv_out_len(n)[-1]) * 400 - 2 * 40000
        assert o.shape == (L,) and o[0] == 1000.0 * i and o[-1] == 1000.0 * i + L - 1


def test_plan_cache_is_lru_and_holds_more_than_a_songs_segments():
    """A 4-minute song cuts into 4 segments of different lengths (vc_infer_pipeline.py:516-545): the per-shape plan
    cache must keep all of them across songs, and evict least-recently-used beyond its capacity."""
    from aicovergen_b200.plans import PLAN_CACHE, PlanCache

    assert PLAN_CACHE >= 5
    built = []
    cache = PlanCache()

    def get(k):
        return cache.get_or_build(k, lambda: built.append(k) or object())

    for _ in range(3):                       # three "songs" with the same four segment lengths
        for k in (6600, 6412, 7031, 5590):
            get(k)
    assert built == [6600, 6412, 7031, 5590]
    for k in range(PLAN_CACHE):              # fill past capacity: the oldest entries go first
        get(100 + k)
    assert 6600 not in cache and len(cache) == PLAN_CACHE
