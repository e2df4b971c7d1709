"""CPU: the N>1 host logic (partition + gather) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from masr_b200 import shard


def fake_engine(waves):
    """Deterministic stand-in for ConformerEngine.transcribe on a CPU box."""
    toks, scores = [], []
    for w in waves:
        n = len(w) // 4000
        toks.append([int(abs(w[i * 17 % len(w)]) * 1000) % 4000 + 1 for i in range(n)])
        scores.append(float(np.float32(w.mean()) * 100))
    return toks, scores


def test_partition_balances_and_covers():
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 480001, 512).tolist()
    for world in (1, 2, 4, 8):
        sh = shard.partition(lens, world)
        assert sorted(i for s in sh for i in s) == list(range(512))
        sizes = [len(s) for s in sh]
        assert max(sizes) - min(sizes) <= 1
        tot = [sum(lens[i] for i in s) for s in sh]
        assert (max(tot) - min(tot)) / np.mean(tot) < 0.02
    assert shard.partition([5, 5, 5, 5], 2) == [[0, 3], [1, 2]]
    assert shard.partition([], 4) == [[], [], [], []]
    assert shard.partition([7], 2) == [[0], []]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    waves = [rng.standard_normal(int(n)).astype(np.float32) for n in rng.integers(8000, 64000, 11)]
    toks, scores = shard.sharded_transcribe(waves, fake_engine, max_tokens=20)
    ref_t, ref_s = fake_engine(waves)
    ok = toks == ref_t and np.allclose(scores, ref_s)
    # ragged worlds: fewer utterances than ranks
    t2, s2 = shard.sharded_transcribe(waves[:1], fake_engine, max_tokens=20)
    ok = ok and t2 == ref_t[:1]
    ret[rank] = ok
    dist.destroy_process_group()


def test_sharded_transcribe_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_single_process_passthrough():
    waves = [np.ones(9000, np.float32), np.ones(5000, np.float32) * 0.5]
    t, s = shard.sharded_transcribe(waves, fake_engine, max_tokens=8)
    assert (t, s) == fake_engine(waves)
