"""CPU test of the N>1 host logic: world_size-2 over gloo.  Each rank renders its sample subset (the oracle stands
in for the GPU here — tests may use it), one reduce(sum), and rank 0 must hold exactly the 1-rank sum."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp, out_path):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pt, orc = ge.load_package(), ge.load_oracle()
    D = __import__("importlib").import_module("pt_b200.distributed")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = pt.Scene.build(3, width=20, spp=spp, seed=1)
    ora = orc.OracleScene(scene.desc, pt)
    begin, count, stride = D.partition_samples(spp, rank, world)
    mean, _ = ora.render(scene.camera, count, seed=4, sample_begin=begin, sample_stride=stride, nan_policy=pt.PT_NAN_DROP, threads=1)
    accum = torch.from_numpy(mean * count)  # per-rank SUM of radiance
    D.reduce_accumulators(accum, 0)
    if rank == 0:
        full, _ = ora.render(scene.camera, spp, seed=4, nan_policy=pt.PT_NAN_DROP, threads=1)
        np.save(out_path, np.stack([accum.numpy() / spp, full]))
    dist.destroy_process_group()


def test_partition_covers_every_sample_once(pt):
    D = __import__("importlib").import_module("pt_b200.distributed")
    for spp in (1, 2, 7, 100, 4000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                b, c, s = D.partition_samples(spp, r, world)
                seen += [b + k * s for k in range(c)]
            assert sorted(seen) == list(range(spp))


def test_progressive_halves_cover_every_sample_once(pt):
    """progressive.py: batches x halves (A even / B odd) x ranks enumerate each sample index exactly once."""
    P = __import__("importlib").import_module("pt_b200.progressive")
    for world in (1, 2, 8):
        seen, done = [], 0
        for batch in (3, 1, 4):                      # samples per half per rank in successive batches
            for r in range(world):
                for b, c, s in P.half_ranges(done, batch, r, world):
                    seen += [b + k * s for k in range(c)]
            done += batch
        assert sorted(seen) == list(range(2 * 8 * world))
        halves = [P.half_ranges(0, 8, r, world) for r in range(world)]
        assert all((b + k * s) // world % 2 == h for rr in halves for h, (b, c, s) in enumerate(rr) for k in range(c))


def test_two_rank_reduce_equals_single_rank(pt, orc, tmp_path):
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(2, _free_port(), 5, out), nprocs=2, join=True)  # odd spp: ragged split 3 + 2
    got, full = np.load(out)
    assert np.allclose(got, full, rtol=1e-12, atol=1e-12)
