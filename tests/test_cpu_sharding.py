"""CPU tests of the N>1 host logic with torch.distributed/gloo, world_size 2 and 3:
 - batch sharding (frames independent, no collective)
 - the MCU-row partition protocol of mgpu.encode_partitioned (four allgathers + send/recv
   stitch) driven by an oracle-backed test double: the stitched stream must equal the oracle's
   whole-image JFIF bytes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, w, h, q, ratio, seed, kind, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import oracle_lib as O
    from band_double import OracleBandEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    mg = pkg.mgpu
    planes = O.synth_planes(w, h, seed, 0, kind)
    bands = mg.plan_bands(h, world)
    row0, bh = bands[rank]
    eng = OracleBandEngine(planes, w, h, q, ratio, row0, bh)
    header = torch.from_numpy(pkg.binding.jfif_header(w, h, q, ratio))
    segment = torch.zeros(w * bh * 3 + 4096, dtype=torch.uint8)
    out = torch.zeros(w * h * 3 + 4096, dtype=torch.uint8) if rank == 0 else None
    n = mg.encode_partitioned(eng, (0, 0, 0), w, bh, q, ratio, header, segment, out, dist, torch.device("cpu"))
    if rank == 0:
        np.save(result_path, out[:n].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h,q,ratio,kind", [(2, 64, 32, 75, 0, 0), (2, 96, 48, 90, 1, 1), (3, 80, 72, 97, 0, 1),
                                                     (2, 16, 16, 50, 1, 0)])
def test_partition_protocol_gloo(tmp_path, world, w, h, q, ratio, kind):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    result = str(tmp_path / "out.npy")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, w, h, q, ratio, 1234, kind, result), nprocs=world, join=True)
    got = np.load(result)
    r, g, b = O.synth_planes(w, h, 1234, 0, kind)
    want = O.encode_jfif(r, g, b, w, h, q, ratio)
    assert got.shape == want.shape and (got == want).all()


def test_plan_bands_and_frame_sharding():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    mg = ge.load_package().mgpu
    for h, world in [(32768, 8), (4320, 4), (72, 3), (16, 2)]:
        bands = mg.plan_bands(h, world)
        assert len(bands) == world and bands[0][0] == 0
        assert all(bh % 8 == 0 and bh > 0 for _, bh in bands)
        assert all(bands[k][0] + bands[k][1] == bands[k + 1][0] for k in range(world - 1))
        assert bands[-1][0] + bands[-1][1] == h
        assert max(bh for _, bh in bands) - min(bh for _, bh in bands) <= 8
    with pytest.raises(ValueError):
        mg.plan_bands(8, 2)
    for n, world in [(256, 8), (10, 4), (3, 8)]:
        seen = sorted(f for r in range(world) for f in mg.shard_frames(n, r, world))
        assert seen == list(range(n))


def _failing_worker(rank, world, port, mode, result_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import oracle_lib as O
    from band_double import OracleBandEngine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = ge.load_package()
    mg = pkg.mgpu
    w, h, q, ratio = 64, 48, 75, 0
    planes = O.synth_planes(w, h, 77, 0, 0)
    row0, bh = mg.plan_bands(h, world)[rank]
    eng = OracleBandEngine(planes, w, h, q, ratio, row0, bh)
    if mode == "engine" and rank == 1:
        def boom(pred):
            raise RuntimeError("engine failure on purpose")
        eng.band_bits = boom
    header = torch.from_numpy(pkg.binding.jfif_header(w, h, q, ratio))
    segment = torch.zeros(w * bh * 3 + 4096, dtype=torch.uint8)
    out = torch.zeros(700 if mode == "small" else w * h * 3 + 4096, dtype=torch.uint8) if rank == 0 else None
    what = "no error"
    try:
        mg.encode_partitioned(eng, (0, 0, 0), w, bh, q, ratio, header, segment, out, dist, torch.device("cpu"))
    except (RuntimeError, ValueError) as exc:
        what = type(exc).__name__
    open(os.path.join(result_dir, "r%d.txt" % rank), "w").write(what)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,expect", [("engine", "RuntimeError"), ("small", "ValueError")])
def test_partition_failure_reaches_every_rank(tmp_path, mode, expect):
    """A failing engine call on one rank, or an output buffer too small on rank 0, must raise on EVERY rank
    instead of leaving the others blocked in an allgather or a send (the status word of each exchange)."""
    import torch.multiprocessing as mp

    world = 3
    port = 29500 + (os.getpid() % 2000) + 17
    mp.spawn(_failing_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(str(tmp_path / ("r%d.txt" % r))).read() == expect
