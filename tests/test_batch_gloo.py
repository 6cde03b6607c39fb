"""CPU tests of the multi-GPU host logic: clip sharding index
maths for world sizes 1..8 and the final gather over a real 2-process gloo group."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")


def test_shard_clips_partition():
    from emspec.batch import shard_clips
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 4096, 4097):
            spans = [shard_clips(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_clips(4096, 3, 8) == (1536, 2048)          # configs[3]: 512 clips per GPU
    with pytest.raises(ValueError):
        shard_clips(8, 8, 8)


def test_summary_pack_roundtrip():
    from emspec.batch import ClipSummary
    s = ClipSummary(clip=4095, frames=11235, energy=1234.5678, checksum=(1 << 61) + 12345)
    assert ClipSummary.unpack(s.pack()) == s


def test_image_checksum_order_dependent():
    from emspec.batch import image_checksum
    a = torch.arange(1000, dtype=torch.int64).remainder(256).to(torch.uint8)
    b = a.flip(0)
    assert image_checksum(a) == image_checksum(a.clone())
    assert image_checksum(a) != image_checksum(b)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    import torch.distributed as dist
    from emspec.batch import ClipSummary, gather_summaries, shard_clips
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_clips(n_clips, rank, world)
    local = [ClipSummary(c, 100 + c, 0.5 * c, (c << 40) | c) for c in range(lo, hi)]
    allc = gather_summaries(local, n_clips)
    q.put((rank, [(s.clip, s.frames, s.energy, s.checksum) for s in allc]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 8])
def test_gather_summaries_gloo_world2(n_clips):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [(c, 100 + c, 0.5 * c, (c << 40) | c) for c in range(n_clips)]
    assert got[0] == want and got[1] == want
