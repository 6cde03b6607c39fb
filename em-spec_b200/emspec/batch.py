"""Batch mode host logic: clips are independent, so they are sharded by
contiguous blocks across ranks with no data-path collective; the only exchange is the
final gather of per-clip summaries (NCCL on GPUs, gloo in the CPU tests).

One process per GPU (torch.distributed); nothing here touches audio samples."""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple


def shard_clips(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r of W owns clips [r*C//W, (r+1)*C//W): contiguous, disjoint, covering, and
    balanced to within one clip."""
    if not (0 <= rank < world) or n_clips < 0:
        raise ValueError("bad shard request")
    return (rank * n_clips) // world, ((rank + 1) * n_clips) // world


@dataclasses.dataclass
class ClipSummary:
    clip: int          # global clip index
    frames: int        # columns produced
    energy: float      # sum of the accumulated grid (energy kept after the drop rule)
    checksum: int      # 62-bit checksum of the colour-index image (order-dependent)

    def pack(self) -> List[float]:
        # float64 carries 53 bits exactly: split the checksum in two 31-bit halves
        return [float(self.clip), float(self.frames), float(self.energy),
                float(self.checksum >> 31), float(self.checksum & 0x7FFFFFFF)]

    @staticmethod
    def unpack(v: Sequence[float]) -> "ClipSummary":
        return ClipSummary(int(v[0]), int(v[1]), float(v[2]), (int(v[3]) << 31) | int(v[4]))


def image_checksum(index_u8) -> int:
    """Order-dependent 62-bit checksum of a u8 image tensor (computed where the tensor lives)."""
    import torch
    flat = index_u8.reshape(-1).to(torch.int64)
    pos = torch.arange(flat.numel(), device=flat.device, dtype=torch.int64)
    mod = (1 << 31) - 1
    a = int(((flat * ((pos % 65521) + 1)) % mod).sum().item()) % mod
    b = int(flat.sum().item()) % mod
    return (a << 31) | b


def gather_summaries(local: Sequence[ClipSummary], n_clips: int, device=None) -> List[ClipSummary]:
    """All ranks receive every clip's summary, ordered by clip index.  One all_gather of a
    fixed-size padded block per rank — the batch job's only collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sorted(local, key=lambda s: s.clip)
    world = dist.get_world_size()
    cap = (n_clips + world - 1) // world
    buf = torch.full((cap, 5), -1.0, dtype=torch.float64, device=device)
    for i, s in enumerate(local):
        buf[i] = torch.tensor(s.pack(), dtype=torch.float64)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = []
    for t in out:
        for row in t.cpu().tolist():
            if row[0] >= 0:
                res.append(ClipSummary.unpack(row))
    res.sort(key=lambda s: s.clip)
    return res
