"""Sequence parallelism for the denoising hot path: one process per GPU, tokens sharded across ranks,
weights replicated, ONE all-gather of packed K|V per attention.

  video tokens     contiguous (f h w) order, L / P rows per rank (L % P == 0 at the BASELINE sizes: 32760 = 8 * 4095)
  geometry tokens  frame-aligned shards (21 frames over 8 ranks -> 3,3,3,3,3,2,2,2) so the per-frame attention of the
                   VGGT frame blocks needs no communication; global attention and the adapter only need
                   "local queries x all keys", so the two streams may be partitioned differently
  everything else  (LayerNorm, modulation, GEMMs, RoPE, gates, text/CLIP cross-attention against the replicated context,
                   camera AdaLN with token-aligned features) is token-local

CFG parallelism on top (SURVEY §8e "alternative"): a denoise step is TWO independent joint_forwards
(conditional and unconditional context, fusion/model_wan21.py:295-317).  With an even number of ranks the two forwards run
concurrently on the two halves of the node, each half sequence-parallel over world/2 ranks, and the halves swap their
64-channel predictions once per step (CFGParallel.exchange: one 8 MB all-gather inside a 2-rank pair group).  Compared
with sequence parallelism over all ranks this halves the K|V bytes every rank receives per attention, doubles the rows per
shard (better tile-wave occupancy) and halves the number of collectives on the critical path.

The reference has no live multi-GPU path (its Ulysses hooks import a module that is not in the tree);
this is new functionality.  Collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def split_even(n: int, parts: int) -> List[int]:
    """Sizes of `parts` contiguous shards of n items, the first (n % parts) shards one larger."""
    base, rem = divmod(n, parts)
    return [base + (1 if r < rem else 0) for r in range(parts)]


def offsets(sizes: List[int]) -> List[int]:
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += s
    return out


@dataclass
class SPLayout:
    """Static description of how one (f, h, w) token grid is sharded over `world` ranks."""
    world: int
    f: int
    h: int
    w: int
    n_special: int = 5
    video_rows: List[int] = field(default_factory=list)     # rows per rank (video stream)
    frames: List[int] = field(default_factory=list)         # frames per rank (geometry stream)

    def __post_init__(self):
        L = self.f * self.h * self.w
        self.video_rows = split_even(L, self.world)
        self.frames = split_even(self.f, self.world)
        self.P = self.n_special + self.h * self.w           # tokens per frame (geometry)

    @property
    def L(self):
        return self.f * self.h * self.w

    @property
    def N(self):
        return self.f * self.P

    def video_range(self, rank) -> Tuple[int, int]:
        o = offsets(self.video_rows)[rank]
        return o, o + self.video_rows[rank]

    def frame_range(self, rank) -> Tuple[int, int]:
        o = offsets(self.frames)[rank]
        return o, o + self.frames[rank]

    def geo_rows(self) -> List[int]:
        return [fr * self.P for fr in self.frames]

    def geo_range(self, rank) -> Tuple[int, int]:
        f0, f1 = self.frame_range(rank)
        return f0 * self.P, f1 * self.P


class SPContext:
    """Per-process handle: rank / world, the process group, and the gather primitives used by the engine."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.layout: Optional[SPLayout] = None
        self.n_gathers = 0
        self.gather_bytes = 0
        # pipelined K|V exchange: the local K|V rows are gathered in `kv_chunks` row slices on a side stream, and the attention
        # over slice c (split-KV partial + merge) overlaps the gather of slice c+1.  1 = single blocking gather.
        self.kv_chunks = 4 if self.world >= 4 else (2 if self.world > 1 else 1)
        self._comm_stream = None

    @property
    def comm_stream(self):
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(priority=-1)
        return self._comm_stream

    def gather_chunks_async(self, x: torch.Tensor, n_chunks: int):
        """Start all-gathers of `n_chunks` contiguous row slices of x (equal [rows, C] block on every rank) on the side
        stream.  Returns [(buffer [world * rows_c, C], event)], in slice order; wait on the event before reading the buffer."""
        rows = x.shape[0]
        sizes = split_even(rows, n_chunks)
        offs = offsets(sizes)
        main = torch.cuda.current_stream()
        bufs = [torch.empty((self.world * n, *x.shape[1:]), device=x.device, dtype=x.dtype) for n in sizes]
        ready = torch.cuda.Event()
        ready.record(main)
        out = []
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            x.record_stream(self.comm_stream)               # x and the buffers are consumed on the side stream: tell the
            for c, n in enumerate(sizes):                    # caching allocator, or a later main-stream allocation could reuse them
                bufs[c].record_stream(self.comm_stream)
                dist.all_gather_into_tensor(bufs[c], x[offs[c]: offs[c] + n], group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                out.append((bufs[c], ev))
                self.n_gathers += 1
                self.gather_bytes += bufs[c].numel() * bufs[c].element_size()
        return out

    def all_gather_rows_async(self, x: torch.Tensor, sizes: List[int]):
        """all_gather_rows on the side stream.  Returns a function that waits for the exchange and yields [sum(sizes), C]."""
        assert x.is_contiguous() and x.shape[0] == sizes[self.rank]
        m = max(sizes)
        C = x.shape[1:]
        if x.shape[0] < m:
            xp = torch.zeros((m, *C), device=x.device, dtype=x.dtype)
            xp[: x.shape[0]] = x
        else:
            xp = x
        buf = torch.empty((self.world * m, *C), device=x.device, dtype=x.dtype)
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            xp.record_stream(self.comm_stream)               # incl. the padded temporary of a short shard (freed on return otherwise)
            buf.record_stream(self.comm_stream)
            dist.all_gather_into_tensor(buf, xp, group=self.group)
            ev.record(self.comm_stream)
        self.n_gathers += 1
        self.gather_bytes += buf.numel() * buf.element_size()

        def finish():
            torch.cuda.current_stream().wait_event(ev)
            if len(set(sizes)) == 1:
                return buf
            b3 = buf.view(self.world, m, *C)
            return torch.cat([b3[r, : sizes[r]] for r in range(self.world)], dim=0)

        return finish

    def set_grid(self, f, h, w):
        if self.layout is None or (self.layout.f, self.layout.h, self.layout.w) != (f, h, w):
            self.layout = SPLayout(self.world, f, h, w)
        return self.layout

    # ---- collectives ---------------------------------------------------------------------------------------------------
    def all_gather_rows(self, x: torch.Tensor, sizes: List[int]) -> torch.Tensor:
        """Concatenate the ranks' row blocks: x is this rank's [sizes[rank], C] block (contiguous); returns [sum(sizes), C].
        Equal sizes: one all_gather_into_tensor.  Unequal (frame-aligned geometry shards): pad to the largest block, one
        all_gather_into_tensor, then compact."""
        assert x.is_contiguous() and x.shape[0] == sizes[self.rank]
        C = x.shape[1:]
        self.n_gathers += 1
        if len(set(sizes)) == 1:
            out = torch.empty((sum(sizes), *C), device=x.device, dtype=x.dtype)
            dist.all_gather_into_tensor(out, x, group=self.group)
            self.gather_bytes += out.numel() * out.element_size()
            return out
        m = max(sizes)
        if x.shape[0] < m:
            xp = torch.empty((m, *C), device=x.device, dtype=x.dtype)
            xp[: x.shape[0]] = x
            xp[x.shape[0]:] = 0
        else:
            xp = x
        buf = torch.empty((self.world * m, *C), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(buf, xp, group=self.group)
        self.gather_bytes += buf.numel() * buf.element_size()
        buf = buf.view(self.world, m, *C)
        return torch.cat([buf[r, : sizes[r]] for r in range(self.world)], dim=0)


class CFGParallel:
    """Classifier-free-guidance parallelism: ranks [0, world/2) evaluate the conditional forward (role 0), ranks
    [world/2, world) the unconditional one (role 1); rank r and rank r + world/2 form a pair that swaps predictions.

    Built collectively (every rank of the default group must construct it at the same point: dist.new_group is collective).
      .role   0 = conditional (context_pos), 1 = unconditional (context_neg)
      .sp     SPContext over this half (None when the half is a single rank)
      .exchange(pred) -> (pred_pos, pred_neg), both on every rank
    """

    def __init__(self):
        world, rank = dist.get_world_size(), dist.get_rank()
        if world < 2 or world % 2:
            raise ValueError(f"CFG parallelism needs an even number of ranks, got {world}")
        half = world // 2
        self.world, self.rank, self.half = world, rank, half
        self.role = 0 if rank < half else 1
        halves = [dist.new_group(ranks=list(range(0, half))), dist.new_group(ranks=list(range(half, world)))]
        pairs = [dist.new_group(ranks=[r, r + half]) for r in range(half)]
        self.pair_group = pairs[rank % half]
        self.sp = SPContext(halves[self.role]) if half > 1 else None
        self.n_exchanges = 0
        self.exchange_bytes = 0

    def exchange(self, pred: torch.Tensor):
        """This rank's prediction (full tensor, identical on all ranks of its half) -> (conditional, unconditional)."""
        pred = pred.contiguous()
        both = torch.empty((2, *pred.shape), device=pred.device, dtype=pred.dtype)
        # pair group rank order = (r, r + half) = (pos, neg); flat views: gloo insists on [world * n] <- [n]
        dist.all_gather_into_tensor(both.view(-1), pred.view(-1), group=self.pair_group)
        self.n_exchanges += 1
        self.exchange_bytes += both.numel() * both.element_size()
        return both[0], both[1]
