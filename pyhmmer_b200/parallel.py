"""Multi-GPU search: shard the TARGET database across ranks, all-gather the hit records (once per wave of profiles).

Every (profile, sequence) comparison is independent, so the path shards with no data-path collective
(SURVEY 8(e)).  One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` for the CPU tests of
this host logic).  Rank r keeps the contiguous slice r of the database, balanced by residues -- the rule
of the reference's target-parallel dispatcher (src/pyhmmer/hmmer/_hmmsearch.py:153-171) -- runs the whole
fused cascade on its slice, and contributes its serialized hit records to ONE all-gather (sizes first, then
padded byte buffers).  Every rank then rebuilds identical `TopHits`: hits are admitted in global target
order with the global running Z, so the result equals the unsharded search exactly.
"""
import ctypes
import os

import numpy as np

from . import _lib
from .easel import DigitalSequenceBlock

__all__ = ["World", "shard_block", "search_sharded", "exchanged_waves", "gathered_waves", "all_gather_bytes", "pack_records", "unpack_records"]


class World:
    """Rank / size of this process in the job (1 process = 1 GPU)."""

    def __init__(self, rank=0, size=1, dist=None, device=None):
        self.rank, self.size, self.dist, self.device = rank, size, dist, device

    @classmethod
    def current(cls):
        try:
            import torch.distributed as dist
        except Exception:
            return cls()
        if dist.is_available() and dist.is_initialized():
            backend = dist.get_backend()
            device = None
            if backend == "nccl":
                import torch
                device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            return cls(dist.get_rank(), dist.get_world_size(), dist, device)
        return cls()


def shard_bounds(lengths, nshards):
    """Contiguous slices balanced by residue count -- the reference's rule, literally (_hmmsearch.py:153-171):
    chunksize = ceil(total / n); walk the targets accumulating residues; start a new chunk at the sequence
    that pushes the running size over chunksize; pad with empty chunks."""
    lengths = [int(v) for v in lengths]
    n = len(lengths)
    total = sum(lengths)
    chunksize = (total + nshards - 1) // nshards
    bounds = [0]
    current = 0
    for i, L in enumerate(lengths):
        current += L
        if current > chunksize and len(bounds) < nshards:
            bounds.append(i)
            current = 0
    while len(bounds) <= nshards:
        bounds.append(n)
    return bounds


def shard_block(block, world):
    """(offset, sub-block) of this rank."""
    b = shard_bounds([len(s) for s in block], world.size)
    lo, hi = b[world.rank], b[world.rank + 1]
    return lo, DigitalSequenceBlock(block.alphabet, list.__getitem__(block, slice(lo, hi)))


def _raw_bytes(recs, rectype):
    """The records as one byte string: straight from the array they came in when the list still mirrors it."""
    raw = getattr(recs, "raw", None)
    if raw is not None and len(raw) == len(recs):
        return bytes(raw)
    n = len(recs)
    return bytes((rectype * max(n, 1))(*recs))[: n * ctypes.sizeof(rectype)]


def pack_records(hits, doms, text, counters, seq_offset, profile_offset=0, profiles=None):
    """Serialize one rank's results: header | HitRec[] | DomainRec[] | text | counters (int64) [| profile indices (int32)]."""
    nh, nd = len(hits), len(doms)
    hb = bytearray(_raw_bytes(hits, _lib.HitRec))
    if nh and seq_offset:
        np.frombuffer(hb, dtype=np.dtype(_lib.HitRec))["seq"] += seq_offset       # local -> global target index
    if nh and profile_offset:
        np.frombuffer(hb, dtype=np.dtype(_lib.HitRec))["profile"] += profile_offset   # local -> global profile index (scan)
    cnt = np.ascontiguousarray(counters, dtype=np.int64)
    prof = np.ascontiguousarray([] if profiles is None else profiles, dtype=np.int32)
    header = np.array([nh, nd, len(text), cnt.size, prof.size], dtype=np.int64).tobytes()
    return b"".join([header, bytes(hb), _raw_bytes(doms, _lib.DomainRec), bytes(text), cnt.tobytes(), prof.tobytes()])


def unpack_records(buf, with_profiles=False):
    nh, nd, nt, nc, npf = (int(v) for v in np.frombuffer(buf[:40], dtype=np.int64))
    o = 40
    hs, ds = ctypes.sizeof(_lib.HitRec), ctypes.sizeof(_lib.DomainRec)
    hits = list((_lib.HitRec * nh).from_buffer_copy(buf[o:o + nh * hs])) if nh else []
    o += nh * hs
    doms = list((_lib.DomainRec * nd).from_buffer_copy(buf[o:o + nd * ds])) if nd else []
    o += nd * ds
    text = bytes(buf[o:o + nt])
    o += nt
    counters = np.frombuffer(buf[o:o + 8 * nc], dtype=np.int64).copy()
    o += 8 * nc
    if with_profiles:
        return hits, doms, text, counters, np.frombuffer(buf[o:o + 4 * npf], dtype=np.int32).tolist()
    return hits, doms, text, counters


class _Exchange:
    """Persistent buffers of the one exchange of the path: page-locked host staging and device slots that grow to the
    largest payload seen, so that a step moves exactly what it has to -- no zero-filled pageable slot, no power-of-two
    padding, no blocking pageable copies."""

    def __init__(self):
        self.key, self.cap = None, 0
        self.h_in = self.h_out = self.d_in = self.d_out = self.h_sz = self.d_sz = self.d_szs = self.h_szs = None
        self.stream = None
        self.slot = 1 << 16                                  # slot size every rank uses: grows only from what ALL ranks saw

    def side_stream(self, torch, dev):
        """The exchange has a (high-priority) stream of its own: the engine's lanes hold the queued cascades of the waves
        still being searched, and a collective issued behind them would wait for all of them."""
        if dev.type != "cuda":
            return None
        if self.stream is None or self.stream.device != dev:
            self.stream = torch.cuda.Stream(device=dev, priority=-1)
        return self.stream

    def ensure(self, torch, dev, world, need):
        key = (str(dev), world.size)
        if key != self.key or need > self.cap:
            cap = max(1 << 16, 1 << (int(need) - 1).bit_length())
            self.slot = max(self.slot, 1 << 16) if key == self.key else 1 << 16       # (a new process group starts over)
            pin = dev.type == "cuda"
            self.h_in = torch.empty(cap, dtype=torch.uint8, pin_memory=pin)
            self.h_out = torch.empty(world.size * cap, dtype=torch.uint8, pin_memory=pin)
            self.d_in = torch.empty(cap, dtype=torch.uint8, device=dev)
            self.d_out = torch.empty(world.size * cap, dtype=torch.uint8, device=dev)
            self.h_sz = torch.zeros(1, dtype=torch.int64, pin_memory=pin)
            self.h_szs = torch.zeros(world.size, dtype=torch.int64, pin_memory=pin)
            self.d_sz = torch.zeros(1, dtype=torch.int64, device=dev)
            self.d_szs = torch.zeros(world.size, dtype=torch.int64, device=dev)
            self.key, self.cap = key, cap


_EXCHANGE = _Exchange()


def all_gather_bytes(data, world):
    """The exchange of the path: a variable-length all-gather of byte strings (NCCL over NVLink, or gloo on CPU).

    ONE collective in the common case: every rank contributes a slot of the agreed size -- 8 bytes of payload length, then
    the payload -- through persistent page-locked / device buffers, and one device->host copy brings the gathered slots
    back.  When some payload does not fit, every rank sees that in the gathered lengths, the agreed slot size doubles up to
    the largest payload, and the gather is repeated once (the size then stays for the rest of the process)."""
    if world.size == 1:
        return [data]
    import torch
    dev = world.device or torch.device("cpu")
    side = _EXCHANGE.side_stream(torch, dev)
    if side is None:
        return _all_gather_bytes(data, world, torch, dev, None)
    with torch.cuda.stream(side):
        return _all_gather_bytes(data, world, torch, dev, side)


def _all_gather_bytes(data, world, torch, dev, side):
    dist = world.dist
    n = len(data)
    ex = _EXCHANGE
    while True:
        slot = ex.slot
        ex.ensure(torch, dev, world, slot)
        h = ex.h_in.numpy()
        h[:8] = np.frombuffer(np.int64(n).tobytes(), dtype=np.uint8)
        fits = n + 8 <= slot
        if n and fits:
            h[8:8 + n] = np.frombuffer(data, dtype=np.uint8)
        used = ((8 + (n if fits else 0)) + 255) & ~255       # bytes of the slot that carry anything: all that goes up
        d_in, d_out, h_out = ex.d_in[:slot], ex.d_out[:world.size * slot], ex.h_out[:world.size * slot]
        d_in[:used].copy_(ex.h_in[:used], non_blocking=True)
        try:
            dist.all_gather_into_tensor(d_out, d_in)
        except (RuntimeError, AttributeError, NotImplementedError):       # a backend without the flat form: list form
            outs = [torch.empty(slot, dtype=torch.uint8, device=dev) for _ in range(world.size)]
            dist.all_gather(outs, d_in)
            d_out.copy_(torch.cat(outs))
        h_out.copy_(d_out, non_blocking=True)
        if side is not None:
            side.synchronize()
        host = h_out.numpy().reshape(world.size, slot)
        sizes = [int(np.frombuffer(host[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(world.size)]
        if max(sizes) + 8 <= slot:
            return [host[r, 8:8 + sizes[r]].tobytes() for r in range(world.size)]
        ex.slot = max(1 << 16, 1 << (max(sizes) + 8 - 1).bit_length())     # the same on every rank: they all saw the same lengths


def merge_rank_records(parts):
    """Concatenate per-rank (hits, doms, text, counters); hits come out ordered by (profile, global seq)."""
    hits, doms, text = [], [], bytearray()
    counters = None
    for h, d, t, c in parts:
        dbase, tbase = len(doms), len(text)
        for r in h:
            r.dom_offset += dbase
        for r in d:
            r.text_offset += tbase
        hits.extend(h)
        doms.extend(d)
        text.extend(t)
        counters = c.copy() if counters is None else counters + c
    hits.sort(key=lambda r: (r.profile, r.seq))
    return hits, doms, bytes(text), counters


def agree_max(value, world):
    """The largest ``value`` of any rank (one 8-byte all-reduce on the exchange stream)."""
    if world.size == 1:
        return int(value)
    import torch
    dev = world.device or torch.device("cpu")
    side = _EXCHANGE.side_stream(torch, dev)
    t = torch.tensor([int(value)], dtype=torch.int64)
    if side is None:
        world.dist.all_reduce(t, op=world.dist.ReduceOp.MAX)
        return int(t.item())
    with torch.cuda.stream(side):
        d = t.to(dev, non_blocking=True)
        world.dist.all_reduce(d, op=world.dist.ReduceOp.MAX)
        out = d.cpu()
    return int(out.item())


def gathered_waves(nlocal, gen, world, lo, P):
    """The rounds of a wave-by-wave exchange: for every round yields ``(local wave tuple, [every rank's packed records])``.
    The first payload carries this rank's number of waves, so the ranks know the number of rounds (the largest of them)
    after the first gather -- no collective of its own; a rank that has run out of waves contributes empty payloads."""
    empty = ([], [], [], b"", np.zeros((P, 4), np.int64))
    rounds, r = nlocal, 0
    while r < rounds:
        local = next(gen, empty)
        profs, hits, doms, text, counters = local
        mine = np.int64(nlocal).tobytes() + pack_records(hits, doms, text, counters, lo, profiles=profs)
        got = all_gather_bytes(mine, world)
        if r == 0:
            rounds = max(int(np.frombuffer(b[:8], dtype=np.int64)[0]) for b in got)
        yield local, [b[8:] for b in got]
        r += 1
    for _ in gen:                                              # (drains the engine's job: nothing is left by construction)
        pass


def exchanged_waves(pipeline, oms, sub, lo, world):
    """This rank's search of its shard ``sub`` (first target = global index ``lo``), wave by wave, with one all-gather of hit
    records per wave: a generator of ``(complete, hits, doms, text, counters)`` -- the records of EVERY rank for the profiles
    in ``complete`` (those every rank has finished), hits ordered by (profile, global target), ``counters`` [P][4] summed over
    the ranks so far.  While a wave's records are exchanged and assembled, the GPUs search the following waves; the ranks
    agree on the number of rounds first (their wave plans may differ when their shards do), a rank that runs out of waves
    contributes empty payloads, and records of profiles some rank has not finished yet wait for the round that completes them."""
    P = len(oms)
    if len(sub) and P >= 16:
        nlocal, gen = pipeline._run_waves(oms, sub)
    elif len(sub):
        nlocal, gen = 1, iter([(list(range(P)),) + tuple(pipeline._run(oms, sub))])
    else:
        nlocal, gen = 1, iter([(list(range(P)), [], [], b"", np.zeros((P, 4), np.int64))])
    done = np.zeros(P, np.int32)
    total = np.zeros((P, 4), np.int64)
    carry = None
    for local, gathered in gathered_waves(nlocal, gen, world, lo, P):
        parts = [unpack_records(b, with_profiles=True) for b in gathered]
        for part in parts:
            done[part[4]] += 1
        hits, doms, text, counters = merge_rank_records(([carry] if carry else []) + [p[:4] for p in parts])
        total += counters.reshape(P, 4)
        ready = done == world.size
        complete = np.nonzero(ready)[0].tolist()
        done[ready] = -(1 << 20)                               # reported once
        rest = [h for h in hits if not ready[h.profile]] if len(complete) < P else []
        carry = (rest, doms, text, np.zeros(P * 4, np.int64)) if rest else None
        yield complete, ([h for h in hits if ready[h.profile]] if rest else hits), doms, text, total


def search_sharded(pipeline, queries, block, local, world):
    """One query batch against the sharded database; every rank returns the same list of `TopHits`."""
    lo, sub = local
    L = len(block[0]) if len(block) else pipeline.L_HINT
    oms = pipeline._optimized_many(queries, L)
    results = [None] * len(oms)
    first = True
    for complete, hits, doms, text, counters in exchanged_waves(pipeline, oms, sub, lo, world):
        for qi, th in zip(complete, pipeline._assemble(queries, oms, block, hits, doms, text, counters, only=complete, count_targets=first)):
            results[qi] = th
        first = False
    return results
