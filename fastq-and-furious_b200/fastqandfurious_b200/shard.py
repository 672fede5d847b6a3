"""Byte-range sharding of one FASTQ stream over the GPUs of a box.

One process per GPU (torch.distributed; NCCL over NVLink on the box, gloo in the CPU tests).  Rank g
holds the bytes [c_g, c_g + own_len) of the stream.  Records straddle the cuts, so every rank receives
a HALO -- the first bytes of its right neighbour's shard -- with ONE neighbour exchange (send/recv ring
shift), parses own + halo in place, and owns the records whose leading newline lies in its own
range.  The only other communication is an all-gather of one int64 per rank (lines per shard): the
"rank mod 4" rule of the fast path needs the global line number of each shard's first line.
The reference has no counterpart (it is single threaded); the contract that a buffer must hold the
largest entry is the reference's own (src/fastqandfurious.py:219-223) and sizes the halo.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib, device

DEFAULT_HALO = 1 << 20
SLOT_RING = 16  # epochs of count slots kept per rank (fused transport): >= the number of shards


class ShardPlan:
    """Pure bookkeeping: who owns which bytes, who exchanges what with whom."""

    def __init__(self, rank, world, own_lens, halo_bytes=DEFAULT_HALO):
        assert len(own_lens) == world and 0 <= rank < world
        self.rank, self.world = rank, world
        self.own_lens = [int(x) for x in own_lens]
        self.offsets = [sum(self.own_lens[:g]) for g in range(world)]
        self.total = sum(self.own_lens)
        self.halo_bytes = int(halo_bytes)

    @property
    def own_len(self):
        return self.own_lens[self.rank]

    @property
    def offset(self):
        return self.offsets[self.rank]

    @property
    def is_last(self):
        return self.rank == self.world - 1

    def halo_len(self, g=None):
        """Bytes rank g receives from rank g + 1 (0 for the last rank)."""
        g = self.rank if g is None else g
        if g == self.world - 1:
            return 0
        return min(self.halo_bytes, self.own_lens[g + 1])

    def send_len(self, g=None):
        """Bytes rank g sends to rank g - 1 (0 for rank 0)."""
        g = self.rank if g is None else g
        return 0 if g == 0 else self.halo_len(g - 1)

    def check(self):
        for g in range(self.world - 1):
            if self.own_lens[g + 1] < self.halo_bytes and g + 1 != self.world - 1:
                raise ValueError('shard %d is shorter than the halo: a halo must come from one neighbour' % (g + 1))


def exchange_halo(buf, plan, group=None):
    """Ring shift: rank g sends its first send_len bytes to g-1 and receives its halo from g+1 into
    buf[own_len : own_len + halo_len].  `buf` is a 1-D uint8 tensor (CUDA with NCCL, CPU with gloo)."""
    ops = []
    own = plan.own_len
    if plan.send_len():
        ops.append(dist.P2POp(dist.isend, buf[:plan.send_len()], _peer(plan.rank - 1, group), group))
    if plan.halo_len():
        ops.append(dist.P2POp(dist.irecv, buf[own:own + plan.halo_len()], _peer(plan.rank + 1, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return plan.halo_len()


def _peer(rank_in_group, group):
    return dist.get_global_rank(group, rank_in_group) if group is not None else rank_in_group


def line_bases(own_lines, plan, group=None):
    """own_lines: int64 tensor [1] (lines this shard owns).  Returns (base, total): int64 tensors [1] with
    the lines owned by all earlier shards and by all shards -- one all-gather of 8 bytes per rank, the
    prefix is computed on the device (no host synchronisation)."""
    gathered = torch.empty(plan.world, dtype=torch.int64, device=own_lines.device)
    dist.all_gather_into_tensor(gathered, own_lines, group=group)
    incl = torch.cumsum(gathered, 0)
    base = (incl[plan.rank:plan.rank + 1] - gathered[plan.rank:plan.rank + 1]).contiguous()
    return base, incl[-1:].contiguous()


class ShardedParser:
    """Per-rank driver of one sharded parse.

    transport 'fused' (default on NVLink boxes): the shard buffers and the count slots live in SYMMETRIC
    memory (torch.distributed._symmetric_memory).  The halo is pulled straight out of the right neighbour's
    buffer with a peer copy over NVLink after one device-side barrier; the kernel that counts a shard's
    lines STORES {count, epoch} into the memory of every later shard over NVLink, and the emit kernel of
    those shards waits for the slots itself -- no collective call, no barrier and no extra kernel between
    scan and emit (fqb_shard_scan_publish / fqb_shard_emit_wait).
    transport 'peer': same memory, but the counts are read through peer-mapped pointers after a second
    device-side barrier.  transport 'nccl': one send/recv ring shift + one 8-byte all-gather."""

    def __init__(self, plan, dev, group=None, cfg=0, transport=None, double_buffer=False):
        """double_buffer (fused transport): two shard buffers that take turns (a streaming caller fills one while the
        other is parsed); the halo of parse k + 1 is then brought in on a second stream while parse k runs -- one warp
        waits for the neighbour's signal, a copy engine moves the bytes -- instead of by a kernel in front of every
        scan."""
        import os
        self.plan, self.dev, self.group, self.cfg = plan, torch.device(dev), group, cfg
        self.double = False
        self._want_double = bool(double_buffer)
        self.flags = _lib.FLAG_CFG(cfg)
        if os.environ.get('FQB_SHARD_TAIL', '0')[:1] == '1':  # count / publish / signal in the scan's epilogue
            self.flags |= _lib.FLAG_SHARD_TAIL
        self.transport = transport or os.environ.get('FQB_SHARD_TRANSPORT', 'fused')
        self.epoch = 0
        self._signalled = 0  # epochs announced to the left neighbour
        self.gepoch = 0      # hand-over epoch of the sharded general path
        self.gws = None
        if plan.world == 1:
            self.transport = 'none'
        n = plan.own_len + plan.halo_len()
        self.n = n
        n_alloc = max(plan.own_lens[g] + plan.halo_len(g) for g in range(plan.world))
        with torch.cuda.device(self.dev):
            self.base = torch.zeros(1, dtype=torch.int64, device=self.dev)
            self.result = torch.zeros(16, dtype=torch.int64, device=self.dev)
            need = _lib.lib().fqb_workspace_bytes(n, 0, self.flags)
            self.ws = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
            if self.transport in ('peer', 'fused'):
                try:
                    self._init_peer(n_alloc)
                except Exception as e:  # API drift, no P2P access, ...: the NCCL transport always works
                    import sys
                    print('fastqandfurious_b200.shard: symmetric-memory transport unavailable (%r); using NCCL' % (e,),
                          file=sys.stderr)
                    self.transport = 'nccl'
            if self.transport not in ('peer', 'fused'):
                self.full = torch.empty(max(n_alloc, 1), dtype=torch.uint8, device=self.dev)
                self.own_lines = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.buf = self.full[:n]
        self.bufs = [self.buf, self.full2[:n]] if self.double else [self.buf]

    def _init_peer(self, n_alloc):
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        self.full = symm_mem.empty(max(n_alloc, 16), dtype=torch.uint8, device=self.dev)
        self.h_buf = symm_mem.rendezvous(self.full, group)
        self.counts = symm_mem.empty(16, dtype=torch.int64, device=self.dev)
        self.counts.zero_()
        self.h_cnt = symm_mem.rendezvous(self.counts, group)
        self.own_lines = self.counts[:1]
        plan = self.plan
        if plan.halo_len():
            self.right = self.h_buf.get_buffer(plan.rank + 1, (self.full.numel(),), torch.uint8)
        ptrs = [int(self.h_cnt.buffer_ptrs[r]) for r in range(plan.rank)]
        self.n_left = len(ptrs)
        self.left_ptrs = (ctypes.c_void_p * max(1, len(ptrs)))(*ptrs)
        # fused exchange: slots[epoch % SLOT_RING][source rank] = {count, epoch} in every rank's memory.  With the
        # ready signal sent as early as possible the first shard can run up to world - 1 parses ahead of the last
        # one (its pull for parse k only needs the scan of parse k - j of the shard j places to its right), so the
        # ring must hold at least `world` epochs.
        world = plan.world
        assert world <= SLOT_RING
        self.slots = symm_mem.empty(SLOT_RING * world * 2, dtype=torch.int64, device=self.dev)
        self.slots.zero_()
        self.h_slots = symm_mem.rendezvous(self.slots, group)
        self.pub_ptrs = []
        for ring in range(SLOT_RING):
            dst = [int(self.h_slots.buffer_ptrs[r]) + ((ring * world + plan.rank) * 2) * 8
                   for r in range(plan.rank + 1, world)]
            self.pub_ptrs.append((ctypes.c_void_p * max(1, len(dst)))(*dst))
        self.n_pub = world - 1 - plan.rank
        # halo handshake: ready[0] of every rank is written by its right neighbour (epoch); status word behind it
        self.ready = symm_mem.empty(4, dtype=torch.int64, device=self.dev)
        self.ready.zero_()
        self.h_ready = symm_mem.rendezvous(self.ready, group)
        self.ready_left = int(self.h_ready.buffer_ptrs[plan.rank - 1]) if plan.rank > 0 else None
        self.right_ptr = int(self.h_buf.buffer_ptrs[plan.rank + 1]) if plan.halo_len() else None
        self.halo_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # sharded general path: hand-over slot {resume, records, ended, epoch}, written by the left neighbour
        self.gslot = symm_mem.empty(8, dtype=torch.int64, device=self.dev)
        self.gslot.zero_()
        self.h_gslot = symm_mem.rendezvous(self.gslot, group)
        self.gslot_right = int(self.h_gslot.buffer_ptrs[plan.rank + 1]) if plan.rank + 1 < world else None
        self.gslot_left = int(self.h_gslot.buffer_ptrs[plan.rank - 1]) if plan.rank > 0 else None
        if self._want_double and self.transport == 'fused' and world > 1:
            self.full2 = symm_mem.empty(max(n_alloc, 16), dtype=torch.uint8, device=self.dev)
            self.h_buf2 = symm_mem.rendezvous(self.full2, group)
            self.rights = [self.right if plan.halo_len() else None,
                           self.h_buf2.get_buffer(plan.rank + 1, (self.full2.numel(),), torch.uint8) if plan.halo_len() else None]
            self.pull_stream = torch.cuda.Stream(device=self.dev)
            self.pull_done = [torch.cuda.Event(), torch.cuda.Event()]
            self.parse_done = [None, None]
            self.double = True
        torch.cuda.synchronize(self.dev)
        self.h_buf.barrier(channel=0)
        torch.cuda.synchronize(self.dev)

    def own(self, which=0):
        """The tensor view the caller fills with this rank's bytes (double_buffer: parse k reads buffer k % 2, the
        first parse is k = 1)."""
        return self.bufs[which][:self.plan.own_len]

    def _enqueue_pull(self, k):
        """double_buffer: the halo of parse k, on the pull stream -- after the parse that last used that buffer, one
        warp waits for the right neighbour's "bytes of parse k in place", then a peer copy (copy engine)."""
        plan, b = self.plan, k % 2
        own, h = plan.own_len, plan.halo_len()
        cur = torch.cuda.current_stream(self.dev)
        if self.parse_done[b] is not None:
            self.pull_stream.wait_event(self.parse_done[b])
        else:
            self.pull_stream.wait_stream(cur)  # the caller's fill of the buffer
        with torch.cuda.stream(self.pull_stream):
            if h:
                _lib.check(_lib.lib().fqb_shard_wait_ready(self.ready.data_ptr(), k, self.halo_status.data_ptr(),
                                                           ctypes.c_void_p(self.pull_stream.cuda_stream)),
                           'fqb_shard_wait_ready')
                device.launch_count += 1
                self.bufs[b][own:own + h].copy_(self.rights[b][:h], non_blocking=True)
            self.pull_done[b].record(self.pull_stream)

    def _step_double(self, table, next_ready, qual, qual_add):
        plan, L = self.plan, _lib.lib()
        n, own = self.n, plan.own_len
        cur = torch.cuda.current_stream(self.dev)
        k = self.epoch + 1
        if not next_ready:
            # streaming caller: the neighbour's "bytes of parse k in place" follows ITS refill for parse k, so the halo
            # of parse k is requested with parse k (a pull enqueued ahead would wait for a refill that may never come)
            self._enqueue_pull(k)
        elif self.epoch == 0:
            self.signal_ready()     # static buffers: our bytes of parse 1 are in place
            self._enqueue_pull(1)
        b = k % 2
        buf = self.bufs[b]
        cur.wait_event(self.pull_done[b])
        stream = ctypes.c_void_p(cur.cuda_stream)
        sentinel = 1 if plan.rank == 0 else 0
        self.epoch = k
        parity = k % SLOT_RING
        sig = next_ready and self.ready_left is not None
        _lib.check(L.fqb_shard_scan_publish_ready(
            buf.data_ptr() if n else None, n, own, sentinel, self.own_lines.data_ptr(), self.pub_ptrs[parity], self.n_pub, k,
            self.ready_left if sig else None, self._signalled + 1 if sig else 0,
            qual.data_ptr() if (qual is not None and n) else None, int(qual_add), self.ws.data_ptr(), self.ws.numel(),
            self.flags, stream), 'fqb_shard_scan_publish_ready')
        if sig:
            self._signalled += 1
        one_kernel = (self.cfg & 15) == 0 and n > 0 and bool(self.flags & _lib.FLAG_SHARD_TAIL)
        device.launch_count += 0 if one_kernel else 1
        wait = self.slots.data_ptr() + parity * plan.world * 2 * 8
        _lib.check(L.fqb_shard_emit_wait(buf.data_ptr() if n else None, n, own, sentinel, 1 if plan.is_last else 0,
                                         plan.offset - sentinel, wait, plan.rank, k, table.data_ptr(), table.shape[0],
                                         self.result.data_ptr(), self.ws.data_ptr(), self.ws.numel(), self.flags, stream),
                   'fqb_shard_emit_wait')
        device.launch_count += 2
        ev = torch.cuda.Event()
        ev.record(cur)
        self.parse_done[b] = ev
        self.buf = buf  # the buffer of the last parse (rows refer to it)
        if next_ready:
            self._enqueue_pull(k + 1)  # static buffers: the neighbour announces parse k + 1 with its scan of parse k

    def signal_ready(self):
        """Fused transport: tell the left neighbour that this shard's bytes for the next parse are in place (it pulls
        its halo from them).  step(next_ready=True) does it right after the scan; a streaming caller that refills
        the shard buffer calls it once the new bytes have landed (and passes next_ready=False)."""
        if self.transport != 'fused' or self.ready_left is None:
            return
        with torch.cuda.device(self.dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().fqb_shard_signal_ready(self.ready_left, self._signalled + 1, stream),
                       'fqb_shard_signal_ready')
        self._signalled += 1
        device.launch_count += 1

    def alloc_qual(self):
        """An int8 Phred mirror for step(qual=...): one byte per byte of own range + halo, congruent to the shard
        buffer modulo 16 (the scan kernel stores it with the 16-byte vectors it loads)."""
        raw = torch.empty(self.n + 32, dtype=torch.int8, device=self.dev)
        shift = (self.buf.data_ptr() - raw.data_ptr()) % 16
        return raw[shift:shift + self.n]

    def step(self, table, exchange=True, next_ready=True, qual=None, qual_add=-33):
        """One asynchronous parse of the shard.  `table`: int64 [cap,6] CUDA tensor.  `qual` (alloc_qual()): also
        decode the quality strings, qual[pos4 - plan.offset : pos5 - plan.offset] of every emitted row is the
        record's quality string + qual_add (arrayadd_b, src/_fastqandfurious.c:161-185)."""
        plan, L = self.plan, _lib.lib()
        n = self.n
        own = plan.own_len
        self._check_table(table)
        if plan.world > 1 and self.transport == 'fused' and not exchange:
            # the halo handshake is what bounds how far a shard may run ahead of its neighbours (at most world - 1
            # parses, which the ring of count slots is sized for): without it a fast rank could lap a slow one
            raise ValueError("transport 'fused': every step exchanges the halo (exchange=False is not supported)")
        if qual is not None:
            if qual.dtype != torch.int8 or qual.device != self.buf.device or qual.numel() < n or not qual.is_contiguous():
                raise ValueError('qual must be a contiguous int8 tensor of at least %d elements on %s' % (n, self.buf.device))
            if n and (qual.data_ptr() - self.buf.data_ptr()) % 16:
                raise ValueError('qual must be congruent to the shard buffer modulo 16 (use alloc_qual())')
        if self.double:
            with torch.cuda.device(self.dev):
                return self._step_double(table, next_ready, qual, qual_add)
        with torch.cuda.device(self.dev):
            if plan.world > 1 and exchange and self.transport == 'fused':
                # wait for the right neighbour's "bytes in place", pull its head over NVLink: one kernel
                stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                if self.epoch == 0:
                    self.signal_ready()  # the first parse: our bytes are in place now
                _lib.check(L.fqb_shard_pull_halo(self.buf.data_ptr() + own, self.right_ptr, plan.halo_len(),
                                                 self.ready.data_ptr(), None, self.epoch + 1,
                                                 self.halo_status.data_ptr(), stream), 'fqb_shard_pull_halo')
                device.launch_count += 1
            elif plan.world > 1 and exchange:
                if self.transport == 'peer':
                    self.h_buf.barrier(channel=0)  # every rank's bytes are in place
                    if plan.halo_len():
                        self.buf[own:own + plan.halo_len()].copy_(self.right[:plan.halo_len()], non_blocking=True)
                else:
                    exchange_halo(self.buf, plan, self.group)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            sentinel = 1 if plan.rank == 0 else 0
            if plan.world > 1 and self.transport == 'fused':
                self.epoch += 1
                parity = self.epoch % SLOT_RING
                # scan + count + publication + (next_ready: the bytes of the NEXT parse are in place already -- static
                # or refilled buffer) the ready signal to the left neighbour: one call (one kernel with FQB_SHARD_TAIL=1 in the environment)
                sig = next_ready and self.ready_left is not None
                _lib.check(L.fqb_shard_scan_publish_ready(
                    self.buf.data_ptr() if n else None, n, own, sentinel, self.own_lines.data_ptr(), self.pub_ptrs[parity],
                    self.n_pub, self.epoch, self.ready_left if sig else None, self._signalled + 1 if sig else 0,
                    qual.data_ptr() if (qual is not None and n) else None, int(qual_add), self.ws.data_ptr(),
                    self.ws.numel(), self.flags, stream), 'fqb_shard_scan_publish_ready')
                if sig:
                    self._signalled += 1
                one_kernel = (self.cfg & 15) == 0 and n > 0 and bool(self.flags & _lib.FLAG_SHARD_TAIL)
                device.launch_count += 0 if one_kernel else 1  # the kernel that counts, publishes and signals
                wait = self.slots.data_ptr() + parity * plan.world * 2 * 8
                _lib.check(L.fqb_shard_emit_wait(self.buf.data_ptr() if n else None, n, own, sentinel,
                                                 1 if plan.is_last else 0, plan.offset - sentinel, wait, plan.rank,
                                                 self.epoch, table.data_ptr(), table.shape[0], self.result.data_ptr(),
                                                 self.ws.data_ptr(), self.ws.numel(), self.flags, stream),
                           'fqb_shard_emit_wait')
                device.launch_count += 2  # scan, emit
                return
            if qual is not None and n:
                _lib.check(L.fqb_shard_scan_decode(self.buf.data_ptr(), n, own, sentinel, self.own_lines.data_ptr(), None, 0,
                                                   0, qual.data_ptr(), int(qual_add), self.ws.data_ptr(), self.ws.numel(),
                                                   self.flags, stream), 'fqb_shard_scan_decode')
            else:
                _lib.check(L.fqb_shard_scan(self.buf.data_ptr() if n else None, n, own, sentinel,
                                            self.own_lines.data_ptr(), self.ws.data_ptr(), self.ws.numel(), self.flags,
                                            stream), 'fqb_shard_scan')
            if plan.world > 1:
                if self.transport == 'peer':
                    self.h_cnt.barrier(channel=1)  # every rank has published its line count
                    _lib.check(L.fqb_sum_u64_ptrs(self.left_ptrs, self.n_left, self.base.data_ptr(), stream),
                               'fqb_sum_u64_ptrs')
                else:
                    self.base, _ = line_bases(self.own_lines, plan, self.group)
                    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.fqb_shard_emit(self.buf.data_ptr() if n else None, n, own, sentinel,
                                        1 if plan.is_last else 0, plan.offset - sentinel, self.base.data_ptr(),
                                        table.data_ptr(), table.shape[0], self.result.data_ptr(), self.ws.data_ptr(),
                                        self.ws.numel(), self.flags, stream), 'fqb_shard_emit')
        device.launch_count += 3 + (1 if self.transport == 'peer' else 0)

    def _check_table(self, table):
        if (not isinstance(table, torch.Tensor) or table.dtype != torch.int64 or table.dim() != 2 or table.shape[1] != 6
                or table.device != self.buf.device or not table.is_contiguous()):
            raise ValueError('table must be a contiguous int64 [cap,6] CUDA tensor on %s' % (self.buf.device,))

    def step_general(self, table, max_lines=None):
        """The same shard through the GENERAL path (multi-line records, damaged entries): call it on every rank
        when any rank's step() reported that its input needs it (`needs_general()`).  The halo of the last
        step() is reused.  Needs the peer-memory transports (the hand-over travels through symmetric memory)."""
        if self.plan.world > 1 and self.transport not in ('fused', 'peer'):
            raise NotImplementedError('the sharded general path needs the peer-memory transports')
        plan, L = self.plan, _lib.lib()
        n, own = self.n, plan.own_len
        self._check_table(table)
        if max_lines is None:
            max_lines = n // 16 + 1024
        self.gepoch += 1
        with torch.cuda.device(self.dev):
            need = L.fqb_workspace_bytes(n, max_lines, self.flags)
            if self.gws is None or self.gws.numel() < need:
                self.gws = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            sentinel = 1 if plan.rank == 0 else 0
            if plan.world > 1:
                local, right, left = self.gslot.data_ptr(), self.gslot_right, self.gslot_left
            else:
                if getattr(self, '_gslot1', None) is None:
                    self._gslot1 = torch.zeros(8, dtype=torch.int64, device=self.dev)
                local, right, left = self._gslot1.data_ptr(), None, None
            _lib.check(L.fqb_shard_general(self.buf.data_ptr() if n else None, n, own, sentinel, 1 if plan.rank == 0 else 0,
                                           1 if plan.is_last else 0, plan.offset - sentinel, local, right, left, self.gepoch,
                                           self.gepoch - 1, table.data_ptr(), table.shape[0], self.result.data_ptr(),
                                           self.gws.data_ptr(), self.gws.numel(), int(max_lines), self.flags, stream),
                       'fqb_shard_general')
        device.launch_count += 15

    def needs_general(self):
        """True on every rank if the last step() of ANY rank found input the 4-line fast path cannot represent
        (one all-reduce of a flag; synchronises)."""
        res = device.read_result(self.result)
        flag = torch.tensor([1 if res.error == _lib.ERR_SHARD_GENERAL else 0], dtype=torch.int32, device=self.dev)
        if self.plan.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(flag.item())

    def read(self):
        """FqbResult of the last step (synchronises)."""
        res = device.read_result(self.result)
        if self.transport == 'fused' and self.plan.world > 1 and int(self.halo_status.item()):
            raise RuntimeError('the right neighbour did not signal its bytes within 10 s (fused halo pull)')
        if res.error == _lib.ERR_HALO:
            raise ValueError('a record runs past the %d-byte halo (or the last shard is shorter than a record): '
                             'raise halo_bytes' % self.plan.halo_bytes)
        if res.error == _lib.ERR_PEER:
            raise RuntimeError('an earlier shard did not publish its line count within 10 s (fused exchange)')
        if res.error == _lib.ERR_SHARD_GENERAL:
            raise NotImplementedError('this input needs the general path (multi-line records or damaged entries): '
                                      'check needs_general() on every rank after step() and run step_general()')
        if res.error:
            raise RuntimeError('fqb_shard_emit: error %d' % res.error)
        return res


class ShardedHostParser:
    """End to end on N GPUs: ONE logical stream whose byte range [c_g, c_g + own_len) sits in rank g's HOST memory ->
    the rows of the records rank g owns (absolute stream offsets, entryfunc_abspos semantics) back in host memory,
    plus the global index of its first record: the ranks' row blocks, in rank order, are the table of the stream.

    Double-buffered shards (ShardedParser(double_buffer=True)): while parse k runs, the bytes of parse k + 1 travel
    host -> device into the other buffer on a copy stream (the "bytes in place" signal to the left neighbour follows
    the copy on that stream), the halo of parse k + 1 is pulled from the right neighbour on the pull stream, and the
    rows of parse k - 1 travel device -> host on a third stream.  A buffer is refilled only after the parse that read it
    has finished; by then the left neighbour has pulled its halo from it (its scan of that parse, which follows its
    pull, has published the count our emit waited for)."""

    def __init__(self, plan, dev, cfg=0, cap=None):
        self.plan, self.dev = plan, torch.device(dev)
        self.parser = ShardedParser(plan, dev, cfg=cfg, double_buffer=True)
        if not self.parser.double:
            raise RuntimeError('ShardedHostParser needs the fused peer-memory transport (world > 1, symmetric memory)')
        cap = cap or (plan.own_len + plan.halo_len()) // 96 + 64
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream()
            self.d2h_stream = torch.cuda.Stream()
            self.res_stream = torch.cuda.Stream()
            self.dtable = [torch.empty((cap, 6), dtype=torch.int64, device=self.dev) for _ in range(2)]
            self.dresult = [torch.zeros(16, dtype=torch.int64, device=self.dev) for _ in range(2)]
            self.htable = [torch.empty((cap, 6), dtype=torch.int64).pin_memory() for _ in range(2)]
            self.hresult = [torch.zeros(16, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.res_ready = [None, None]   # events: result header of the parse in slot b is on the host
        self.rows_ready = [None, None]
        self.pending = []               # parses submitted and not yet returned: (k, slot)
        self.stats = {'h2d_bytes': 0, 'd2h_bytes': 0}

    def submit(self, host_own):
        """Enqueue parse k of this rank's bytes (1-D uint8 CPU tensor of plan.own_len bytes, pinned for full PCIe
        speed).  Asynchronous; collect() returns the results in submission order."""
        P, plan = self.parser, self.plan
        if not isinstance(host_own, torch.Tensor) or host_own.is_cuda or host_own.dtype != torch.uint8 or \
                host_own.numel() != plan.own_len:
            raise TypeError('host_own must be a 1-D uint8 CPU tensor of %d bytes' % plan.own_len)
        if len(self.pending) >= 2:
            raise RuntimeError('two parses are in flight: collect() one first')
        k = P.epoch + 1
        b = k % 2
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream()
            if P.parse_done[b] is not None:
                self.copy_stream.wait_event(P.parse_done[b])  # the parse that last read this buffer (and the left neighbour's pull)
            else:
                self.copy_stream.wait_stream(cur)
            with torch.cuda.stream(self.copy_stream):
                P.own(b).copy_(host_own, non_blocking=True)
                P.signal_ready()  # "my bytes of parse k are in place", behind the copy
                h2d = torch.cuda.Event()
                h2d.record(self.copy_stream)
            cur.wait_event(h2d)
            # the rows of this slot's previous parse must have left the device table
            if self.rows_ready[b] is not None:
                cur.wait_event(self.rows_ready[b])
            P.result = self.dresult[b]
            P.step(self.dtable[b], next_ready=False)
            self.res_stream.wait_event(P.parse_done[b])
            with torch.cuda.stream(self.res_stream):  # its own stream: the rows of the parse before travel meanwhile
                self.hresult[b].copy_(self.dresult[b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.res_stream)
            self.res_ready[b] = ev
        self.pending.append((k, b))
        self.stats['h2d_bytes'] += plan.own_len

    def collect(self):
        """(k0, rows) of the oldest parse in flight: global index of this rank's first record, int64 [n, 6] ndarray of
        absolute offsets (a view of pinned memory, valid until the slot is reused two submits later)."""
        k, b = self.pending.pop(0)
        self.res_ready[b].synchronize()
        res = _lib.FqbResult.from_buffer_copy(self.hresult[b].numpy().tobytes())
        if res.error:
            self.parser.result = self.dresult[b]
            self.parser.read()  # raises with the matching message
        n = int(res.n_records)
        with torch.cuda.device(self.dev):
            self.d2h_stream.wait_event(self.res_ready[b])
            with torch.cuda.stream(self.d2h_stream):
                if n:
                    self.htable[b][:n].copy_(self.dtable[b][:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.d2h_stream)
            self.rows_ready[b] = ev
            ev.synchronize()
        self.stats['d2h_bytes'] += n * 48 + 128
        return int(res.reserved[0]), self.htable[b][:n].numpy(), res

    def parse(self, host_own):
        self.submit(host_own)
        return self.collect()


class ShardedJob:
    """bench.py's multi-GPU step: one synthetic shard per rank of a single logical stream."""

    def __init__(self, parser, table, rec_bytes):
        self.parser, self.table, self.rec_bytes = parser, table, rec_bytes
        self.buf = parser.own()

    @classmethod
    def synthetic(cls, shard_bytes, rec_bytes, rank, world, dev, cfg=0, halo_bytes=DEFAULT_HALO, double_buffer=None):
        plan = ShardPlan(rank, world, [shard_bytes] * world, halo_bytes)
        plan.check()
        if double_buffer is None:
            double_buffer = os.environ.get('FQB_SHARD_DOUBLE', '1')[:1] == '1'
        parser = ShardedParser(plan, dev, cfg=cfg, double_buffer=double_buffer)
        with torch.cuda.device(dev):
            src = device.synth_fixed(0, device=dev, first_byte=plan.offset, n_bytes=plan.own_len)
            for which in range(len(parser.bufs)):
                parser.own(which).copy_(src)
            del src
            table = torch.empty((shard_bytes // rec_bytes + 64, 6), dtype=torch.int64, device=dev)
        return cls(parser, table, rec_bytes)

    def step(self, table=None, result=None, flags=None):
        self.parser.step(self.table)

    def result(self):
        return self.parser.read()

    def records_per_step(self):
        return self.result().n_records

    def verify_fixed(self, header_len=32, read_len=150):
        """Closed-form check of this rank's rows for the fixed-geometry synthetic stream (outside any timed region):
        row i of the table is global record k0 + i, whose '@' sits at byte (k0 + i) * record_bytes; the shards'
        record ranges must tile the stream (checked with one all-gather of (k0, n))."""
        res = self.result()
        n, k0 = int(res.n_records), int(res.reserved[0])
        rec = self.rec_bytes
        rows = self.table[:n]
        k = torch.arange(k0, k0 + n, dtype=torch.int64, device=rows.device)
        p0 = k * rec
        want = torch.stack([p0, p0 + header_len, p0 + header_len + 1, p0 + header_len + 1 + read_len,
                            p0 + header_len + read_len + 4, p0 + header_len + 2 * read_len + 4], dim=1)
        if not torch.equal(rows, want):
            bad = int((rows != want).any(dim=1).nonzero()[0])
            raise AssertionError('rank %d: row %d of the sharded table is %s, expected %s' %
                                 (self.parser.plan.rank, bad, rows[bad].tolist(), want[bad].tolist()))
        mine = torch.tensor([k0, n], dtype=torch.int64, device=rows.device)
        every = torch.empty(2 * self.parser.plan.world, dtype=torch.int64, device=rows.device)
        dist.all_gather_into_tensor(every, mine)
        every = every.view(-1, 2).tolist()
        nxt = 0
        for g, (a, c) in enumerate(every):
            if a != nxt:
                raise AssertionError('shard %d starts at record %d, expected %d' % (g, a, nxt))
            nxt = a + c
        return nxt

    def global_bytes(self):
        return self.parser.plan.total

    def global_records(self):
        n = torch.tensor([self.result().n_records], dtype=torch.int64, device=self.parser.dev)
        dist.all_reduce(n)
        return int(n.item())


class SynthJob:
    """bench.py / tests: one synthetic stream of variable record geometry (synth.SynthStream: 'illumina', 'ont',
    'multiline'; BASELINE.json configs[2..4]) of `total_bytes`, cut at arbitrary bytes into `world` equal shards
    (STRONG scaling: the stream is fixed, a rank holds total / world bytes + halo).  world == 1 parses through
    fqb_parse, world > 1 through the sharded calls (general=True: the sharded general path)."""

    def __init__(self, kind, total_bytes, rank, world, dev, halo_bytes=DEFAULT_HALO, cfg=0, general=None, seed=None,
                 qtable=None):
        from . import synth
        self.kind, self.rank, self.world, self.dev, self.cfg = kind, rank, world, torch.device(dev), cfg
        self.general = (kind == 'multiline') if general is None else bool(general)
        self.exact = False  # general path at world == 1: the exact resolution instead of the speculative pass
        with torch.cuda.device(self.dev):
            self.stream = synth.SynthStream.for_bytes(kind, total_bytes, seed=seed, dev=self.dev, qtable=qtable)
            total = self.stream.total
            cuts = [(g * total) // world for g in range(world + 1)]
            self.cuts = cuts
            own_lens = [cuts[g + 1] - cuts[g] for g in range(world)]
            self.plan = ShardPlan(rank, world, own_lens, halo_bytes)
            self.plan.check()
            self.k_lo, self.k_hi = self.stream.records_from(cuts[rank], cuts[rank + 1])
            n_own = self.k_hi - self.k_lo
            self.table = torch.empty((n_own + 64, 6), dtype=torch.int64, device=self.dev)
            self.result = torch.empty(16, dtype=torch.int64, device=self.dev)
            if world == 1:
                self.parser = None
                self.buf = self.stream.fill(0, total)
                self.n_lines = None
            else:
                self.parser = ShardedParser(self.plan, self.dev, cfg=cfg)
                self.stream.fill(self.plan.offset, self.plan.own_len, out=self.parser.own())
                self.buf = self.parser.buf
            self.max_lines = 0
            if self.general:  # line budget of the general path: the generator knows its lines (+ slack for the halo)
                per_rec = {'multiline': 12, 'illumina': 4, 'ont': 4}[kind]
                k_a, k_b = self.stream.records_from(cuts[rank], min(total, cuts[rank + 1] + self.plan.halo_len()))
                self.max_lines = (k_b - k_a + 2) * per_rec + 4096

    def step(self):
        if self.parser is None:
            if not self.general:
                flags, ml = _lib.FLAG_FAST_ONLY, 0
            elif self.exact:  # line table + hierarchical resolution
                flags, ml = _lib.FLAG_FORCE_GENERAL | _lib.FLAG_NO_SPEC, self.max_lines
            else:             # the speculative single pass alone (no line table)
                flags, ml = _lib.FLAG_FORCE_GENERAL | _lib.FLAG_SPEC_ONLY, 0
            device.parse_raw(self.buf, 1, -1, self.table, None, 0, self.result, _lib.FLAG_CFG(self.cfg) | flags, max_lines=ml)
        elif self.general:
            self.parser.step_general(self.table, max_lines=self.max_lines)
        else:
            self.parser.step(self.table)

    def prepare(self):
        """world > 1, general path: one fast step first (it brings the halo in; step_general reuses it).  world == 1,
        general path: what the product's first call does -- the speculative pass; the exact resolution if it declines."""
        if self.parser is not None and self.general:
            self.parser.step(self.table)
            torch.cuda.synchronize(self.dev)
        elif self.general:
            self.step()
            res = device.read_result(self.result)
            if res.need_general:
                self.exact = True
            # what a streaming caller does from its second chunk on (device.HostParser): the scan geometry that the
            # line density of the first parse asks for
            self.cfg = device.geometry_for_density(self.cfg, res.n_lines, self.buf.numel())

    def read(self):
        if self.parser is None:
            res = device.read_result(self.result)
            if res.error or res.need_general:
                raise RuntimeError('fqb_parse: error %d need_general %d' % (res.error, res.need_general))
            return res
        return self.parser.read()

    def verify_local(self):
        """Rows of this rank against the generator's truth (all of them) and the record range the generator says this
        shard owns.  No collective; returns the number of records verified, raises on any difference (the caller
        sums over the ranks: the ranges tile the stream iff the sum is stream.n)."""
        res = self.read()
        n = int(res.n_records)
        k0 = int(res.reserved[0]) if self.parser is not None else 0
        # The stream ends with the newline of its last record: like the reference's entrypos (pos5 + 2 >= len,
        # src/_fastqandfurious.c:130) the device reports that record as MISSING_QUAL_END with pos0..pos4 set, and
        # readfastq_iter's end-of-stream rule (src/fastqandfurious.py:259-266) completes it on the host.
        last = self.rank == self.world - 1
        want_n = self.k_hi - self.k_lo - (1 if last else 0)
        if (k0, n) != (self.k_lo, want_n):
            raise AssertionError('rank %d emitted records [%d, %d), the generator says [%d, %d)' %
                                 (self.rank, k0, k0 + n, self.k_lo, self.k_lo + want_n))
        bad = self.stream.mismatches(self.table[:n], self.k_lo)
        if bad:
            raise AssertionError('rank %d: %d rows differ from the generator truth' % (self.rank, bad))
        if last:
            goff = self.plan.offset - (1 if self.rank == 0 else 0)
            want = self.stream.truth(self.k_hi - 1, self.k_hi)[0].tolist()
            got = [p + goff for p in res.tail_pos[:5]]
            if res.tail_status != _lib.MISSING_QUAL_END or got != want[:5] or want[5] != self.stream.total - 1:
                raise AssertionError('rank %d: the open last record is %d %s, expected status 5 %s' %
                                     (self.rank, res.tail_status, got, want[:5]))
            n += 1
        return n

    def seam_windows(self, n_seams=None, half=None):
        """[(k_a, k_b, host bytes of records k_a .. k_b - 1, truth rows relative to the window)] for the windows of
        +-`half` bytes around the seams this rank can see in one piece (own tail + halo at world > 1; at world == 1
        the seams an 8-way split would have).  bench.py parses them with the compiled reference."""
        total = self.stream.total
        half = half or max(1 << 20, self.plan.halo_bytes // 2)
        if self.world > 1:
            seams = [self.cuts[self.rank + 1]] if self.rank + 1 < self.world else []
            lo_have, hi_have = self.plan.offset, self.plan.offset + self.buf.numel()
        else:
            m = n_seams or 7
            seams = [(i * total) // (m + 1) for i in range(1, m + 1)]
            lo_have, hi_have = 0, total
        out = []
        for c in seams:
            a, b = max(lo_have, c - half), min(hi_have, c + half)
            q = torch.tensor([a, b], dtype=torch.int64, device=self.dev)
            k_a = int(torch.searchsorted(self.stream.off, q[:1], right=False).item())       # first record starting >= a
            k_b = int(torch.searchsorted(self.stream.off, q[1:], right=True).item()) - 1    # records ending <= b
            if k_b <= k_a:
                continue
            o_a, o_b = int(self.stream.off[k_a].item()), int(self.stream.off[k_b].item())
            data = self.buf[o_a - lo_have:o_b - lo_have].cpu().numpy().tobytes()
            truth = (self.stream.truth(k_a, k_b) - o_a).cpu().numpy()
            out.append((k_a, k_b, data, truth))
        return out

    def global_bytes(self):
        return self.stream.total

    def global_records(self):
        return self.stream.n

    def free(self):
        self.parser = self.buf = self.table = self.stream = None
        device._ws_cache.clear()
        torch.cuda.empty_cache()


def parse_shards_local(data, cuts, halo_bytes, dev='cuda', cfg=0, fused=False, epoch=1, quals_out=None, qual_add=-33,
                       tail=False):
    """The sharded protocol with every shard on ONE device and the exchanges replaced by local copies
    (tests; also documents the protocol).  data: 1-D uint8 CUDA tensor holding the whole stream; cuts:
    increasing byte offsets where shards 1.. start.  Returns (list of per-shard row tensors with absolute
    offsets, FqbResult of the last shard).  fused: the publish / wait exchange (fqb_shard_scan_publish,
    fqb_shard_emit_wait) with every shard's slots in one local tensor instead of peer memory.
    quals_out: a list -> Phred decode (fqb_shard_scan_decode); it receives one (stream offset of the shard, int8
    mirror of own bytes + halo) per shard.  tail: FQB_FLAG_SHARD_TAIL (the scan's last CTA counts, publishes, signals)."""
    dev = torch.device(dev)
    L = _lib.lib()
    flags = _lib.FLAG_CFG(cfg) | (_lib.FLAG_SHARD_TAIL if tail else 0)
    total = data.numel()
    bounds = [0] + list(cuts) + [total]
    world = len(bounds) - 1
    own_lens = [bounds[g + 1] - bounds[g] for g in range(world)]
    plans = [ShardPlan(g, world, own_lens, halo_bytes) for g in range(world)]
    shards = []
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        # fused: slots[destination shard][source shard] = {count, epoch}
        slots = torch.zeros((world, world, 2), dtype=torch.int64, device=dev)
        # one device runs the shards one after the other, so a shard's right neighbour cannot signal in time:
        # every ready flag starts at `epoch` (the signalling store itself is still exercised)
        ready = torch.full((world, 1), int(epoch), dtype=torch.int64, device=dev)
        halo_status = torch.zeros(1, dtype=torch.int32, device=dev)
        tail_ready = torch.zeros((world, 1), dtype=torch.int64, device=dev)  # signals of fqb_shard_scan_publish_ready
        for g, plan in enumerate(plans):
            n = plan.own_len + plan.halo_len()
            if fused:  # the halo pull kernel with local "peer" pointers; the ready flags are set by the shards themselves
                buf = torch.empty(n + 16, dtype=torch.uint8, device=dev)[:n]
                buf[:plan.own_len].copy_(data[plan.offset:plan.offset + plan.own_len])
                src = data[plan.offset + plan.own_len:] if plan.halo_len() else None
                _lib.check(L.fqb_shard_pull_halo(buf.data_ptr() + plan.own_len, src.data_ptr() if src is not None else None,
                                                 plan.halo_len(), ready[g].data_ptr(),
                                                 ready[g - 1].data_ptr() if g else None, epoch, halo_status.data_ptr(),
                                                 stream), 'fqb_shard_pull_halo')
            else:
                buf = data[plan.offset:plan.offset + n].clone()  # "exchange": own bytes + the neighbour's first bytes
            ws = torch.empty(L.fqb_workspace_bytes(n, 0, flags) + 256, dtype=torch.uint8, device=dev)
            own_lines = torch.zeros(1, dtype=torch.int64, device=dev)
            sentinel = 1 if g == 0 else 0
            dst = [slots[r, g].data_ptr() for r in range(g + 1, world)] if fused else []
            pub = (ctypes.c_void_p * max(1, len(dst)))(*dst)
            qual = None
            if quals_out is not None and n:
                raw = torch.empty(n + 32, dtype=torch.int8, device=dev)
                shift = (buf.data_ptr() - raw.data_ptr()) % 16
                qual = raw[shift:shift + n]
                quals_out.append((plan.offset, qual))
            if fused and (g & 1):  # every other shard through the one-kernel form with its ready signal
                _lib.check(L.fqb_shard_scan_publish_ready(buf.data_ptr() if n else None, n, plan.own_len, sentinel,
                                                          own_lines.data_ptr(), pub, len(dst), epoch,
                                                          tail_ready[g - 1].data_ptr(), epoch + 7,
                                                          qual.data_ptr() if qual is not None else None, int(qual_add),
                                                          ws.data_ptr(), ws.numel(), flags, stream),
                           'fqb_shard_scan_publish_ready')
            elif qual is not None:
                _lib.check(L.fqb_shard_scan_decode(buf.data_ptr(), n, plan.own_len, sentinel, own_lines.data_ptr(),
                                                   pub if fused else None, len(dst), epoch if fused else 0, qual.data_ptr(),
                                                   int(qual_add), ws.data_ptr(), ws.numel(), flags, stream),
                           'fqb_shard_scan_decode')
            elif fused:
                _lib.check(L.fqb_shard_scan_publish(buf.data_ptr() if n else None, n, plan.own_len, sentinel,
                                                    own_lines.data_ptr(), pub, len(dst), epoch, ws.data_ptr(), ws.numel(),
                                                    flags, stream), 'fqb_shard_scan_publish')
            else:
                _lib.check(L.fqb_shard_scan(buf.data_ptr() if n else None, n, plan.own_len, sentinel,
                                            own_lines.data_ptr(), ws.data_ptr(), ws.numel(), flags, stream),
                           'fqb_shard_scan')
            shards.append((plan, buf, ws, own_lines, sentinel))
        assert int(halo_status.item()) == 0
        if fused:  # odd shards signalled their left neighbour from the scan's epilogue
            assert tail_ready[:, 0].tolist() == [epoch + 7 if (g + 1 < world and (g + 1) & 1) else 0 for g in range(world)]
        gathered = torch.cat([s[3] for s in shards])  # "all-gather"
        incl = torch.cumsum(gathered, 0)
        rows, last = [], None
        for g, (plan, buf, ws, own_lines, sentinel) in enumerate(shards):
            base = (incl[g:g + 1] - gathered[g:g + 1]).contiguous()
            n = buf.numel()
            table = torch.empty((n // 8 + 64, 6), dtype=torch.int64, device=dev)
            result = torch.zeros(16, dtype=torch.int64, device=dev)
            if fused:
                _lib.check(L.fqb_shard_emit_wait(buf.data_ptr() if n else None, n, plan.own_len, sentinel,
                                                 1 if plan.is_last else 0, plan.offset - sentinel,
                                                 slots[g].data_ptr() if g else None, g, epoch, table.data_ptr(),
                                                 table.shape[0], result.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                                                 stream), 'fqb_shard_emit_wait')
            else:
                _lib.check(L.fqb_shard_emit(buf.data_ptr() if n else None, n, plan.own_len, sentinel,
                                            1 if plan.is_last else 0, plan.offset - sentinel, base.data_ptr(),
                                            table.data_ptr(), table.shape[0], result.data_ptr(), ws.data_ptr(), ws.numel(),
                                            flags, stream), 'fqb_shard_emit')
            res = device.read_result(result)
            if res.error:
                return None, res
            rows.append(table[:res.n_records].clone())
            last = res
    return rows, last


def parse_shards_local_general(data, cuts, halo_bytes, dev='cuda', cfg=0, epoch=1, max_lines=None):
    """The sharded GENERAL path with every shard on one device, one after the other (the hand-over slots are
    local).  Returns (list of per-shard row tensors, list of per-shard FqbResult)."""
    dev = torch.device(dev)
    L = _lib.lib()
    flags = _lib.FLAG_CFG(cfg)
    total = data.numel()
    bounds = [0] + list(cuts) + [total]
    world = len(bounds) - 1
    own_lens = [bounds[g + 1] - bounds[g] for g in range(world)]
    plans = [ShardPlan(g, world, own_lens, halo_bytes) for g in range(world)]
    rows, results = [], []
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        slots = torch.zeros((world + 1, 8), dtype=torch.int64, device=dev)
        for g, plan in enumerate(plans):
            n = plan.own_len + plan.halo_len()
            buf = data[plan.offset:plan.offset + n].clone()
            ml = (n // 2 + 64) if max_lines is None else max_lines
            ws = torch.empty(L.fqb_workspace_bytes(n, ml, flags) + 256, dtype=torch.uint8, device=dev)
            table = torch.empty((n // 8 + 64, 6), dtype=torch.int64, device=dev)
            result = torch.zeros(16, dtype=torch.int64, device=dev)
            sentinel = 1 if g == 0 else 0
            _lib.check(L.fqb_shard_general(buf.data_ptr() if n else None, n, plan.own_len, sentinel, 1 if g == 0 else 0,
                                           1 if plan.is_last else 0, plan.offset - sentinel, slots[g].data_ptr(),
                                           slots[g + 1].data_ptr() if g + 1 < world else None,
                                           slots[g - 1].data_ptr() if g else None, epoch, 0, table.data_ptr(),
                                           table.shape[0], result.data_ptr(), ws.data_ptr(), ws.numel(), ml, flags, stream),
                       'fqb_shard_general')
            res = device.read_result(result)
            results.append(res)
            rows.append(table[:res.n_records].clone() if not res.error else table[:0].clone())
    return rows, results


# ---- FASTA over byte-range shards (SURVEY.md 8e x 8f4) ---------------------------------------------------------------
# The chain of entrypos_fasta calls (csrc/fq_fasta.cuh) needs no state from the shard before it except the PARITY of the
# run of consecutive "\n>" lines a shard may start in, and that parity is reset by any line that is not a header.  So a
# shard parses a window [c_g - look-behind, c_{g+1} + halo) of the stream with the ordinary single-buffer call; when
# the look-behind holds one newline that is not followed by '>' (any sequence line), every on-chain decision behind it
# equals the whole-stream parse's.  The shard owns the records whose "\n>" newline lies in [c_g, c_{g+1}) (shard 0 also
# the virtual sentinel); the record behind its last own one must start inside the halo (else FastaShardError: the
# reference's "buffer must hold the largest entry" contract, src/fastqandfurious.py:219-223).  Communication: the
# neighbours' bytes (one exchange) and one all-gather of two int64 per rank (own records, error code).  The reference
# has no counterpart.
class FastaShardError(RuntimeError):
    pass


FASTA_ERR_PARITY, FASTA_ERR_HALO, FASTA_ERR_NO_START, FASTA_ERR_PARSE = 1, 2, 3, 4
_FASTA_ERR_TEXT = {
    FASTA_ERR_PARITY: 'the look-behind of a shard holds no line that is not a header: enlarge lookbehind_bytes',
    FASTA_ERR_HALO: 'a record a shard owns does not end inside its halo: enlarge halo_bytes',
    FASTA_ERR_NO_START: 'the last shard (with its look-behind) holds no record start while earlier shards do',
    FASTA_ERR_PARSE: 'the single-buffer parse of the shard\'s window failed (see the exception chained on that rank)',
}


def fasta_shard_window(total, cuts, g, halo_bytes, lookbehind_bytes):
    """The byte window of the stream shard g parses and the newline positions it owns.
    cuts = [0, c_1, ..., total].  Returns (lo, hi, own_lo, own_hi): window [lo, hi); the shard owns the records whose
    "\\n>" newline sits at a stream offset in [own_lo, own_hi) (-1: the virtual sentinel in front of the stream)."""
    world = len(cuts) - 1
    lo = 0 if g == 0 else max(0, cuts[g] - int(lookbehind_bytes))
    hi = total if g == world - 1 else min(total, cuts[g + 1] + int(halo_bytes))
    return lo, hi, (-1 if g == 0 else cuts[g]), cuts[g + 1]


def fasta_shard_own(window, lo, hi, own_lo, own_hi, total, parse):
    """Parse one shard's window (a 1-D uint8 tensor holding stream[lo:hi]) and keep what the shard owns.
    parse(buf, sentinel, goff) -> (rows int64[n,4] tensor, status, tail_pos[4], resume_offset) is the single-buffer
    FASTA call (device.parse_fasta_buffer on the GPU).  Returns (rows of the own records as ABSOLUTE stream offsets,
    err, tail) with tail = (status, tail_pos in the blob coordinates of a whole-stream parse with sentinel) for a
    window that reaches the end of the stream, else None."""
    sentinel = lo == 0
    if not sentinel:  # the parity of the first run must be decided inside the look-behind
        seg = window[:own_lo - lo + 1]
        if not bool(((seg[:-1] == 10) & (seg[1:] != 62)).any()):
            return None, FASTA_ERR_PARITY, None
    rows, status, tail_pos, _ = parse(window, sentinel, -1 if sentinel else lo)
    shift = -1 if sentinel else lo  # blob position of the window's parse -> stream offset
    pos0 = rows[:, 0].contiguous()
    bounds = torch.tensor([own_lo + 1, own_hi + 1], dtype=torch.int64, device=pos0.device)
    i_lo, i_hi = (int(v) for v in torch.searchsorted(pos0, bounds))
    reaches_end = hi == total
    if status != 0 and not reaches_end:
        # the call that is not COMPLETE: an own record that does not end inside the halo?
        if tail_pos[0] + shift - 1 < own_hi:
            return None, FASTA_ERR_HALO, None
    tail = None
    if reaches_end:
        tail = (int(status), [int(p) + shift + 1 if p >= 0 else -1 for p in tail_pos])
    return rows[i_lo:i_hi], 0, tail


def _fasta_device_parse(cfg=0):
    def parse(buf, sentinel, goff):
        res = device.parse_fasta_buffer(buf, sentinel=sentinel, goff=goff, cfg=cfg)
        return res.table, res.tail_status, res.tail_pos, res.resume_offset
    return parse


def _fasta_finish(counts, errs, tail):
    """Global result from the shards' counts, error codes and the last shard's tail."""
    for g, e in enumerate(errs):
        if e:
            raise FastaShardError('shard %d: %s' % (g, _FASTA_ERR_TEXT[e]))
    n = int(sum(counts))
    status, tail_pos = tail
    if status == 0 and n > 0:
        raise FastaShardError(_FASTA_ERR_TEXT[FASTA_ERR_NO_START])
    # the offset of the call that is not COMPLETE: pos3 of the record before it = the newline of its own "\n>"
    resume = tail_pos[0] - 1 if (n >= 1 and status != 0) else 0
    return n, status, tail_pos, resume


def parse_fasta_shards_local(data, cuts, halo_bytes=DEFAULT_HALO, lookbehind_bytes=1 << 16, parse=None, cfg=0):
    """All shards of one stream on ONE device, one after the other (tests, and the reference point for the distributed
    form): `data` is the whole stream (1-D uint8 tensor), `cuts` the interior cut offsets.  Returns
    (rows per shard -- ABSOLUTE offsets, the concatenation is the whole-stream table --, n_records, tail_status,
    tail_pos, resume_offset) exactly as one parse_fasta_buffer(data) call reports them."""
    parse = parse or _fasta_device_parse(cfg)
    total = int(data.numel())
    cuts = [0] + [int(c) for c in cuts] + [total]
    out, counts, errs, tail = [], [], [], None
    for g in range(len(cuts) - 1):
        lo, hi, own_lo, own_hi = fasta_shard_window(total, cuts, g, halo_bytes, lookbehind_bytes)
        rows, err, t = fasta_shard_own(data[lo:hi], lo, hi, own_lo, own_hi, total, parse)
        out.append(rows)
        counts.append(0 if rows is None else int(rows.shape[0]))
        errs.append(err)
        if g == len(cuts) - 2:
            tail = t
    n, status, tail_pos, resume = _fasta_finish(counts, errs, tail if tail is not None else (0, [-1] * 4))
    return out, n, status, tail_pos, resume


class ShardedFastaParser:
    """One process per GPU: rank g holds stream[c_g : c_{g+1}] on its device, receives the last `lookbehind_bytes` of
    its left neighbour and the first `halo_bytes` of its right neighbour (one batch of send / recv), parses the window
    with the single-buffer FASTA call and keeps the records it owns; one all-gather of (own records, error code) gives
    every rank the index of its first record, and the last rank's tail is broadcast.  Works on CUDA tensors over NCCL
    and on CPU tensors over gloo (tests, with `parse` = an oracle-backed callable)."""

    def __init__(self, rank, world, own_lens, halo_bytes=DEFAULT_HALO, lookbehind_bytes=1 << 16, group=None, parse=None, cfg=0):
        self.rank, self.world, self.group = int(rank), int(world), group
        self.own_lens = [int(x) for x in own_lens]
        self.cuts = [0]
        for x in self.own_lens:
            self.cuts.append(self.cuts[-1] + x)
        self.total = self.cuts[-1]
        self.halo, self.lookbehind = int(halo_bytes), int(lookbehind_bytes)
        for g in range(self.world):  # a window must come from the direct neighbours alone
            if g > 0 and self.own_lens[g - 1] < min(self.lookbehind, self.cuts[g]) and g - 1 != 0:
                raise ValueError('shard %d is shorter than the look-behind' % (g - 1))
            if g + 1 < self.world - 1 and self.own_lens[g + 1] < self.halo:
                raise ValueError('shard %d is shorter than the halo' % (g + 1))
        self.parse_fn = parse or _fasta_device_parse(cfg)

    def window_of(self, g):
        return fasta_shard_window(self.total, self.cuts, g, self.halo, self.lookbehind)

    def parse(self, own):
        """own: 1-D uint8 tensor with this rank's bytes.  Returns (first_record_index, rows int64[n,4] of ABSOLUTE
        stream offsets, n_records, tail_status, tail_pos, resume_offset) -- the last four as a whole-stream
        parse_fasta_buffer call reports them, identical on every rank."""
        r, w = self.rank, self.world
        if own.numel() != self.own_lens[r]:
            raise ValueError('own bytes: %d, plan: %d' % (own.numel(), self.own_lens[r]))
        lo, hi, own_lo, own_hi = self.window_of(r)
        c0, c1 = self.cuts[r], self.cuts[r + 1]
        window = torch.empty(hi - lo, dtype=torch.uint8, device=own.device)
        window[c0 - lo:c1 - lo] = own
        ops = []
        if r > 0:  # my head is the left neighbour's halo, its tail my look-behind
            l_lo, l_hi, _, _ = self.window_of(r - 1)
            if l_hi > c0:
                ops.append(dist.P2POp(dist.isend, own[:l_hi - c0].contiguous(), _peer(r - 1, self.group), self.group))
            if c0 > lo:
                ops.append(dist.P2POp(dist.irecv, window[:c0 - lo], _peer(r - 1, self.group), self.group))
        if r < w - 1:
            r_lo, r_hi, _, _ = self.window_of(r + 1)
            if r_lo < c1:
                ops.append(dist.P2POp(dist.isend, own[r_lo - c0:].contiguous(), _peer(r + 1, self.group), self.group))
            if hi > c1:
                ops.append(dist.P2POp(dist.irecv, window[c1 - lo:], _peer(r + 1, self.group), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        failure = None
        try:
            rows, err, tail = fasta_shard_own(window, lo, hi, own_lo, own_hi, self.total, self.parse_fn)
        except Exception as exc:  # every rank must still reach the collectives below; the error is agreed there
            rows, err, tail, failure = None, FASTA_ERR_PARSE, None, exc
        mine = torch.tensor([0 if rows is None else int(rows.shape[0]), int(err)], dtype=torch.int64, device=own.device)
        allv = torch.empty(2 * w, dtype=torch.int64, device=own.device)
        dist.all_gather_into_tensor(allv, mine, group=self.group)
        allv = allv.cpu().tolist()
        counts, errs = allv[0::2], allv[1::2]
        tail_t = torch.full((5,), -1, dtype=torch.int64, device=own.device)
        if r == w - 1 and tail is not None:
            tail_t = torch.tensor([tail[0]] + list(tail[1]), dtype=torch.int64, device=own.device)
        dist.broadcast(tail_t, _peer(w - 1, self.group), group=self.group)
        tail_l = tail_t.cpu().tolist()
        if failure is not None:
            raise FastaShardError('shard %d: %s' % (r, _FASTA_ERR_TEXT[FASTA_ERR_PARSE])) from failure
        n, status, tail_pos, resume = _fasta_finish(counts, errs, (tail_l[0], tail_l[1:]))
        return int(sum(counts[:r])), rows, n, status, tail_pos, resume
