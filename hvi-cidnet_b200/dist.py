"""Multi-GPU plumbing for the CIDNet forward: images are independent units (the attention is per
image), so a batch is partitioned contiguously over the ranks of one node, every rank
runs the single-GPU path on its share and NO data-path collective is needed.  The only collective is
the optional gather of the outputs onto every rank."""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced partition of `n_items` over `world_size` ranks: (start, count).
    The first `n_items % world_size` ranks get one extra item; empty shards are allowed."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def forward_sharded(model, x_full, group=None, gather=True):
    """Run `model` on this rank's share of the batch `x_full` ([B,3,H,W], present on every rank, on the
    rank's own device).  With `gather` the per-rank outputs are all-gathered (ragged shards supported)
    and the full [B,3,H,W] result is returned on every rank; otherwise only the local shard's output."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    start, count = shard_range(x_full.shape[0], world, rank)
    local = x_full[start:start + count]
    out = model(local) if count > 0 else x_full.new_empty((0,) + tuple(x_full.shape[1:]))
    if not gather or world == 1:
        return out
    counts = [shard_range(x_full.shape[0], world, r)[1] for r in range(world)]
    cmax = max(counts)
    pad = out.new_zeros((cmax,) + tuple(out.shape[1:]))
    pad[:count] = out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


# ----------------------------------------------------------------------------------------------
# Single image, rows sharded over the GPUs of one node (BASELINE.json configs[4]).
# The C library drives the schedule (cidnet_forward_sharded) and calls back into the two
# functions below whenever data has to cross ranks: halo rows of NHWC activations to / from the
# two neighbours, and one all-reduce of the raw [Gram | sum q^2 | sum k^2] per LCA stage.
# ----------------------------------------------------------------------------------------------
import ctypes as _C


class StripComm:
    """Transport of the halo / all-reduce callbacks over a torch.distributed process group.

    `buffer` is the torch tensor (uint8, 1-D) the library's pointers point into (the workspace).
      * NCCL group, CUDA buffer: P2P send/recv and all_reduce act directly on views of the
        workspace -- rows are contiguous in NHWC, nothing is packed or staged; NVLink carries them.
      * any other combination (gloo group; CPU buffer in the dry-run tests; two ranks sharing one
        GPU in the single-GPU parity test): the same rows are staged through host memory.
    """

    def __init__(self, buffer, group=None):
        self.buf = buffer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.direct = buffer.is_cuda and dist.get_backend(group) == "nccl"
        self.error = None
        self.log = []                                   # ("halo", [(offset, row_bytes, rows, top, bot)...]) / ("allreduce", offset, count)
        self.bytes_sent = 0
        self.trace = None                               # tests: set to [] to record a fingerprint of every block sent / received
        self.halo_cb = _clib().HALO_FN(self._halo)
        self.allreduce_cb = _clib().ALLREDUCE_FN(self._allreduce)

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def _view(self, ptr, nbytes):
        off = ptr - self.buf.data_ptr()
        if off < 0 or off + nbytes > self.buf.numel():
            raise RuntimeError("callback pointer outside the registered workspace")
        return self.buf[off:off + nbytes]

    def _halo(self, user, reqs, n):
        try:
            sends, recvs, entry = [], [], []
            for i in range(n):
                q = reqs[i]
                t = self._view(q.base, q.rows * q.row_bytes).view(q.rows, q.row_bytes)
                entry.append((q.base - self.buf.data_ptr(), q.row_bytes, q.rows, q.halo_top, q.halo_bot))
                if q.halo_top > 0:                       # neighbour above: rank - 1
                    sends.append((t[q.halo_top:2 * q.halo_top], self.rank - 1))
                    recvs.append((t[:q.halo_top], self.rank - 1))
                if q.halo_bot > 0:                       # neighbour below: rank + 1
                    sends.append((t[q.rows - 2 * q.halo_bot:q.rows - q.halo_bot], self.rank + 1))
                    recvs.append((t[q.rows - q.halo_bot:], self.rank + 1))
            self.log.append(("halo", entry))
            self.bytes_sent += sum(s.numel() for s, _ in sends)
            if self.trace is not None:
                sent_fp = [(p, _fingerprint(s_)) for s_, p in sends]
            if self.direct:
                ops = [dist.P2POp(dist.isend, s, self._peer(p), self.group) for s, p in sends]
                ops += [dist.P2POp(dist.irecv, r, self._peer(p), self.group) for r, p in recvs]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                             # stream-ordered for NCCL (no host sync)
            else:
                hs = [s.cpu().contiguous() for s, _ in sends]
                hr = [torch.empty(r.shape, dtype=r.dtype) for r, _ in recvs]
                ops = [dist.P2POp(dist.isend, s, self._peer(p), self.group) for s, (_, p) in zip(hs, sends)]
                ops += [dist.P2POp(dist.irecv, r, self._peer(p), self.group) for r, (_, p) in zip(hr, recvs)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                for h, (r, _) in zip(hr, recvs):
                    r.copy_(h)
            if self.trace is not None:
                self.trace.append(("halo", sent_fp, [(p, _fingerprint(r_)) for r_, p in recvs]))
            return 0
        except BaseException as e:                       # never let an exception cross the C boundary
            self.error = e
            return -1

    def _allreduce(self, user, ptr, count):
        try:
            t = self._view(ptr, 4 * count).view(torch.float32)
            self.log.append(("allreduce", ptr - self.buf.data_ptr(), int(count)))
            before = t.detach().double().cpu().clone() if self.trace is not None else None
            if self.direct:
                dist.all_reduce(t, group=self.group)
            else:
                h = t.cpu()
                dist.all_reduce(h, group=self.group)
                t.copy_(h)
            if self.trace is not None:
                self.trace.append(("allreduce", before, t.detach().double().cpu().clone()))
            return 0
        except BaseException as e:
            self.error = e
            return -1


def _clib():
    from . import _lib
    return _lib


def _fingerprint(t):
    """(byte sum, position-weighted byte sum) of a uint8 block: equal for the block a rank sent and the block its
    neighbour received (StripComm.trace, used by the CPU tests of the exchange schedule)"""
    b = t.detach().reshape(-1).to("cpu", torch.int64)
    return int(b.sum()), int((b * (torch.arange(b.numel(), dtype=torch.int64) % 8191 + 1)).sum())


def strip_plan(H, world_size, rank, halo=16):
    """cidnet_shard for `rank`: balanced split of the H/8 coarsest rows (no device needed)."""
    L = _clib()
    sh = L.Shard()
    L.check(L.lib().cidnet_shard_plan(int(H), int(world_size), int(rank), int(halo), _C.byref(sh)))
    return sh


def strip_local_range(sh):
    """(first, last+1) global rows of the LOCAL image of shard `sh` (owned rows + halos)."""
    top = sh.halo if (sh.nranks > 1 and sh.rank > 0) else 0
    bot = sh.halo if (sh.nranks > 1 and sh.rank < sh.nranks - 1) else 0
    return sh.row_begin - top, sh.row_end + bot


class RowShardedCIDNet:
    """One image, rows partitioned over the ranks of `group` (one process per GPU).

        net = RowShardedCIDNet(model)            # model: hvi_cidnet_b200 CIDNet on this rank's GPU
        y = net(x)                               # x: [1,3,H,W] full image (CPU pinned or CUDA), same on every rank
                                                 # y: full [1,3,H,W] result on every rank (gather=True)

    Each rank uploads / computes only its strip (+ halo); conv halos and the partial Gram sums
    travel over NCCL (NVLink) -- see include/cidnet_b200.h, "rows sharded over the GPUs of one node".

    CUDA-graph replay (EXPERIMENTAL, opt-in: `graph=True` or CIDNET_SHARD_GRAPH=1): the ~90 kernel launches of the
    strip AND the NCCL halo send/recvs and all-reduces the library's callbacks issue are captured into one CUDA
    graph on the third call with the same (shape, flags) and replayed afterwards -- at 8 GPUs the eager schedule
    spends more time in host-side launches (89 kernels + 15 NCCL calls driven from Python) than on the GPU.
    Every rank captures / replays in the same call, so the collectives stay matched.  With replay the returned
    tensor is a STATIC buffer that the next call overwrites.  Only taken with an NCCL group on CUDA buffers (the
    gloo-staged transport of the single-GPU tests copies through the host and cannot be captured).
    """

    def __init__(self, model, group=None, halo=16, graph=None, transport=None):
        """transport: "peer" -- the strips' workspaces are CUDA-IPC mapped into every rank and the halo rows / partial
        attention statistics are read straight from the neighbours' memory by kernels of the library (no NCCL on the data
        path, the strip forward is one CUDA graph inside the library; needs one GPU per rank on one node);
        "callbacks" -- torch.distributed send/recv + all_reduce issued from the library's host callbacks (NCCL, or gloo
        staged through the host); None -- "peer" when every rank has its own GPU and the group is NCCL, else "callbacks"
        (CIDNET_SHARD_TRANSPORT overrides)."""
        import os
        self.model, self.group, self.halo = model, group, int(halo)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._ws = {}
        self.comm = None
        self.use_graph = bool(graph) if graph is not None else os.environ.get("CIDNET_SHARD_GRAPH", "0") == "1"
        self._graphs = {}                 # key -> dict(seen, g, x, out, sig)
        self.replays = 0
        transport = transport or os.environ.get("CIDNET_SHARD_TRANSPORT")
        if transport is None:
            own_gpu = (torch.cuda.is_available() and self.world > 1 and self.world <= torch.cuda.device_count()
                       and dist.is_initialized() and dist.get_backend(group) == "nccl")
            transport = "peer" if own_gpu else "callbacks"
        if transport not in ("peer", "callbacks"):
            raise ValueError(f"unknown transport {transport!r}")
        self.transport = transport if self.world > 1 else "callbacks"
        self._peer = {}                   # (H, W, device) -> dict(own, ptrs, arr, nbytes)
        self.peer_forwards = 0

    def close(self):
        """Drop the captured CUDA graphs and the peer mappings.  Call before `dist.destroy_process_group()`: tearing the
        NCCL communicator down while instantiated graphs still hold its kernels blocks (measured: the 2-rank probe hung
        in destroy_process_group until the graphs were released first); the peer workspaces are unmapped by every rank
        before their owners free them (barrier)."""
        if self._graphs or self._peer:
            torch.cuda.synchronize()
        self._graphs = {}
        if self._peer:
            L = _clib()
            for ent in self._peer.values():
                for r, p in enumerate(ent["ptrs"]):
                    if r != self.rank and p:
                        L.lib().cidnet_peer_close(_C.c_void_p(p))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for ent in self._peer.values():
                L.lib().cidnet_peer_free(_C.c_void_p(ent["own"]))
            self._peer = {}

    # ------------------------------------------------------------------ peer-memory transport
    def _peer_setup(self, H_global, W, dev):
        """allocate this rank's strip workspace, exchange the CUDA IPC handles, map every other rank's workspace"""
        key = (int(H_global), int(W), str(dev))
        ent = self._peer.get(key)
        if ent is not None:
            return ent
        L = _clib()
        lib = L.lib()
        nbytes = int(lib.cidnet_peer_workspace_bytes(int(H_global), int(W), self.world, self.halo))
        if nbytes <= 0:
            raise RuntimeError("cidnet_peer_workspace_bytes failed: " + lib.cidnet_last_error().decode())
        own = _C.c_void_p()
        handle = _C.create_string_buffer(64)
        L.check(lib.cidnet_peer_alloc(dev.index, nbytes, _C.byref(own), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(own.value)
            else:
                p = _C.c_void_p()
                L.check(lib.cidnet_peer_open(dev.index, handles[r], _C.byref(p)))
                ptrs.append(p.value)
        dist.barrier(group=self.group)                      # everybody has mapped everybody before the first exchange
        ent = {"own": own.value, "ptrs": ptrs, "arr": (_C.c_void_p * self.world)(*ptrs), "nbytes": nbytes}
        self._peer = {key: ent}
        return ent

    def peer_error(self):
        """1 if a handshake of the peer transport timed out (a partner never arrived), else 0"""
        L = _clib()
        err = _C.c_int(0)
        for ent in self._peer.values():
            L.check(L.lib().cidnet_peer_error(_C.c_void_p(ent["own"]), _C.byref(err)))
        return int(err.value)

    def _forward_strip_peer(self, x_local, H_global, sh, rows):
        L = _clib()
        m, t = self.model, self.model.trans
        dev = x_local.device
        W = x_local.shape[3]
        with torch.cuda.device(dev):
            ctx = m._ensure_ctx(dev)
            ent = self._peer_setup(H_global, W, dev)
            t._note_hvit_called()
            kd = t.density_k.detach() if t.density_k.dtype == torch.float32 else t._this_k_dev
            out = torch.empty_like(x_local)
            L.check(L.lib().cidnet_forward_sharded_peer(
                ctx, x_local.data_ptr(), out.data_ptr(), W, _C.byref(sh), ent["arr"], ent["nbytes"], kd.data_ptr(),
                int(bool(t.gated)), float(t.alpha_s), int(bool(t.gated2)), float(t.alpha), L.stream_ptr(dev)))
            self.peer_forwards += 1
        return out, sh

    def _workspace(self, rows, W, device):
        key = (rows, W, str(device))
        if key not in self._ws:
            L = _clib()
            n = L.lib().cidnet_workspace_bytes(1, rows, W)
            ws = torch.empty(n + 1024, dtype=torch.uint8, device=device)
            self._ws = {key: (ws, StripComm(ws, self.group) if self.world > 1 else None)}
            self._graphs = {}             # captured graphs point into the previous workspace
        return self._ws[key]

    def forward_strip(self, x_local, H_global):
        """x_local: CUDA fp32 [1,3,local_rows,W] = this rank's local image.  Returns the local output
        [1,3,local_rows,W]; only the owned rows (see strip_plan / strip_local_range) are meaningful."""
        L = _clib()
        m = self.model
        dev = x_local.device
        sh = strip_plan(H_global, self.world, self.rank, self.halo)
        rows = L.lib().cidnet_shard_local_rows(_C.byref(sh))
        if tuple(x_local.shape) != (1, 3, rows, x_local.shape[3]) or x_local.dtype != torch.float32 or not x_local.is_cuda:
            raise RuntimeError(f"forward_strip expects a CUDA fp32 [1,3,{rows},W] local image, got {tuple(x_local.shape)} on {dev}")
        W = x_local.shape[3]
        x_local = x_local.contiguous()
        t = m.trans
        if self.transport == "peer" and self.world > 1:
            return self._forward_strip_peer(x_local, H_global, sh, rows)
        with torch.cuda.device(dev):
            ctx = m._ensure_ctx(dev)
            ws, comm = self._workspace(rows, W, dev)
            self.comm = comm
            t._note_hvit_called()
            key = (rows, W, str(dev), bool(t.gated), float(t.alpha_s), bool(t.gated2), float(t.alpha), m._synced,
                   t.density_k.data_ptr())
            ent = self._graphs.setdefault(key, {"seen": 0, "g": None})
            can_graph = self.use_graph and (comm is None or comm.direct)
            if can_graph and ent["g"] is not None:
                ent["x"].copy_(x_local, non_blocking=True)
                ent["g"].replay()
                self.replays += 1
                return ent["out"], sh

            def run(xin, xout):
                if comm is not None:
                    comm.log, comm.bytes_sent = [], 0            # per-forward record of the exchanges
                off = (-ws.data_ptr()) % 1024
                kd = t.density_k.detach()
                rc = L.lib().cidnet_forward_sharded(
                    ctx, xin.data_ptr(), xout.data_ptr(), W, _C.byref(sh), ws.data_ptr() + off, ws.numel() - off,
                    kd.data_ptr() if kd.dtype == torch.float32 else None,
                    int(bool(t.gated)), float(t.alpha_s), int(bool(t.gated2)), float(t.alpha),
                    comm.halo_cb if comm else L.HALO_FN(), comm.allreduce_cb if comm else L.ALLREDUCE_FN(), None,
                    L.stream_ptr(dev))
                if comm is not None and comm.error is not None:
                    err, comm.error = comm.error, None
                    raise err
                L.check(rc)

            ent["seen"] += 1
            if can_graph and ent["seen"] >= 3:
                # two eager calls have set the kernels' attributes and opened NCCL's peer connections
                sx, sout = x_local.clone(), torch.empty_like(x_local)
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                try:
                    with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                        run(sx, sout)
                except Exception as e:      # keep working eagerly; every rank sees the same failure mode
                    self.use_graph = False
                    self.graph_error = repr(e)
                    torch.cuda.synchronize(dev)
                else:
                    torch.cuda.current_stream(dev).wait_stream(side)
                    ent.update(g=g, x=sx, out=sout)
                    g.replay()
                    self.replays += 1
                    return sout, sh
            out = torch.empty_like(x_local)
            run(x_local, out)
        return out, sh

    def __call__(self, x, gather=True):
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 3:
            raise RuntimeError(f"RowShardedCIDNet expects one image [1,3,H,W], got {tuple(x.shape)}")
        H, W = int(x.shape[2]), int(x.shape[3])
        dev = self.model.trans.density_k.device
        sh = strip_plan(H, self.world, self.rank, self.halo)
        a, b = strip_local_range(sh)
        x_local = x[:, :, a:b, :].to(dev, torch.float32, non_blocking=True).contiguous()
        out, sh = self.forward_strip(x_local, H)
        own = out[:, :, sh.row_begin - a:sh.row_end - a, :]
        if not gather or self.world == 1:
            return own.contiguous()
        plans = [strip_plan(H, self.world, r, self.halo) for r in range(self.world)]
        nmax = max(p.row_end - p.row_begin for p in plans)
        pad = own.new_zeros((1, 3, nmax, W))
        pad[:, :, :own.shape[2]] = own
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([bf[:, :, :p.row_end - p.row_begin] for bf, p in zip(bufs, plans)], dim=2)
