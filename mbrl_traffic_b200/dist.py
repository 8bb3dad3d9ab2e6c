"""Multi-GPU plumbing: one process per GPU over ``torch.distributed``.

Two ways the path shards:

* **Road batch** -- roads are independent (arz.py:508-513,
  evaluate_model.py:388-403): ``shard_range`` gives each rank a contiguous
  range; there is no data-path collective.
* **One long road** -- ``DomainDecomposedRoad`` splits the cell axis into
  contiguous blocks with ``halo`` ghost cells per internal edge.  The step
  kernel leaves ghost edges alone (``MT_BC_HALO``); after every ``halo`` steps
  neighbouring ranks swap their edge cells (point-to-point send/recv: NCCL on
  GPUs, gloo in the CPU tests).  Stencil radius is one cell per step for LWR
  and ARZ, so ``halo`` steps can run between exchanges (temporal blocking:
  the exchange is latency-bound, 2*halo values per side).
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous ``[lo, hi)`` of ``n_items`` owned by ``rank`` (remainder
    spread over the first ranks)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DomainDecomposedRoad(object):
    """One road of ``n_cells`` cells split over the ranks of a process group.

    Parameters
    ----------
    n_cells : int
        global number of cells
    rank, world : int
        this rank and the number of ranks
    halo : int
        ghost cells per internal edge = steps between exchanges
    stepper : callable
        ``stepper(local_obs, bc_left, bc_right) -> new_local_obs`` advancing
        the local block ``[rho | v]`` (shape ``(2 * n_local,)`` or
        ``(1, 2 * n_local)``) by one step; ``bc_*`` is ``"physical"`` on a
        true road end and ``"halo"`` on an internal edge.
    exchange : callable or None
        ``exchange(send_left, send_right) -> (recv_left, recv_right)`` with
        None on missing sides; defaults to ``torch.distributed`` send/recv.
    """

    def __init__(self, n_cells, rank, world, halo, stepper, exchange=None):
        if halo < 1:
            raise ValueError("halo must be >= 1")
        self.n_cells, self.rank, self.world, self.halo = n_cells, rank, world, halo
        self.lo, self.hi = shard_range(n_cells, rank, world)
        if self.hi - self.lo < halo + 1:
            raise ValueError("blocks must be wider than the halo")
        self.has_left = rank > 0
        self.has_right = rank < world - 1
        self.n_local = (self.hi - self.lo) + halo * (self.has_left + self.has_right)
        physical = getattr(stepper, "physical", ())
        if world > 1 and any(code == 0 for code in physical):   # MT_BC_LOOP
            raise ValueError(
                "the reference 'loop' rule copies cells of the road's far end "
                "(lwr.py:237-244, arz.py:316-328), which a block of a decomposed "
                "road does not hold: use extend / constant road ends")
        self.stepper = stepper
        self.exchange = exchange or self._torch_exchange
        self.peer = None               # a PeerHalo: exchange over peer memory instead
        self._since_exchange = 0
        self._push_pending = False     # the fused path has pushed; the ghosts await a pull
        #: False keeps the exchange in its own push / pull launches even when the
        #: fused path (PeerHalo.advance) applies -- for A/B measurements
        self.fuse_exchange = True

    # ------------------------------------------------------------------ #
    def scatter(self, global_obs):
        """Cut this rank's block (with ghosts) out of a global ``[rho | v]``."""
        n, h = self.n_cells, self.halo
        a = self.lo - (h if self.has_left else 0)
        b = self.hi + (h if self.has_right else 0)
        g = np.asarray(global_obs).reshape(-1) if not _is_torch(global_obs) \
            else global_obs.reshape(-1)
        rho, v = g[a:b], g[n + a:n + b]
        if _is_torch(g):
            import torch
            return torch.cat((rho, v)).contiguous()
        return np.concatenate((rho, v))

    def begin(self, global_obs):
        """``scatter`` for a NEW run on this object: besides cutting the block
        out it forgets the previous run's exchange state.  A run that ended on
        the fused path leaves one push in the mailboxes (its ghost cells were
        never pulled); that push is consumed here, so that the first launch of
        the new run uses the fresh ghosts it was handed instead of pulling the
        old run's edge cells.  Every rank must call it."""
        x = self.scatter(global_obs)
        if _is_torch(x):
            x = x.clone()
        if self._push_pending:
            self.peer.pull(x.clone(), self.n_local)     # (result discarded)
            self._push_pending = False
        self._since_exchange = 0
        return x

    def own(self, local_obs):
        """The cells this rank owns (ghosts stripped) as ``(rho, v)``."""
        x = local_obs.reshape(-1)
        h, nl = self.halo, self.n_local
        a = h if self.has_left else 0
        b = nl - (h if self.has_right else 0)
        return x[a:b], x[nl + a:nl + b]

    # ------------------------------------------------------------------ #
    def needs_refresh(self):
        """True when the ghost cells have been consumed (``halo`` steps ran)."""
        return self._since_exchange >= self.halo

    def step(self, local_obs):
        """One global time step; exchanges halos when the ghosts are used up."""
        if self.needs_refresh():
            local_obs = self.refresh(local_obs)
        out = self.stepper(local_obs,
                           "halo" if self.has_left else "physical",
                           "halo" if self.has_right else "physical")
        self._since_exchange += 1
        return out

    def advance(self, local_obs, n_steps):
        """``n_steps`` global time steps.  Between two exchanges the ghosts
        allow ``halo`` steps; when the stepper offers ``block(local_obs, k,
        bc_left, bc_right)`` (``engine_stepper`` does: one ``mt_rollout`` call,
        k back-to-back launches) each such run of steps is one call instead of
        k Python round trips."""
        fused = self._fused_advance(local_obs, int(n_steps))
        if fused is not None:
            return fused
        block = getattr(self.stepper, "block", None)
        left = int(n_steps)
        while left > 0:
            if self.needs_refresh():
                local_obs = self.refresh(local_obs)
            k = min(left, self.halo - self._since_exchange)
            bl = "halo" if self.has_left else "physical"
            br = "halo" if self.has_right else "physical"
            if block is not None:
                local_obs = block(local_obs, k, bl, br)
            else:
                for _ in range(k):
                    local_obs = self.stepper(local_obs, bl, br)
            self._since_exchange += k
            left -= k
        return local_obs

    def _fused_advance(self, local_obs, n_steps):
        """The whole run as ``mt_halo_advance`` launches (exchange folded into
        the step kernel), or None when that path does not apply."""
        engine = getattr(self.stepper, "engine", None)
        buffers = getattr(self.stepper, "buffers", None)
        if (self.peer is None or engine is None or buffers is None or n_steps < 1
                or not self.fuse_exchange or not _is_torch(local_obs)):
            return None
        esz = local_obs.element_size()
        if (self.halo * esz) % 16 or self.n_local % 4 or self.n_local <= 4096 // esz * 2:
            return None      # (mt_halo_advance would refuse: short block or odd halo)
        if not self._push_pending and 0 < self._since_exchange:
            local_obs = self.refresh(local_obs)      # partly used ghosts: start from fresh ones
        fresh = not self._push_pending
        x, other = buffers(local_obs,
                           "halo" if self.has_left else "physical",
                           "halo" if self.has_right else "physical")
        try:
            out = self.peer.advance(engine, x, other, n_steps, fresh)
        except ValueError:
            return None
        self._since_exchange = self.halo
        self._push_pending = True
        return out.reshape(local_obs.shape)

    def edge_cells(self, local_obs):
        """``(send_left, send_right)``: my first / last ``halo`` own cells as
        ``[rho(h) | v(h)]`` (None on a true road end)."""
        x = local_obs.reshape(-1)
        h, nl = self.halo, self.n_local
        a = h if self.has_left else 0            # first own cell
        b = nl - (h if self.has_right else 0)    # one past the last own cell
        cat = _cat(x)
        send_l = cat((x[a:a + h], x[nl + a:nl + a + h])) if self.has_left else None
        send_r = cat((x[b - h:b], x[nl + b - h:nl + b])) if self.has_right else None
        return send_l, send_r

    def set_ghosts(self, local_obs, recv_l, recv_r):
        """Write the neighbours' edge cells into my ghost cells (in place)."""
        x = local_obs.reshape(-1)
        h, nl = self.halo, self.n_local
        b = nl - (h if self.has_right else 0)
        if self.has_left:
            x[0:h] = recv_l[:h]
            x[nl:nl + h] = recv_l[h:]
        if self.has_right:
            x[b:b + h] = recv_r[:h]
            x[nl + b:nl + b + h] = recv_r[h:]
        self._since_exchange = 0
        return local_obs

    def refresh(self, local_obs):
        """Swap edge cells with the neighbours."""
        if self.peer is not None:      # peer-memory mailboxes: in place, on the GPU
            if self._push_pending:     # the fused path has already pushed this exchange
                self.peer.pull(local_obs, self.n_local)
                self._push_pending = False
            else:
                self.peer.exchange(local_obs, self.n_local)
            self._since_exchange = 0
            return local_obs
        send_l, send_r = self.edge_cells(local_obs)
        recv_l, recv_r = self.exchange(send_l, send_r)
        return self.set_ghosts(local_obs, recv_l, recv_r)

    def _torch_exchange(self, send_l, send_r):
        import torch
        import torch.distributed as dist
        probe = send_l if send_l is not None else send_r
        was_numpy = probe is not None and not _is_torch(probe)
        ops, recv_l, recv_r = [], None, None
        if send_l is not None:
            send_l = _as_tensor(send_l)
            recv_l = torch.empty_like(send_l)
            ops += [dist.P2POp(dist.isend, send_l, self.rank - 1),
                    dist.P2POp(dist.irecv, recv_l, self.rank - 1)]
        if send_r is not None:
            send_r = _as_tensor(send_r)
            recv_r = torch.empty_like(send_r)
            ops += [dist.P2POp(dist.isend, send_r, self.rank + 1),
                    dist.P2POp(dist.irecv, recv_r, self.rank + 1)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if was_numpy:
            recv_l = None if recv_l is None else recv_l.numpy()
            recv_r = None if recv_r is None else recv_r.numpy()
        return recv_l, recv_r

    def gather(self, local_obs):
        """All ranks' own cells assembled into the global ``[rho | v]`` on
        every rank (for checks and output, not on the hot path)."""
        import torch
        import torch.distributed as dist
        rho, v = self.own(local_obs)
        mine = _as_tensor(_cat(rho)((rho, v))).contiguous()
        sizes = [2 * (shard_range(self.n_cells, r, self.world)[1]
                      - shard_range(self.n_cells, r, self.world)[0])
                 for r in range(self.world)]
        parts = [torch.empty(s, dtype=mine.dtype, device=mine.device) for s in sizes]
        dist.all_gather(parts, mine) if len(set(sizes)) == 1 else \
            _all_gather_ragged(parts, mine, self.rank, self.world)
        rhos = [p[:p.numel() // 2] for p in parts]
        vs = [p[p.numel() // 2:] for p in parts]
        return torch.cat(rhos + vs)


class PeerHalo(object):
    """Ghost-cell exchange through peer-mapped mailboxes (``mt_halo_*``): the
    edge cells travel as plain NVLink stores of a small kernel, a second small
    kernel waits for the neighbours' flags and fills the ghosts -- no NCCL and
    no host round trip on the data path.  One object per rank; every rank must
    call ``exchange`` the same number of times.

    ``PeerHalo(halo, dtype, device).connect_group(rank, world)`` wires the
    ranks of a ``torch.distributed`` group together (the 64-byte IPC handles
    travel through ``all_gather_object``); ``PeerHalo.connect_local`` wires
    objects living in one process (tests)."""

    def __init__(self, halo, dtype, device=0):
        import ctypes
        from mbrl_traffic_b200 import _lib
        from mbrl_traffic_b200.engine import _DTYPES, dtype_name
        self._lib = _lib.load()
        self._check = _lib.check
        self.halo, self.device = int(halo), int(device)
        h = ctypes.c_void_p()
        self._check(self._lib.mt_halo_create(self.device, _DTYPES[dtype_name(dtype)], self.halo,
                                             ctypes.byref(h)))
        self._h = h

    def handle(self):
        import ctypes
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.mt_halo_handle(self._h, buf))
        return buf.raw

    def connect_group(self, rank, world, group=None):
        import torch.distributed as dist
        handles = [None] * world
        dist.all_gather_object(handles, self.handle(), group=group)
        left = handles[rank - 1] if rank > 0 else None
        right = handles[rank + 1] if rank < world - 1 else None
        self._check(self._lib.mt_halo_connect(self._h, left, right))
        return self

    @staticmethod
    def connect_local(halos):
        """Wire ``halos[r]`` (same process, rank order) to their neighbours."""
        for r, h in enumerate(halos):
            left = halos[r - 1]._h if r > 0 else None
            right = halos[r + 1]._h if r < len(halos) - 1 else None
            h._check(h._lib.mt_halo_connect_local(h._h, left, right))
        return halos

    def push(self, local_obs, n_local, stream=None):
        from mbrl_traffic_b200.engine import _ptr, _stream_ptr
        self._check(self._lib.mt_halo_push(self._h, _ptr(local_obs), int(n_local),
                                           _stream_ptr(stream)))

    def pull(self, local_obs, n_local, stream=None):
        from mbrl_traffic_b200.engine import _ptr, _stream_ptr
        self._check(self._lib.mt_halo_pull(self._h, _ptr(local_obs), int(n_local),
                                           _stream_ptr(stream)))

    def exchange(self, local_obs, n_local, stream=None):
        self.push(local_obs, n_local, stream)
        self.pull(local_obs, n_local, stream)

    def advance(self, engine, obs_a, obs_b, n_steps, ghosts_fresh, stream=None):
        """``mt_halo_advance``: ``n_steps`` steps of the block with the exchange
        folded into the step launches (one launch per ``halo`` steps).  Returns
        the buffer that holds the final state; the ghost cells are stale and
        one push is pending afterwards."""
        import ctypes
        from mbrl_traffic_b200.engine import _ptr, _stream_ptr
        flag = ctypes.c_int32(0)
        self._check(self._lib.mt_halo_advance(
            engine._h, self._h, _ptr(obs_a), _ptr(obs_b), int(n_steps),
            1 if ghosts_fresh else 0, _stream_ptr(stream), ctypes.byref(flag)))
        return obs_b if flag.value else obs_a

    def timed_out(self):
        import ctypes
        flag = ctypes.c_int32(0)
        self._check(self._lib.mt_halo_status(self._h, ctypes.byref(flag)))
        return bool(flag.value)

    def close(self):
        if self._h:
            self._lib.mt_halo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _all_gather_ragged(parts, mine, rank, world):
    import torch.distributed as dist
    for r in range(world):
        if r == rank:
            parts[r].copy_(mine)
        dist.broadcast(parts[r], src=r)


def _is_torch(x):
    return hasattr(x, "data_ptr")   # (NumPy 2 arrays also have `.device`)


def _cat(x):
    if _is_torch(x):
        import torch
        return torch.cat
    return np.concatenate


def _as_tensor(x):
    import torch
    if _is_torch(x):
        return x.contiguous()
    return torch.from_numpy(np.ascontiguousarray(x))


def engine_stepper(engine, physical_left, physical_right, left_val=(0.0, 0.0),
                   right_val=(0.0, 0.0), **params):
    """A ``stepper`` for ``DomainDecomposedRoad`` backed by one ``Engine``
    sized for the local block; ``physical_*`` are ``mt_bc`` codes of the true
    road ends (``_lib.MT_BC_EXTEND`` / ``MT_BC_CONSTANT``)."""
    import torch
    from mbrl_traffic_b200 import _lib
    state = {"bc": None, "bufs": None}

    def prepare(local_obs, bc_left, bc_right):
        cl = _lib.MT_BC_HALO if bc_left == "halo" else physical_left
        cr = _lib.MT_BC_HALO if bc_right == "halo" else physical_right
        if state["bc"] != (cl, cr):
            engine.set_params(
                boundary_conditions={"codes": (cl, cr, left_val, right_val)},
                **params)
            state["bc"] = (cl, cr)
        x = local_obs.reshape(1, -1)
        # the stepper ping-pongs between two buffers of its own; a block handed
        # in by the caller (the scattered initial state) is copied, never
        # overwritten
        if state["bufs"] is None:
            state["bufs"] = [torch.empty_like(x), torch.empty_like(x)]
        bufs = state["bufs"]
        for i in (0, 1):
            if bufs[i].data_ptr() == x.data_ptr():
                return bufs[i], bufs[1 - i]
        bufs[0].copy_(x)
        return bufs[0], bufs[1]

    def step(local_obs, bc_left, bc_right):
        x, out = prepare(local_obs, bc_left, bc_right)
        engine.step(x, out)
        return out.reshape(local_obs.shape)

    def block(local_obs, k, bc_left, bc_right):
        x, other = prepare(local_obs, bc_left, bc_right)
        final = engine.rollout(x, other, k)
        return final.reshape(local_obs.shape)

    step.block = block
    step.physical = (physical_left, physical_right)
    step.engine = engine
    step.buffers = prepare
    return step
