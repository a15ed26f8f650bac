"""Multi-GPU partition + gather for the sr / es / relax steps.

Replaces the mpi4py split of the reference (round-robin lateral points + pickled
bcast/gather at dbspm:505-515,540-556,585-591; the dead copy in pydbspm/parallel.py:4-64):

  sr / es  ranks own contiguous slabs of the window's z-planes; every rank holds the inputs
           (broadcast once) and computes its slab; rank 0 gathers and interleaves along z.
  relax    ranks own contiguous x-tiles of lateral positions; every rank holds the static
           window; rank 0 gathers (tiles are contiguous in C order: no permutation).

One process per GPU (torchrun); torch.distributed is the plumbing: NCCL on the GPUs,
gloo in the CPU unit tests.  There is no data-path collective other than the final gather
(the partitions are independent).  The compute callables are injected so the partition /
gather logic can be tested on CPU with world_size 2."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def split_range(lo: int, hi: int, parts: int):
    """Contiguous near-equal split of [lo,hi): part g = [lo + g*n//parts, lo + (g+1)*n//parts)."""
    n = hi - lo
    return [(lo + (g * n) // parts, lo + ((g + 1) * n) // parts) for g in range(parts)]


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for one process).
    Returns (rank, world, device)."""
    rank, world = env_rank_world()
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, device


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def barrier() -> None:
    """All ranks wait for each other (the reference's comm.Barrier() between steps); no-op for one process.  With NCCL
    the host blocks until the barrier collective has run on the device, i.e. after everything enqueued before it."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src)
    return t


def _gather_padded(local: torch.Tensor, sizes, dst: int = 0):
    """Gather tensors that differ only in their first dimension: pad to the largest, gather, trim."""
    rank, world = _world()
    if world == 1:
        return [local]
    big = max(sizes)
    pad = local
    if local.shape[0] < big:
        pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return [b[:n] for b, n in zip(bufs, sizes)]


def corr3d_sharded(compute_slab, nx, ny, z_lo, z_hi, device):
    """compute_slab(zl, zh) -> (nx,ny,zh-zl) tensor on `device`.  Rank 0 returns the full
    (nx,ny,z_hi-z_lo) window, other ranks None."""
    rank, world = _world()
    slabs = split_range(z_lo, z_hi, world)
    zl, zh = slabs[rank]
    if zh > zl:
        mine = compute_slab(zl, zh)
    else:
        mine = torch.empty((nx, ny, 0), dtype=torch.float64, device=device)
    if world == 1:
        return mine
    # z is the fastest axis: ship slab-major (nz_g, nx, ny), interleave on rank 0
    parts = _gather_padded(mine.permute(2, 0, 1).contiguous(), [b - a for a, b in slabs])
    if rank != 0:
        return None
    return torch.cat(parts, dim=0).permute(1, 2, 0).contiguous()


def interaction_plan(world: int, z_lo: int, z_hi: int, n_inter: int = 2):
    """Which (interaction, z-slab) jobs each rank runs when several independent interaction
    grids (sr, es) are wanted at once.

    The forward transforms of a correlation cannot be split by output plane (every output plane
    needs the whole spectrum), so z-slab sharding alone does not scale; sr and es, however, share
    no work at all.  With world >= n_inter the ranks are divided into n_inter groups, one per
    interaction, and each group splits the window's planes among its ranks.  With fewer ranks
    every rank runs the interactions one after the other on its slab of a plain z-slab split."""
    if world < n_inter:
        slabs = split_range(z_lo, z_hi, world)
        return [[(i, zl, zh) for i in range(n_inter)] for (zl, zh) in slabs]
    plan = []
    for i, (r_lo, r_hi) in enumerate(split_range(0, world, n_inter)):
        for zl, zh in split_range(z_lo, z_hi, r_hi - r_lo):
            plan.append([(i, zl, zh)])
    return plan


def gather_windows(plan, mine, nx, ny, z_lo, z_hi, device, n_inter: int = 2, dst: int = 0):
    """Assemble the n_inter full (nx,ny,z_hi-z_lo) windows on rank `dst` from the slabs each rank
    computed under `plan` (`mine`: this rank's slabs, in the order of its jobs).  One gather of
    slab-major buffers per job slot; other ranks get None."""
    rank, world = _world()
    if world == 1:
        return list(mine)
    nslots = max(len(jobs) for jobs in plan)
    nzw = z_hi - z_lo
    out = [torch.empty((nx, ny, nzw), dtype=torch.float64, device=device) for _ in range(n_inter)] if rank == dst else None
    for slot in range(nslots):
        jobs = [p[slot] if slot < len(p) else None for p in plan]
        sizes = [(j[2] - j[1]) * nx * ny if j else 0 for j in jobs]
        # slabs travel as flat buffers (z stays the fastest axis); rank dst drops each one into its
        # z-range of the full window with a single strided copy -- no transposes on either side
        local = mine[slot].reshape(-1) if slot < len(mine) else torch.empty(0, dtype=torch.float64, device=device)
        parts = _gather_padded(local, sizes, dst)
        if rank == dst:
            for j, part in zip(jobs, parts):
                if j and part.numel() > 0:
                    out[j[0]][:, :, j[1] - z_lo:j[2] - z_lo] = part.view(nx, ny, j[2] - j[1])
    return out


# ---- sr / es with the transform itself distributed (x-slab decomposition) ----------------------------------
_GROUPS: dict = {}


def dist_groups(world: int, n_inter: int = 2):
    """Rank groups for the distributed transform: n_inter equal groups of G = world / n_inter >= 2 consecutive
    ranks (group i computes interaction i), or None when the world does not divide that way (the callers then
    use interaction_plan: replicated inputs, z-slab split of the output)."""
    if world % n_inter or world // n_inter < 2:
        return None
    G = world // n_inter
    return [list(range(i * G, (i + 1) * G)) for i in range(n_inter)]


def _process_groups(groups):
    """torch.distributed groups for `groups` (created once; every rank creates every group, in order)."""
    key = tuple(tuple(g) for g in groups)
    if key not in _GROUPS:
        _GROUPS[key] = [dist.new_group(ranks=g) for g in groups]
    return _GROUPS[key]


def slot_order(n: int):
    """Frequencies of an axis of even length n in slot order 0, n/2, 1, n-1, 2, n-2, ... (dist_slot_of in
    csrc/corr_body.h): k and -k are neighbours, so equal even-sized blocks of slots are closed under k -> -k,
    which is what the fused z kernel needs of the rows a rank owns."""
    order = [0, n // 2]
    for k in range(1, n // 2):
        order += [k, n - k]
    return order[:n]


def x_slab(nx: int, G: int, h: int):
    return h * (nx // G), (h + 1) * (nx // G)


_SYMM: dict = {}


MAX_PEERS = 8          # DBSPM_MAX_PEERS of csrc/corr_body.h: pointer slots of the fused-exchange kernels


def _agreed(ok: bool, pg, device) -> bool:
    """True only if `ok` holds on EVERY rank of the group: the choice between the peer-store path and the NCCL
    exchange must be collective, or some ranks would sit in a symmetric-memory barrier while the others call
    all_to_all_single."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)
    return bool(int(flag.item()))


def peer_buffer(key, dims, pg, device):
    """A float64 buffer of shape `dims` in symmetric memory, mapped into every process of group `pg` (cached under
    `key`): (buffer, handle) or None when symmetric memory is unavailable on any rank of the group or the group has
    more ranks than the kernels have peer-pointer slots.  Collective over the group on first use."""
    if device.type != "cuda" or os.environ.get("DBSPM_NO_PEER") or dist.get_world_size(pg) > MAX_PEERS:
        return None
    key = (key, tuple(int(v) for v in dims), id(pg), device.index)
    if key not in _SYMM:
        got = None
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(tuple(int(v) for v in dims), dtype=torch.float64, device=device)
            hdl = symm.rendezvous(buf, pg)
            got = (buf, hdl)
        except Exception as exc:                                   # noqa: BLE001 -- any failure = no peer path
            print(f"dbspm_b200: symmetric memory unavailable ({exc!r}); exchanging through NCCL", flush=True)
        _SYMM[key] = got if _agreed(got is not None, pg, device) else None
    return _SYMM[key]


def peer_receive_buffer(shape, G: int, pg, device):
    """This rank's (2, nx, ny/G, nz) float64 receive buffer of the fused exchange, allocated in symmetric memory and
    mapped into every process of the group (torch.distributed._symmetric_memory): returns (buffer, handle) or None
    when symmetric memory is unavailable (CPU tests, no peer access) -- the callers then exchange through NCCL."""
    nx, ny, nz = (int(v) for v in shape)
    return peer_buffer("recv", (2, nx, ny // int(G), nz), pg, device)


def sr_es_distributed(stages, a_slab, b_slab, alpha, shape, dr, z_lo, z_hi, groups, dst: int = 0, on_stage=None,
                      peer: bool = True, tip_support=None, gather: bool = True):
    """The distributed sr / es step (replaces the rank-0-only dbspm:320-352).

    groups[i] = the G ranks that compute interaction i; this rank, member h of its group, passes the x-slab
    x_slab(nx, G, h) of its interaction's two arrays.  `stages` provides the three compute stages
    (dbspm_b200.interactions: dist_forward / dist_middle / dist_finish; the CPU tests inject numpy ones).
    Data-path exchanges: the transpose every distributed FFT needs (16N/G bytes per rank; fused into the y pass as
    peer stores when symmetric memory is available, else one all-to-all per array), a second, small all-to-all of the
    packed window rows back to x-slabs (so the final inverse y pass is distributed as well), and one gather of the
    finished real slabs into place on rank `dst`.  Returns the list of (nx,ny,nzw) windows on `dst`, None elsewhere.
    on_stage(name) is called after each stage has been enqueued (bench.py records a CUDA event there).
    tip_support = (s_lo, s_len): the caller's claim that the second array is exactly zero outside those z-planes
    (interactions.tip_zsupport); when pruning pays, the forward pass reads only the z-ranges the window can see and
    every later stage runs on the shorter z length (same result, see dbspm_corr3d_pruned_f64).
    gather=False: no gather -- every rank returns its finished real x-slab (nx/G, ny, nzw) of its interaction's window
    (the relax step that follows only needs the slab and a halo, exchange_halo)."""
    mark = on_stage or (lambda name: None)
    rank, world = _world()
    G = len(groups[0])
    cut = stages.pruned_cut(shape[2], z_lo, z_hi, tip_support) if (tip_support is not None and hasattr(stages, "pruned_cut")) else None
    kw = {}
    if cut is not None:
        # the pruned problem: z length Mp, window at [w_lo, w_lo + nzw)
        kw = {"cut": cut}
        shape = (int(shape[0]), int(shape[1]), int(cut[1]))
        z_lo, z_hi = int(cut[6]), int(cut[6]) + (int(z_hi) - int(z_lo))
    which = next(i for i, g in enumerate(groups) if rank in g)
    h = groups[which].index(rank)
    pg = _process_groups(groups)[which]
    symm = peer_receive_buffer(shape, G, pg, a_slab.device) if (peer and hasattr(stages, "dist_forward_peer")) else None
    if symm is not None:
        # compute + exchange in one kernel: the y pass stores each row into its owner's buffer over NVLink.
        # barrier 0: every rank is done reading its buffer (previous call); barrier 1: everyone's stores have landed
        buf, hdl = symm
        hdl.barrier(channel=0)
        stages.dist_forward_peer(a_slab, b_slab, G, h, hdl.buffer_ptrs, alpha, alpha,
                                 flags=int(os.environ.get("DBSPM_DIST_FLAGS", "0")), **kw)
        mark("fwd")
        hdl.barrier(channel=1)
        recvA, recvB = buf[0], buf[1]
        mark("a2a")
    else:
        sendA, sendB = stages.dist_forward(a_slab, b_slab, G, alpha, alpha, **kw)
        recvA, recvB = torch.empty_like(sendA), torch.empty_like(sendB)
        mark("fwd")
        dist.all_to_all_single(recvA, sendA, group=pg)
        dist.all_to_all_single(recvB, sendB, group=pg)
        mark("a2a")
    # (G, nxl, ny/G, nz) received = (nx, ny/G, nz): source ranks are consecutive x-slabs
    nx, ny = int(shape[0]), int(shape[1])
    nxl, nwh = nx // G, (int(z_hi) - int(z_lo) + 1) // 2
    symm2 = peer_buffer("w2", (G, nxl, ny // G, 2 * nwh), pg, a_slab.device) if (symm is not None and hasattr(stages, "dist_middle_peer")) else None
    if symm2 is not None:
        # second exchange fused into the inverse x pass of the middle stage (peer stores); the previous reader of
        # this buffer (the final pass of the previous call) finished before barrier 0 above
        W2, hdl2 = symm2
        stages.dist_middle_peer(recvA, recvB, dr, shape, G, h, z_lo, z_hi, hdl2.buffer_ptrs)
        mark("mid")
        hdl2.barrier(channel=0)
        slab = stages.dist_finish(W2, (nxl,) + tuple(int(v) for v in shape[1:]), G, z_lo, z_hi)
        mark("fin")
        return _gather_slabs(slab, groups, G, nx, nxl, dst, mark) if gather else slab
    W = stages.dist_middle(recvA, recvB, dr, shape, G, h, z_lo, z_hi)
    mark("mid")
    # second exchange (small: the packed window rows, 16 N_w / G bytes per rank): rank g receives the rows of x-slab g
    # for every y-frequency, laid out (source rank, nxl, ny/G, .) -- exactly what the row-mapped inverse y pass reads --
    # so the final pass is distributed too and rank `dst` only receives finished real slabs, straight into place
    W2 = torch.empty_like(W)
    dist.all_to_all_single(W2, W, group=pg)
    slab = stages.dist_finish(W2.view((G, nxl) + tuple(W.shape[1:])), (nxl,) + tuple(int(v) for v in shape[1:]), G, z_lo, z_hi)
    mark("fin")
    return _gather_slabs(slab, groups, G, nx, nxl, dst, mark) if gather else slab


def _gather_slabs(slab, groups, G, nx, nxl, dst, mark=None, async_op: bool = False):
    """Finished real x-slabs of every interaction -> the full windows on rank `dst` (received straight into place).
    async_op: returns (windows, work); the collective runs beside whatever is enqueued next (work.wait() joins it)."""
    rank, world = _world()
    wins = None
    gl = None
    if rank == dst:
        assert sorted(r for g in groups for r in g) == list(range(world)), "groups must cover the world"
        wins = [torch.empty((nx,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device) for _ in groups]
        gl = [None] * world
        for i, g in enumerate(groups):
            for hh, r in enumerate(g):
                gl[r] = wins[i][hh * nxl:(hh + 1) * nxl]
    work = dist.gather(slab, gl, dst=dst, async_op=async_op)
    if mark is not None:
        mark("gather")
    return (wins, work) if async_op else wins


def gather_x_slabs(slab, nx, dst: int = 0, async_op: bool = False):
    """One window from the x-slabs of ALL ranks (rank r holds planes [r*nx/W, (r+1)*nx/W)) onto rank `dst`."""
    rank, world = _world()
    if world == 1:
        return (slab, None) if async_op else slab
    out = _gather_slabs(slab, [list(range(world))], world, nx, nx // world, dst, None, async_op)
    if async_op:
        return (out[0][0] if out[0] is not None else None), out[1]
    return out[0] if out is not None else None


class PeerAssembler:
    """Rank `dst` assembles an array whose first axis is split evenly over the ranks, WITHOUT a collective kernel: dst's
    buffer lives in symmetric memory (mapped into every process), every rank copies its part straight into its rows there
    with a device-to-device copy -- for contiguous parts a `cudaMemcpyAsync` that runs on a copy engine over NVLink.  A copy
    engine needs no SM, so the transfer overlaps any kernel, including the persistent relax grid that holds every register of
    every SM (NCCL's gather kernels cannot start beside it: measured at N = 8, the window gathers either delayed the halo
    exchange by 1.0 ms or ran after the relax).

        asm = PeerAssembler.get("sr", (nx, ny, nzw), device)          # collective on first use; None -> use dist.gather
        asm.begin(stream)              # device barrier: dst is done with the previous contents
        asm.put(part, stream)          # this rank's rows, asynchronously on `stream`
        full = asm.end()               # current stream: device barrier = everyone's copy has landed; the array on dst, else None

    The two barriers are single-block kernels on the symmetric-memory signal pads; `end` is meant to be called when the SMs
    are free again (after the kernel the copies were hidden behind)."""

    _cache: dict = {}

    @classmethod
    def get(cls, key, shape, device, dst: int = 0):
        rank, world = _world()
        if world == 1 or int(shape[0]) % world:
            return None
        k = (key, tuple(int(v) for v in shape), device.index, int(dst))
        if k not in cls._cache:
            pg = _process_groups([list(range(world))])[0]          # the group of the distributed transform in slab mode
            symm = peer_buffer(("assemble", key, int(dst)), shape, pg, device)
            cls._cache[k] = cls(symm, shape, rank, world, dst) if symm is not None else None
        return cls._cache[k]

    def __init__(self, symm, shape, rank, world, dst):
        self.buf, self.hdl = symm
        self.shape = tuple(int(v) for v in shape)
        self.rank, self.world, self.dst = rank, world, int(dst)
        nl = self.shape[0] // world
        self.rows = (rank * nl, (rank + 1) * nl)
        self.remote = self.buf if rank == dst else self.hdl.get_buffer(self.dst, self.shape, torch.float64, 0)

    def begin(self, stream=None):
        with torch.cuda.stream(stream) if stream is not None else _null():
            self.hdl.barrier(channel=0)

    def put(self, part, stream=None):
        with torch.cuda.stream(stream) if stream is not None else _null():
            self.remote[self.rows[0]:self.rows[1]].copy_(part, non_blocking=True)

    def end(self):
        self.hdl.barrier(channel=1)
        return self.buf if self.rank == self.dst else None


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def exchange_halo(slab: torch.Tensor, halo: int) -> torch.Tensor:
    """Rank r holds planes [r*nxl, (r+1)*nxl) of a periodic array split evenly along its first axis; returns the
    (nxl + 2*halo, ...) array of its planes plus `halo` planes of each neighbour (periodic wrap): what the relax
    kernel's x-slab mode needs of the static potential (interactions.relax_grid(x_slab=...)).  Two point-to-point
    messages per rank (replaces the broadcast of the whole `static` to every rank, dbspm:540-556); when the halo is
    wider than a slab the slabs are all-gathered instead."""
    rank, world = _world()
    nxl = int(slab.shape[0])
    halo = int(halo)
    if halo == 0:
        return slab
    if world == 1 or halo > nxl:
        full = slab
        if world > 1:
            parts = [torch.empty_like(slab) for _ in range(world)]
            dist.all_gather(parts, slab.contiguous())
            full = torch.cat(parts, dim=0)
        n = int(full.shape[0])
        idx = (torch.arange(rank * nxl - halo, (rank + 1) * nxl + halo, device=full.device) % n)
        return full.index_select(0, idx)
    out = torch.empty((nxl + 2 * halo,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
    out[halo:halo + nxl] = slab
    left, right = (rank - 1) % world, (rank + 1) % world
    first, last = slab[:halo].contiguous(), slab[nxl - halo:].contiguous()
    # order matters when left == right (two ranks): the peer's first receive (its left halo) must meet my LAST planes
    ops = [dist.P2POp(dist.isend, last, right), dist.P2POp(dist.isend, first, left),
           dist.P2POp(dist.irecv, out[:halo], left), dist.P2POp(dist.irecv, out[halo + nxl:], right)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return out


def relax_sharded(compute_tile, nx, device):
    """compute_tile(i_lo, i_hi) -> dict of tensors whose first axis is the tile's x extent
    (E, theta, phi, tip_x, tip_y, tip_z: (ni,ny,nz)).  Rank 0 returns the concatenated dict."""
    rank, world = _world()
    tiles = split_range(0, nx, world)
    i_lo, i_hi = tiles[rank]
    mine = compute_tile(i_lo, i_hi)
    if world == 1:
        return mine
    out = {}
    sizes = [b - a for a, b in tiles]
    for key in sorted(mine):
        local = mine[key].contiguous()
        if len(set(sizes)) == 1:
            # equal tiles: rank 0 receives every tile straight into its rows of the result (no padded copies, no concatenation)
            gl = None
            if rank == 0:
                out[key] = torch.empty((nx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                gl = [out[key][a:b] for a, b in tiles]
            dist.gather(local, gl, dst=0)
        else:
            parts = _gather_padded(local, sizes)
            if rank == 0:
                out[key] = torch.cat(parts, dim=0)
    return out if rank == 0 else None


def sweep_assignment(param_sets, world: int):
    """Which parameter sets (indices into `param_sets`, each (alpha, V0, kappa)) a rank relaxes in a
    multi-GPU sweep.  Sets are dealt in contiguous blocks after sorting by alpha,
    so that a rank needs as few distinct sr grids (one forward correlation per alpha) as possible;
    block sizes differ by at most one."""
    order = sorted(range(len(param_sets)), key=lambda i: (float(param_sets[i][0]), i))
    return [[order[k] for k in range(lo, hi)] for lo, hi in split_range(0, len(order), world)]


def gather_sweep(results, assignment, keys=("E", "theta", "phi", "tip_x", "tip_y", "tip_z"), dst: int = 0):
    """results: this rank's relax dicts in the order of assignment[rank].  Rank `dst` returns the dicts
    of every set in the original order of the parameter list (tensors only for `keys`), others None."""
    rank, world = _world()
    if world == 1:
        out = [None] * sum(len(a) for a in assignment)
        for i, r in zip(assignment[0], results):
            out[i] = r
        return out
    shape = None
    for r in results:
        shape = tuple(r[keys[0]].shape)
    meta = [None] * world
    dist.all_gather_object(meta, shape)
    shape = next(m for m in meta if m is not None)
    dev = results[0][keys[0]].device if results else torch.device("cuda", torch.cuda.current_device()) \
        if torch.cuda.is_available() else torch.device("cpu")
    if results:
        mine = torch.stack([torch.stack([r[k] for k in keys]) for r in results])
    else:
        mine = torch.empty((0, len(keys)) + shape, dtype=torch.float64, device=dev)
    parts = _gather_padded(mine.contiguous(), [len(a) for a in assignment], dst=dst)
    if rank != dst:
        return None
    out = [None] * sum(len(a) for a in assignment)
    for a, part in zip(assignment, parts):
        for j, i in enumerate(a):
            out[i] = {k: part[j, q] for q, k in enumerate(keys)}
    return out
