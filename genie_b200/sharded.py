"""Grid sharding of the product-graph front end over the GPUs of one node (BASELINE.json configs[4]).

The reference runs on one device (`module.py:2`); it has no counterpart of this file.  For networks whose node features
exceed one GPU (2000 stations x 200000 grid nodes = 4e8 product nodes) the grid nodes are dealt to the ranks:

* `GridPartition` — every rank computes the same partition on the host: the compact grid-node groups of
  `plan.bisection_groups` (depth-first order, spatially coherent) are assigned to ranks in contiguous runs of about equal
  node counts.  A rank's local grid = its owned nodes followed by the 1-hop halo of their source-graph in-neighbours.
* `ShardedFrontEnd.forward` — per window:
    1. layer 0 + 1 of DataAggregation on the local grid (inputs of halo nodes are computed locally: Slice/Mask are cheap,
       so layer 0 needs no exchange; layer 1 is only computed for owned nodes),
    2. ONE exchange: the layer-2 message rows of halo nodes are fetched from their owners (`all_to_all_single`, NCCL over
       NVLink; S x 64 B per halo grid node),
    3. layer 2 + Bipartite_ReadIn for the owned nodes (a grid node's stations all live on its rank, so the sum over stations
       never crosses shards: no all-reduce of station partials is needed in this layout),
    4. all-gather of the `[G,15]` read-in rows; SpatialAggregation x3 (0.1 % of the work) runs replicated on every rank.
  With `PeerHalo` (the product path on several GPUs) step 2 is no collective at all: every rank's layer-1 station pass stores
  the rows its peers need straight into the peers' landing buffers while it produces them (peer stores over NVLink,
  genie_plan_set_halo_export), and a one-element all-reduce orders "all layer-1 passes done" before layer 2.
The compute steps go through a backend object; the product backend is `CudaBackend` (C-ABI kernels).  Tests substitute a CPU
backend to check the partition / exchange logic with world_size 2 over gloo.
"""
import numpy as np
import torch
import torch.distributed as dist

from .plan import GROUP_SIZE, bisection_groups, csr_by_destination


class GridPartition(object):
    """Deterministic assignment of grid nodes to ranks + the halo / exchange index lists of every rank (host side)."""

    def __init__(self, A_src_src, n_grid, world, group_size=GROUP_SIZE):
        self.n_grid, self.world = int(n_grid), int(world)
        rowptr, col = csr_by_destination(A_src_src.cpu(), n_grid)
        self.rowptr, self.col = rowptr.numpy(), col.numpy().astype(np.int64)
        gp, gn = bisection_groups(self.rowptr, self.col, n_grid, group_size)
        # contiguous runs of groups with ~ n_grid / world nodes each
        bounds = np.searchsorted(gp, np.arange(1, world) * (n_grid / float(world)), side='left')
        bounds = np.concatenate(([0], np.clip(bounds, 0, len(gp) - 1), [len(gp) - 1]))
        self.owned = [np.ascontiguousarray(gn[gp[bounds[r]]:gp[bounds[r + 1]]].astype(np.int64)) for r in range(world)]
        # the groups a rank owns, as ranges of its local ids (its owned nodes are those groups' nodes, in group order)
        self.group_ptr = [np.ascontiguousarray(gp[bounds[r]:bounds[r + 1] + 1] - gp[bounds[r]]).astype(np.int32)
                          for r in range(world)]
        self.owner = np.empty(n_grid, dtype=np.int64)
        for r in range(world):
            self.owner[self.owned[r]] = r
        self.halo = []
        for r in range(world):
            nb = np.unique(np.concatenate([self.col[self.rowptr[g]:self.rowptr[g + 1]] for g in self.owned[r]]
                                          or [np.zeros(0, np.int64)]))
            nb = nb[self.owner[nb] != r]
            self.halo.append(nb[np.lexsort((nb, self.owner[nb]))])             # grouped by owner, ascending id inside

    def local_nodes(self, rank):
        """Global ids of the local grid of `rank`: owned nodes first, then the halo."""
        return np.concatenate((self.owned[rank], self.halo[rank]))

    def local_groups(self, rank):
        """(grp_ptr, grp_nodes) of the local plan: the compact groups of the WHOLE grid's bisection that `rank` owns.  (A
        bisection of the local graph, whose halo nodes have no in-edges, gives ragged groups: the layer-1 source pass of the
        eight ranks of C5 took 4.1 - 6.7 ms with it against 3.9 ms for the same node count on one GPU.)"""
        gp = self.group_ptr[rank]
        keep = np.concatenate(([True], np.diff(gp) > 0))
        return gp[keep], np.arange(len(self.owned[rank]), dtype=np.int32)

    def local_graph(self, rank):
        """Source graph restricted to the local grid (local ids): in-edges of owned nodes only; halo rows are empty."""
        nodes = self.local_nodes(rank)
        g2l = -np.ones(self.n_grid, dtype=np.int64)
        g2l[nodes] = np.arange(len(nodes))
        own = self.owned[rank]
        deg = self.rowptr[own + 1] - self.rowptr[own]
        tgt = np.repeat(np.arange(len(own)), deg)
        src = g2l[np.concatenate([self.col[self.rowptr[g]:self.rowptr[g + 1]] for g in own] or [np.zeros(0, np.int64)])]
        assert (src >= 0).all()
        return torch.from_numpy(np.stack((src, tgt), axis=0)).long()

    def exchange_lists(self, rank):
        """(send_rows, send_counts, recv_counts): local row ids (owned part) this rank sends, ordered by destination rank
        and, per destination, in the order of that rank's halo list; counts per peer."""
        g2l = -np.ones(self.n_grid, dtype=np.int64)
        g2l[self.owned[rank]] = np.arange(len(self.owned[rank]))
        send, send_counts = [], []
        for q in range(self.world):
            want = self.halo[q][self.owner[self.halo[q]] == rank] if q != rank else np.zeros(0, np.int64)
            send.append(g2l[want])
            send_counts.append(len(want))
        recv_counts = [int((self.owner[self.halo[rank]] == q).sum()) for q in range(self.world)]
        return np.concatenate(send), send_counts, recv_counts


class PeerHalo(object):
    """Halo rows over peer memory: landing buffers (genie_peer_alloc), their handles exchanged over the process group, the
    peers' buffers mapped (genie_peer_open) and the export tables of this rank installed on its plan.  Works for ranks on
    different GPUs of one node and for ranks sharing a GPU (the tests)."""

    def __init__(self, partition, rank, plan, n_sta, device, group=None):
        import ctypes
        from . import capi
        lib = capi.load()
        self.lib, self.device, self.group = lib, torch.device(device), group
        world = partition.world
        n_halo = len(partition.halo[rank])
        row_bytes = n_sta * 16 * (2 if plan.storage == 'bf16' else 4)
        self.bytes = max(n_halo, 1) * row_bytes
        # Every rank goes through the same collectives whatever happens locally, and all ranks agree on the outcome: a rank that
        # cannot allocate, export or map a buffer (no peer access between two GPUs, inter-process handles disabled) makes the
        # whole group fall back — `PeerHalo.create` then returns None and the caller keeps the all-to-all.
        self.local_ptr, self.peer_ptrs, err = None, [], None
        with torch.cuda.device(self.device):
            ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(capi.PEER_HANDLE_BYTES)
            try:
                capi.check(lib.genie_peer_alloc(self.bytes, ctypes.byref(ptr), handle))
                self.local_ptr = ptr.value
            except capi.GenieError as e:
                err = str(e)
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw if err is None else None, group=group)
            for q in range(world):
                if q == rank:
                    self.peer_ptrs.append(self.local_ptr)
                    continue
                pq = ctypes.c_void_p()
                try:
                    if handles[q] is None:
                        raise capi.GenieError('rank %d has no landing buffer' % q)
                    capi.check(lib.genie_peer_open(handles[q], ctypes.byref(pq)))
                except capi.GenieError as e:
                    err = err or str(e)
                self.peer_ptrs.append(pq.value)
            errs = [None] * world
            dist.all_gather_object(errs, err, group=group)
        self.plan, self.rank = None, rank
        if any(e is not None for e in errs):
            self._release()
            raise capi.GenieError('halo rows over peer memory are not available: %s' % next(e for e in errs if e is not None))
        ptr_, peer, row = self.export_tables(partition, rank)
        dev = self.device
        self.exp_ptr = torch.from_numpy(ptr_).to(dev)
        pad = lambda v: v if len(v) else np.zeros(1, dtype=np.int32)       # (a rank that exports nothing still passes tables)
        self.exp_peer = torch.from_numpy(pad(peer)).to(dev)
        self.exp_row = torch.from_numpy(pad(row)).to(dev)
        self.peer_base = torch.tensor(self.peer_ptrs, dtype=torch.int64, device=dev)
        self.export_bytes = int(len(row)) * row_bytes                # bytes this rank stores into peer memory per window
        plan.set_halo_export(self.exp_ptr, self.exp_peer, self.exp_row, self.peer_base, self.local_ptr)
        self.plan, self.rank = plan, rank
        self._flag = torch.zeros(1, device=dev)

    @staticmethod
    def export_tables(partition, rank):
        """Export CSR over the owned nodes of `rank` (int32 arrays): exp_ptr [n_owned + 1], and per export the peer rank and
        the position of the node in that peer's halo list (= row of its landing buffer)."""
        own = partition.owned[rank]
        g2l = -np.ones(partition.n_grid, dtype=np.int64)
        g2l[own] = np.arange(len(own))
        loc, peer, row = [], [], []
        for q in range(partition.world):
            if q == rank:
                continue
            pos = np.nonzero(partition.owner[partition.halo[q]] == rank)[0]
            loc.append(g2l[partition.halo[q][pos]])
            peer.append(np.full(len(pos), q, dtype=np.int64))
            row.append(pos)
        loc, peer, row = (np.concatenate(a) if a else np.zeros(0, np.int64) for a in (loc, peer, row))
        order = np.argsort(loc, kind='stable')
        counts = np.zeros(len(own) + 1, dtype=np.int64)
        np.add.at(counts, loc + 1, 1)
        return np.cumsum(counts).astype(np.int32), peer[order].astype(np.int32), row[order].astype(np.int32)

    def fence(self):
        """Stream-ordered: returns (on the stream) once every rank's work enqueued before its own fence has completed."""
        dist.all_reduce(self._flag, group=self.group)

    @classmethod
    def create(cls, partition, rank, plan, n_sta, device, group=None):
        """A PeerHalo, or None (on every rank alike) when peer memory cannot be set up on this node."""
        from . import capi
        try:
            return cls(partition, rank, plan, n_sta, device, group)
        except capi.GenieError:
            return None

    def _release(self):
        with torch.cuda.device(self.device):
            for q, p in enumerate(self.peer_ptrs):
                if q != self.rank and p:
                    self.lib.genie_peer_close(p)
            if self.local_ptr:
                self.lib.genie_peer_free(self.local_ptr)
        self.peer_ptrs, self.local_ptr = [], None

    def close(self):
        if self.plan is not None:
            self.plan.set_halo_export(None, None, None, None, None)
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                dist.barrier(group=self.group)                      # nobody still stores into a buffer that is about to go
            self._release()
            self.plan = None


class CudaBackend(object):
    """The product compute path: libgenie_b200 kernels on the local plan (see include/genie_b200.h)."""

    def __init__(self, model, A_sta_sta, A_src_local, n_sta, n_local, n_owned, read_in_attr_local, A_src_global, n_grid,
                 device, grid_groups=None):
        """grid_groups: GridPartition.local_groups(rank) — the source-pass groups of the local plan."""
        from .plan import GraphPlan
        self.model, self.device = model, torch.device(device)
        self.plan = GraphPlan.cartesian(A_sta_sta, A_src_local, n_sta, n_local, device=device, n_grid_owned=n_owned,
                                        grid_groups=grid_groups)
        self.plan_grid = GraphPlan.grid_only(A_src_global, n_grid, device)
        self.attr = read_in_attr_local.to(device).float().contiguous()
        self._mask = None

    def layer1(self, Slice, Mask):
        from . import ops
        self._mask = Mask
        ops.da_layer1_fwd(self.plan, self.model._packed_weights(self.device), Slice, Mask)

    def message_rows(self):
        from . import ops
        return ops.message_rows(self.plan)

    def layer2_readin(self):
        from . import ops
        return ops.da_layer2_readin_fwd(self.plan, self.model._packed_weights(self.device), self._mask, self.attr)[0]

    def spatial(self, read_in, pos, scale_rel):
        from . import ops
        packed = self.model._packed_weights(self.device)
        x = read_in
        for layer in range(3):
            x = ops.spatial_aggregation_fwd(self.plan_grid, packed, layer, x, pos, scale_rel)
        return x


def sharded_heads(model, x_spatial, x_grid_cart, x_query_cart, t_query, rank, world, group=None):
    """The read-out heads of forward_fixed_source (module.py:1015-1020) split by rows over the ranks: rank r computes y for its
    block of grid nodes and x for its block of query points from the (replicated) x_spatial, two all-gathers assemble
    (y [G,T,1], x [Q,T,1]) on every rank.  Needs the kernel-supported head shapes (ops.HeadsWeights.supported)."""
    from . import capi, ops
    dev = x_spatial.device
    if not ops.HeadsWeights.supported(model):
        raise capi.GenieError('sharded_heads: head shapes without a kernel (use model._heads on one rank)')
    if model._heads_w is None or model._heads_w.device != dev:
        model._heads_w = ops.HeadsWeights(dev)
    hp, fold, T = model._heads_w.update(model, t_query)
    G, Q = x_spatial.shape[0], x_query_cart.shape[0]
    gs, qs = -(-G // world), -(-Q // world)
    g0, g1 = min(G, rank * gs), min(G, (rank + 1) * gs)
    q0, q1 = min(Q, rank * qs), min(Q, (rank + 1) * qs)
    nbr = model.SpatialAttention._nbr_table(model.SpatialAttention._edges(x_query_cart, x_grid_cart, 10), Q)
    y, x = ops.heads_fwd(model._heads_w, hp, fold, T, x_spatial, x_grid_cart, x_query_cart, nbr,
                         float(model.SpatialAttention.scale_rel), grid_rows=(g0, g1), query_rows=(q0, q1))
    out = []
    for part, n, blk in ((y, G, gs), (x, Q, qs)):
        pad = torch.zeros((blk, T), dtype=part.dtype, device=dev)
        pad[:part.shape[0]] = part.view(-1, T)
        full = torch.empty((world * blk, T), dtype=part.dtype, device=dev)
        dist.all_gather_into_tensor(full, pad, group=group)
        out.append(full[:n].unsqueeze(-1))
    return out[0], out[1]


class ShardedFrontEnd(object):
    """DataAggregation -> Bipartite_ReadIn -> SpatialAggregation x3 with the grid nodes sharded over the process group."""

    def __init__(self, partition, rank, backend, device, group=None, peer_halo=None):
        """peer_halo: a PeerHalo installed on the backend's plan — the halo rows then travel inside the layer-1 kernel and
        `forward` only fences; None: one all_to_all_single per window."""
        self.part, self.rank, self.backend, self.group = partition, int(rank), backend, group
        self.peer_halo = peer_halo
        self.device = torch.device(device)
        self.n_owned = len(partition.owned[rank])
        send_rows, self.send_counts, self.recv_counts = partition.exchange_lists(rank)
        self.send_rows = torch.from_numpy(send_rows).to(self.device)
        self.max_owned = max(len(o) for o in partition.owned)
        order = np.concatenate([np.concatenate((o, -np.ones(self.max_owned - len(o), dtype=np.int64)))
                                for o in partition.owned])
        keep = order >= 0
        inv = np.empty(partition.n_grid, dtype=np.int64)
        inv[order[keep]] = np.nonzero(keep)[0]
        self.gather_index = torch.from_numpy(inv).to(self.device)          # global node -> row of the padded all-gather
        self._pad, self._allr, self.exchange_bytes = None, None, 0

    def exchange(self, rows):
        """rows [n_local, W]: fills the halo part (rows >= n_owned) from the owners' copies.  The halo rows of a rank are
        ordered by owner, so the all-to-all receives straight into the workspace; only the send side needs a gather (a
        boundary row can be wanted by several peers)."""
        send = rows.index_select(0, self.send_rows)
        recv = rows[self.n_owned:]
        if recv.shape[0] != sum(self.recv_counts):
            raise RuntimeError('halo size mismatch')
        dist.all_to_all_single(recv, send, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts,
                               group=self.group)
        return send.numel() * send.element_size()

    def forward(self, Slice_local, Mask_local, pos_grid, scale_rel, events=None):
        """Slice/Mask of the LOCAL product nodes ([n_local * S, 4], owned grid nodes first) -> x_spatial [G,30] (replicated).
        `events`: optional list that receives (begin, end) CUDA events bracketing the halo exchange."""
        be = self.backend
        be.layer1(Slice_local, Mask_local)
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.peer_halo is not None:
            self.peer_halo.fence()
            self.exchange_bytes = self.peer_halo.export_bytes
        else:
            self.exchange_bytes = self.exchange(be.message_rows())
        if events is not None:
            e1.record()
            events.append((e0, e1))
        r_own = be.layer2_readin()                                           # [n_owned, 15]
        if self._pad is None or self._pad.device != r_own.device:
            self._pad = torch.zeros((self.max_owned, r_own.shape[1]), dtype=r_own.dtype, device=r_own.device)
            self._allr = torch.empty((self.part.world * self.max_owned, r_own.shape[1]), dtype=r_own.dtype, device=r_own.device)
        self._pad[:self.n_owned] = r_own
        dist.all_gather_into_tensor(self._allr, self._pad, group=self.group)
        read_in = self._allr.index_select(0, self.gather_index)              # [G,15] in global node order
        return be.spatial(read_in, pos_grid, scale_rel), read_in
