"""Graph plans: the compact, destination-sorted form of the reference's edge lists that the CUDA kernels consume.

The reference passes explicit int64 `[2, E]` edge lists (row 0 = message source j, row 1 = target i;
process_utils.py:718-722) to `GCN_Detection_Network_extended.set_adjacencies` (module.py:941).  In its default dense
mode those product-graph lists are pure index patterns over the Cartesian product of the station kNN graph and the grid
kNN graph; `GraphPlan.from_edge_lists` recognises the pattern and keeps only the two small graphs (the product edges
stay implicit — at 1000 stations x 50000 grid nodes the explicit lists would be 24 GB).  Any other product graph
(sub-graph mode, process_utils.py:744-849) is kept as an explicit CSR.
"""
import ctypes

import os

import numpy as np
import torch

from . import capi


def csr_by_destination(edge_index, n_nodes):
    """[2,E] (source, target) -> (rowptr int64 [n+1], col int32 [E]) sorted by target, source order kept."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    if src.numel() and (int(dst.max()) >= n_nodes or int(dst.min()) < 0 or int(src.min()) < 0):
        raise ValueError('edge target out of range')
    order = torch.sort(dst, stable=True)[1]
    col = src[order].to(torch.int32).contiguous()
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=edge_index.device)
    if src.numel():
        rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_nodes), 0)
    return rowptr.contiguous(), col


def locality_order(rowptr, col, n):
    """A permutation of 0..n-1 in which graph neighbours are close together: reverse Cuthill-McKee of the symmetrised
    graph (host side, once per plan).  The kernels only use it to ORDER their tiles so that the neighbour tiles a CTA
    fetches were fetched recently by another CTA and still sit in L2; results do not depend on it."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    rp = rowptr.cpu().numpy().astype(np.int64)
    cl = col.cpu().numpy().astype(np.int32)
    if n == 0 or len(cl) == 0:
        return np.arange(n, dtype=np.int32)
    a = csr_matrix((np.ones(len(cl), dtype=np.int8), cl, rp), shape=(n, n))
    a = (a + a.T).tocsr()
    return np.ascontiguousarray(reverse_cuthill_mckee(a, symmetric_mode=True), dtype=np.int32)


# ---- on-chip tiling tables of the split (source pass / station pass) kernels ----------------------------------------------
TILE_M = 128        # product-node rows of one station tile (= MMA M)
ROWS_MAX = 288      # station rows (tile + halo) staged in shared memory per tile; row ROWS_MAX is an all-zero row
GROUP_SIZE = 256    # grid nodes per source-pass group (measured best on B200 together with 8-station slabs)


def bisection_groups(rowptr, col, n, size):
    """Partitions the nodes of a graph into compact groups of <= `size` nodes by recursive bisection along reverse
    Cuthill-McKee orders of the (symmetrised) sub-graphs: no coordinates needed, and the union of the neighbour sets of a
    group stays small (the on-chip reuse factor of the kernels).  Returns (ptr int32 [ng+1], nodes int32 [n]); groups
    are emitted in depth-first order, so consecutive groups are close in the graph."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    rp = np.asarray(rowptr.cpu() if torch.is_tensor(rowptr) else rowptr).astype(np.int64)
    cl = np.asarray(col.cpu() if torch.is_tensor(col) else col).astype(np.int32)
    if n == 0:
        return np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32)
    a = csr_matrix((np.ones(len(cl), dtype=np.int8), cl, rp), shape=(n, n))
    a = (a + a.T).tocsr()
    groups = []
    stack = [np.arange(n, dtype=np.int64)]
    while stack:
        idx = stack.pop()
        m = len(idx)
        if m <= size:
            groups.append(idx)
            continue
        o = reverse_cuthill_mckee(a[idx][:, idx], symmetric_mode=True)
        nl = (m // 2 + size - 1) // size * size
        if nl >= m:
            nl = m // 2
        stack.append(idx[o[nl:]])
        stack.append(idx[o[:nl]])           # left half is processed first (depth-first order)
    ptr = np.zeros(len(groups) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(x) for x in groups])
    return ptr, np.concatenate(groups).astype(np.int32)


def station_tiles(rowptr, col, n_sta, tile_m=TILE_M, rows_max=ROWS_MAX):
    """Station tiles of the station-pass kernels: every tile is a compact set of <= tile_m stations plus the halo of
    their in-neighbours, <= rows_max rows in all.  Returns None when no tile size >= 32 fits, else a dict of numpy arrays
      rows   int32  [NT, rows_max]   station id of each staged row (tile stations first, then the halo; padding 0)
      meta   int32  [NT, 2]          (stations in the tile, staged rows)
      nbr    uint16 [NT, tile_m, 16] staged-row index of every in-neighbour (padding: rows_max = the zero row)
      invdeg fp32   [NT, tile_m]     1 / in-degree (0 for isolated stations: PyG's mean of nothing is 0)."""
    rp = np.asarray(rowptr.cpu() if torch.is_tensor(rowptr) else rowptr).astype(np.int64)
    cl = np.asarray(col.cpu() if torch.is_tensor(col) else col).astype(np.int64)
    deg = rp[1:] - rp[:-1]
    if n_sta == 0 or (len(deg) and deg.max() > 16):
        return None
    for t in range(tile_m, 31, -16):
        ptr, nodes = bisection_groups(rp, cl, n_sta, t)
        nt = len(ptr) - 1
        rows = np.zeros((nt, rows_max), dtype=np.int32)
        meta = np.zeros((nt, 2), dtype=np.int32)
        nbr = np.full((nt, tile_m, 16), rows_max, dtype=np.uint16)
        invdeg = np.zeros((nt, tile_m), dtype=np.float32)
        ok = True
        for i in range(nt):
            own = nodes[ptr[i]:ptr[i + 1]].astype(np.int64)
            local = {int(s): r for r, s in enumerate(own)}
            lst = list(own)
            for r, s in enumerate(own):
                js = cl[rp[s]:rp[s + 1]]
                for q, j in enumerate(js):
                    j = int(j)
                    if j not in local:
                        local[j] = len(lst)
                        lst.append(j)
                    nbr[i, r, q] = local[j]
                if len(js):
                    invdeg[i, r] = np.float32(1.0) / np.float32(len(js))
            if len(lst) > rows_max:
                ok = False
                break
            rows[i, :len(lst)] = lst
            meta[i] = (len(own), len(lst))
        if ok:
            _pair_neighbour_order(nbr, rows_max)
            return dict(rows=rows, meta=meta, nbr=nbr, invdeg=invdeg, tile=t)
    return None


def _pair_neighbour_order(nbr, pad):
    """Orders every row's neighbour list so that the layer-2 gather of 64-byte rows is (mostly) free of shared-memory bank
    conflicts.  A 64-byte staged row covers half of the 32 banks — which half is the parity of its staged index — and the lanes
    r and r + 4 of a quarter-warp read the same 16-byte chunk position at every step (da_s2_kernel.cu), so they collide exactly
    when their step-j neighbours have the same parity.  For every such pair the two lists are re-ordered (a sum does not care)
    so that at as many steps as possible one lane reads an even and the other an odd staged row; padding goes last."""
    nt, tm, k = nbr.shape
    for t in range(nt):
        for r in range(tm):
            if r & 4:
                continue
            s = r + 4
            a = [int(x) for x in nbr[t, r] if x != pad]
            b = [int(x) for x in nbr[t, s] if x != pad] if s < tm else []
            ea, oa = [x for x in a if x % 2 == 0], [x for x in a if x % 2]
            eb, ob = [x for x in b if x % 2 == 0], [x for x in b if x % 2]
            n1 = min(len(ea), len(ob))                      # steps with (even, odd)
            n2 = min(len(oa), len(eb))                      # steps with (odd, even)
            new_a = ea[:n1] + oa[:n2] + ea[n1:] + oa[n2:]
            new_b = ob[:n1] + eb[:n2] + ob[n1:] + eb[n2:]
            nbr[t, r, :] = pad
            nbr[t, r, :len(new_a)] = new_a
            if s < tm:
                nbr[t, s, :] = pad
                nbr[t, s, :len(new_b)] = new_b


def _is_cartesian(A_in_sta, A_in_src, prod_target, S, G):
    """Checks the index patterns of process_utils.py:720-722; returns (A_sta_sta, A_src_src) or None."""
    P = S * G
    dev = A_in_sta.device
    if prod_target is not None:
        if prod_target.numel() != P:
            return None
        if not torch.equal(prod_target.long(), torch.arange(G, device=dev).repeat_interleave(S)):
            return None
    Es, Eg = A_in_sta.shape[1], A_in_src.shape[1]
    if G == 0 or S == 0 or Es % G or Eg % S:
        return None
    es, eg = Es // G, Eg // S
    base_sta = A_in_sta[:, :es].long()
    if es and (int(base_sta.max()) >= S):
        return None
    # A_prod_sta_sta = A_sta_sta.repeat(1, G) + S * g          (:720)
    chunk = max(1, (1 << 24) // max(es, 1))
    for g0 in range(0, G, chunk):
        g1 = min(G, g0 + chunk)
        want = base_sta.unsqueeze(1) + S * torch.arange(g0, g1, device=dev).view(1, -1, 1)
        if not torch.equal(A_in_sta[:, g0 * es:g1 * es].long().view(2, g1 - g0, es), want):
            return None
    # A_prod_src_src = S * A_src_src.repeat(1, S) + s           (:721)
    first = A_in_src[:, :eg].long()
    if eg and bool((first % S != 0).any()):
        return None
    base_src = first // S
    chunk = max(1, (1 << 24) // max(eg, 1))
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        want = S * base_src.unsqueeze(1) + torch.arange(s0, s1, device=dev).view(1, -1, 1)
        if not torch.equal(A_in_src[:, s0 * eg:s1 * eg].long().view(2, s1 - s0, eg), want):
            return None
    return base_sta, base_src


class GraphPlan(object):
    """Owns the device index arrays and the C-side plan handle (genie_plan_t)."""

    def __init__(self, mode, n_sta, n_grid, n_prod, sta, src, grid, grid_outdeg, prod_grid, device, grid_order=None,
                 tiling=True, group_size=GROUP_SIZE, n_grid_owned=0, grid_groups=None):
        self.mode, self.n_sta, self.n_grid, self.n_prod = mode, int(n_sta), int(n_grid), int(n_prod)
        self.n_grid_owned = int(n_grid_owned)          # grid sharding: nodes >= n_grid_owned are halo copies (0 = none)
        self.device = torch.device(device)
        self.sta_max_deg, self.grid_order = 0, None
        if mode == capi.GRAPH_CARTESIAN and n_sta > 0 and n_grid > 0:
            deg = sta[0][1:] - sta[0][:-1]
            self.sta_max_deg = int(deg.max()) if deg.numel() else 0
            if grid_order is None:
                grid_order = locality_order(src[0], src[1], n_grid)
            grid_order = np.ascontiguousarray(grid_order, dtype=np.int32)
            if not np.array_equal(np.sort(grid_order), np.arange(n_grid)):
                raise ValueError('grid_order must be a permutation of the grid nodes')
            self.grid_order = torch.from_numpy(grid_order).to(self.device).contiguous()
        # on-chip tiling tables of the split kernels (dense mode; None -> the one-pass kernels are used)
        self.tiles = None
        if mode == capi.GRAPH_CARTESIAN and n_sta >= 2 and n_grid > 0 and 1 <= self.sta_max_deg <= 16 and tiling:
            st = station_tiles(sta[0], sta[1], self.n_sta)
            if st is not None:
                if grid_groups is not None:                 # the caller's groups (grid sharding: the groups of the WHOLE grid's
                    gp, gn = (np.ascontiguousarray(a, dtype=np.int32) for a in grid_groups)      # bisection that a rank owns)
                    if gp[0] != 0 or gp[-1] != len(gn) or (np.diff(gp) <= 0).any() or \
                            not np.array_equal(np.sort(gn), np.arange(self.n_grid_owned or self.n_grid)):
                        raise ValueError('grid_groups must partition the (owned) grid nodes into non-empty groups')
                else:
                    gp, gn = bisection_groups(src[0], src[1], self.n_grid, int(group_size))
                if self.n_grid_owned and grid_groups is None:   # halo nodes are gathered FROM, never computed: drop them
                    keep = gn < self.n_grid_owned
                    cnt = np.add.reduceat(keep.astype(np.int64), gp[:-1]) if len(gp) > 1 else np.zeros(0, np.int64)
                    cnt = cnt[cnt > 0]
                    gn = gn[keep]
                    gp = np.concatenate(([0], np.cumsum(cnt))).astype(np.int32)
                put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
                self.tiles = dict(n_tiles=int(st['meta'].shape[0]), n_groups=int(len(gp) - 1), tile=int(st['tile']),
                                  rows=put(st['rows']), meta=put(st['meta']),
                                  nbr=put(st['nbr'].view(np.int16)), invdeg=put(st['invdeg']),
                                  grp_ptr=put(gp), grp_nodes=put(gn))
        self._keep = (sta, src, grid, grid_outdeg, prod_grid, self.grid_order, self.tiles)     # keep the tensors alive
        self.sta_rowptr, self.sta_col = sta
        self.src_rowptr, self.src_col = src
        self.grid_rowptr, self.grid_col = grid
        self.grid_outdeg = grid_outdeg
        self.prod_grid = prod_grid
        d = capi.GraphDesc()
        d.mode, d.n_sta, d.n_grid, d.sta_max_deg, d.n_prod = mode, self.n_sta, self.n_grid, self.sta_max_deg, self.n_prod
        d.sta_rowptr = capi.dptr(self.sta_rowptr, torch.int64, 'sta_rowptr')
        d.sta_col = capi.dptr(self.sta_col, torch.int32, 'sta_col')
        d.src_rowptr = capi.dptr(self.src_rowptr, torch.int64, 'src_rowptr')
        d.src_col = capi.dptr(self.src_col, torch.int32, 'src_col')
        d.grid_rowptr = capi.dptr(self.grid_rowptr, torch.int64, 'grid_rowptr')
        d.grid_col = capi.dptr(self.grid_col, torch.int32, 'grid_col')
        d.grid_outdeg = capi.dptr(self.grid_outdeg, torch.int32, 'grid_outdeg')
        d.prod_grid = capi.dptr(self.prod_grid, torch.int32, 'prod_grid') if prod_grid is not None else None
        d.grid_order = capi.dptr(self.grid_order, torch.int32, 'grid_order') if self.grid_order is not None else None
        if self.tiles is not None:
            t = self.tiles
            d.n_sta_tiles, d.n_grid_groups = t['n_tiles'], t['n_groups']
            d.sta_tile_rows = capi.dptr(t['rows'], torch.int32, 'sta_tile_rows')
            d.sta_tile_meta = capi.dptr(t['meta'], torch.int32, 'sta_tile_meta')
            d.sta_tile_nbr = capi.dptr(t['nbr'], torch.int16, 'sta_tile_nbr')
            d.sta_tile_invdeg = capi.dptr(t['invdeg'], torch.float32, 'sta_tile_invdeg')
            d.grid_grp_ptr = capi.dptr(t['grp_ptr'], torch.int32, 'grid_grp_ptr')
            d.grid_grp_nodes = capi.dptr(t['grp_nodes'], torch.int32, 'grid_grp_nodes')
        d.n_grid_owned = self.n_grid_owned
        self._desc = d
        lib = capi.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            capi.check(lib.genie_plan_create(ctypes.byref(d), ctypes.byref(handle)))
        self.handle = handle
        self.workspace_bytes = int(lib.genie_plan_workspace_bytes(handle))
        self._workspace = None
        self.storage = 'fp32'

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None and h.value:
            try:
                capi.load().genie_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_storage(self, storage):
        """genie_plan_set_storage: 'fp32' (default; within 1e-4 of the reference) or 'bf16' (the gathered intermediate rows
        kept as bf16: the fast inference mode of BASELINE.json configs[1], ~1e-3 relative).  Re-sizes the workspace."""
        code = {'fp32': capi.STORAGE_FP32, 'bf16': capi.STORAGE_BF16}[storage]
        capi.check(capi.load().genie_plan_set_storage(self.handle, code))
        if storage != self.storage:
            self.storage = storage
            self.workspace_bytes = int(capi.load().genie_plan_workspace_bytes(self.handle))
            self._workspace = None

    def set_halo_export(self, exp_ptr, exp_peer, exp_row, peer_base, halo_ptr):
        """genie_plan_set_halo_export (grid sharding): int32 device tensors of the export CSR, an int64 device tensor of the
        peers' landing-buffer addresses and the address of this rank's landing buffer; all None switches the export off."""
        if exp_ptr is None:
            capi.check(capi.load().genie_plan_set_halo_export(self.handle, None, None, None, None, None))
            self._halo_export = None
            return
        if exp_ptr.numel() != self.n_grid_owned + 1 or exp_peer.numel() != exp_row.numel() or exp_peer.numel() == 0:
            raise capi.GenieError('halo export tables do not match the plan')
        capi.check(capi.load().genie_plan_set_halo_export(
            self.handle, capi.dptr(exp_ptr, torch.int32, 'exp_ptr'), capi.dptr(exp_peer, torch.int32, 'exp_peer'),
            capi.dptr(exp_row, torch.int32, 'exp_row'), capi.dptr(peer_base, torch.int64, 'peer_base'),
            ctypes.c_void_p(int(halo_ptr))))
        self._halo_export = (exp_ptr, exp_peer, exp_row, peer_base)      # keep the tensors alive

    def set_edge_terms(self, t_sta, t_src):
        """genie_plan_set_edge_terms: per-node additive terms of the edge-feature model ([n, 48] fp32 each), or None, None."""
        if t_sta is None:
            capi.check(capi.load().genie_plan_set_edge_terms(self.handle, None, None))
        else:
            rows = (self.n_sta, self.n_grid) if self.mode == capi.GRAPH_CARTESIAN else (self.n_prod, self.n_prod)
            for t, n in ((t_sta, rows[0]), (t_src, rows[1])):
                if tuple(t.shape) != (n, capi.EDGE_TERM_LD):
                    raise capi.GenieError('edge-term table must be [%d, %d]' % (n, capi.EDGE_TERM_LD))
            capi.check(capi.load().genie_plan_set_edge_terms(self.handle, capi.dptr(t_sta, torch.float32, 'edge_sta'),
                                                             capi.dptr(t_src, torch.float32, 'edge_src')))
        self._edge_terms = (t_sta, t_src)              # keep the tensors alive

    def set_init_terms(self, t_sta, t_src):
        """genie_plan_set_init_terms (use_absolute_pos): [S,32] + [G,32] (CARTESIAN) or [P,32] + None (EXPLICIT); None, None = off."""
        if t_sta is not None:
            rows = (self.n_sta, self.n_grid) if self.mode == capi.GRAPH_CARTESIAN else (self.n_prod, None)
            if tuple(t_sta.shape) != (rows[0], 32) or (rows[1] is None) != (t_src is None) or \
                    (t_src is not None and tuple(t_src.shape) != (rows[1], 32)):
                raise capi.GenieError('init-term tables must be [%s, 32] and [%s, 32]' % rows)
        capi.check(capi.load().genie_plan_set_init_terms(self.handle, capi.dptr(t_sta, torch.float32, 'init_sta'),
                                                         capi.dptr(t_src, torch.float32, 'init_src')))
        self._init_terms = (t_sta, t_src)              # keep the tensors alive

    def set_assoc_terms(self, init_sta, init_src, edge_sta, edge_src):
        """genie_assoc_set_terms: the init / edge-term tables of the association branch's model variants (None = off)."""
        capi.check(capi.load().genie_assoc_set_terms(
            self.handle, capi.dptr(init_sta, torch.float32, 'assoc init_sta'), capi.dptr(init_src, torch.float32, 'assoc init_src'),
            capi.dptr(edge_sta, torch.float32, 'assoc edge_sta'), capi.dptr(edge_src, torch.float32, 'assoc edge_src')))
        self._assoc_terms = (init_sta, init_src, edge_sta, edge_src)       # keep the tensors alive

    def node_grid_index(self):
        """int64 [P]: grid node of every product node (CARTESIAN: i // n_sta; EXPLICIT: the read-in target list)."""
        if self.prod_grid is not None:
            return self.prod_grid.long()
        return torch.arange(self.n_prod, device=self.device) // self.n_sta

    def workspace(self):
        if self._workspace is None:
            self._workspace = torch.empty(max(self.workspace_bytes, 256), dtype=torch.uint8, device=self.device)
        return self._workspace

    # ---- constructors ------------------------------------------------------------------------------------------------
    @staticmethod
    def _grid_parts(A_src, n_grid):
        grid = csr_by_destination(A_src, n_grid)
        outdeg = torch.bincount(A_src[0].long(), minlength=n_grid).to(torch.int32).contiguous()
        return grid, outdeg

    @classmethod
    def cartesian(cls, A_sta_sta, A_src_src, n_sta, n_grid, A_src=None, device=None, grid_order=None, tiling=True,
                  group_size=GROUP_SIZE, n_grid_owned=0, grid_groups=None):
        """Dense mode from the two small kNN graphs (process_utils.py:718-719); product edges stay implicit."""
        device = torch.device(device if device is not None else A_sta_sta.device)
        A_sta_sta, A_src_src = A_sta_sta.to(device), A_src_src.to(device)
        sta = csr_by_destination(A_sta_sta, n_sta)
        src = csr_by_destination(A_src_src, n_grid)
        if A_src is None or A_src is A_src_src:
            grid = src
            outdeg = torch.bincount(A_src_src[0].long(), minlength=n_grid).to(torch.int32).contiguous()
        else:
            grid, outdeg = cls._grid_parts(A_src.to(device), n_grid)
        return cls(capi.GRAPH_CARTESIAN, n_sta, n_grid, n_sta * n_grid, sta, src, grid, outdeg, None, device,
                   grid_order=grid_order, tiling=tiling, group_size=group_size, n_grid_owned=n_grid_owned,
                   grid_groups=grid_groups)

    @classmethod
    def grid_only(cls, A_src, n_grid, device):
        """A plan with no product nodes: only the grid graph of SpatialAggregation (module.py:243)."""
        device = torch.device(device)
        empty = torch.zeros((2, 0), dtype=torch.long, device=device)
        grid, outdeg = cls._grid_parts(A_src.to(device), n_grid)
        return cls(capi.GRAPH_EXPLICIT, 0, n_grid, 0, csr_by_destination(empty, 0), csr_by_destination(empty, 0), grid,
                   outdeg, torch.zeros(0, dtype=torch.int32, device=device), device)

    @classmethod
    def explicit(cls, A_in_sta, A_in_src, prod_target, A_src, n_prod, n_grid, device=None):
        """Arbitrary product graph (sub-graph mode): CSR over the product nodes themselves."""
        device = torch.device(device if device is not None else A_in_sta.device)
        sta = csr_by_destination(A_in_sta.to(device), n_prod)
        src = csr_by_destination(A_in_src.to(device), n_prod)
        grid, outdeg = cls._grid_parts(A_src.to(device), n_grid)
        prod_grid = prod_target.to(device).to(torch.int32).contiguous()
        return cls(capi.GRAPH_EXPLICIT, 0, n_grid, n_prod, sta, src, grid, outdeg, prod_grid, device)

    @classmethod
    def from_edge_lists(cls, A_in_sta, A_in_src, read_in_index, A_src, n_sta, n_grid, device=None):
        """What set_adjacencies receives (module.py:941): picks CARTESIAN when the lists follow :720-722."""
        device = torch.device(device if device is not None else A_in_sta.device)
        A_in_sta, A_in_src, A_src = A_in_sta.to(device), A_in_src.to(device), A_src.to(device)
        n_prod = int(read_in_index.shape[1])
        src_nodes = read_in_index[0].to(device)
        if not torch.equal(src_nodes.long(), torch.arange(n_prod, device=device)):
            raise ValueError('A_src_in_edges.edge_index[0] must be arange(P) (module.py:229, process_utils.py:722)')
        prod_target = read_in_index[1].to(device)
        if n_sta * n_grid == n_prod:
            small = _is_cartesian(A_in_sta, A_in_src, prod_target, n_sta, n_grid)
            if small is not None:
                return cls.cartesian(small[0], small[1], n_sta, n_grid, A_src=A_src, device=device)
        return cls.explicit(A_in_sta, A_in_src, prod_target, A_src, n_prod, n_grid, device=device)
