"""GraphDD's location network on the libgenie_b200 kernels — the second consumer of the DataAggregation kernel family
(SURVEY.md §8f rank 4): a drop-in for the model classes of `Relocation/train_double_difference_model.py:333-536`
(`DataAggregation`, `BipartiteGraphOperator`, `BipartiteGraphOperatorSta`, `GNN_Location`; same constructors, `forward`,
`forward_fixed`, `set_adjacencies`, same state_dict keys).

GraphDD trains this network (the file is a training script), so every operator here is differentiable.  What differs from
the detection model: every message passes through a per-EDGE layer before the mean, `merge_edges([x_j | (pos_j - pos_i) /
scale_rel])` (:386-388), and the two read-outs apply a per-edge MLP as well (:412-414) — PyG materialises `[E, C]` tensors with
index_select and reduces them with scatter.  Here an edge list becomes an `EdgeGraph` once (edges sorted by target), and
  * the gather x[source(e)]            is `genie_kron_spmm_fwd` with one entry per row      (backward: the transposed CSR, a
                                                                                          gather again — no atomics),
  * the per-edge layers                are `genie_node_mlp_fwd / _bwd` over the E edge rows (no concatenated tensor),
  * the mean over a node's in-edges    is `genie_kron_spmm_fwd` over the sorted edge rows   (backward: a one-entry gather).
Small per-source / per-station tensors (the projection heads, `embed_inpt`) stay torch modules.  CUDA tensors only.
"""
import numpy as np
import torch
from torch import nn

from . import capi, ops
from .training import mlp, note_zero_slopes


class _Csr(torch.autograd.Function):
    """y = A x for an explicit CSR pair (A by rows, A^T by rows): forward and backward are both gathers."""

    @staticmethod
    def forward(ctx, x, fwd, rev, n_out):
        ctx.rev, ctx.n_in = rev, x.shape[0]
        return ops.csr_apply(fwd, x, n_out)

    @staticmethod
    def backward(ctx, gy):
        return ops.csr_apply(ctx.rev, gy, ctx.n_in), None, None, None


def _csr(rows, cols, vals, n_rows, device):
    """COO (row, col, val) -> CSR by row, the order inside a row kept."""
    order = torch.sort(rows, stable=True)[1]
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n_rows), 0)
    return rowptr.contiguous(), cols[order].to(torch.int32).contiguous(), vals[order].float().contiguous()


class EdgeGraph(object):
    """One edge list [2,E] (row 0 = message source j, row 1 = target i; PyG flow source_to_target) over `n_src` source rows and
    `n_tgt` target rows, with its edges sorted by target.  gather(x): x[source(e)] per sorted edge; mean(m): mean of the edge
    rows m over every target's in-edges (0 for a target without edges: PyG's mean of nothing)."""

    def __init__(self, edge_index, n_src, n_tgt):
        dev = edge_index.device
        src, tgt = edge_index[0].long(), edge_index[1].long()
        order = torch.sort(tgt, stable=True)[1]
        self.src, self.tgt = src[order].contiguous(), tgt[order].contiguous()
        self.order = order
        E = int(src.numel())
        self.E, self.n_src, self.n_tgt = E, int(n_src), int(n_tgt)
        ar = torch.arange(E, device=dev)
        ones = torch.ones(E, device=dev)
        deg = torch.bincount(self.tgt, minlength=n_tgt).clamp(min=1).float()
        w = 1.0 / deg[self.tgt]
        # gather: E x n_src with one entry per row; its transpose n_src x E
        self.g_fwd = _csr(ar, self.src, ones, E, dev)
        self.g_rev = _csr(self.src, ar, ones, n_src, dev)
        # mean: n_tgt x E; its transpose E x n_tgt with one entry per row
        self.m_fwd = _csr(self.tgt, ar, w, n_tgt, dev)
        self.m_rev = _csr(ar, self.tgt, w, E, dev)

    def gather(self, x):
        return _Csr.apply(x, self.g_fwd, self.g_rev, self.E)

    def mean(self, m):
        return _Csr.apply(m, self.m_fwd, self.m_rev, self.n_tgt)


_graph_cache = {}


def edge_graph(edge_index, n_src, n_tgt):
    """EdgeGraph of an edge tensor, cached on the tensor itself (kept referenced, so its address cannot be recycled)."""
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(n_src), int(n_tgt), str(edge_index.device))
    ent = _graph_cache.get(key)
    if ent is None or ent[1] is not edge_index:
        if len(_graph_cache) > 64:
            _graph_cache.clear()
        ent = (EdgeGraph(edge_index, n_src, n_tgt), edge_index)
        _graph_cache[key] = ent
    return ent[0]


class DataAggregation(nn.Module):
    """train_double_difference_model.py:333-388 (parameters :339-364 incl. the unused l1_t1_1 / l1_t2_1, kept for the state_dict)."""

    def __init__(self, in_channels, out_channels, n_hidden=30, scale_rel=30.0, n_dim=3, n_dim_mask=2, ndim_proj=3):
        super().__init__()
        self.in_channels, self.out_channels, self.n_hidden = in_channels, out_channels, n_hidden
        self.activate = nn.PReLU()
        self.init_trns = nn.Linear(in_channels + n_dim_mask, n_hidden)
        self.l1_t1_1 = nn.Linear(n_hidden, n_hidden)
        self.l1_t1_2 = nn.Linear(2 * n_hidden + n_dim_mask, n_hidden)
        self.l1_t2_1 = nn.Linear(in_channels, n_hidden)
        self.l1_t2_2 = nn.Linear(2 * n_hidden + n_dim_mask, n_hidden)
        self.activate11, self.activate12, self.activate1 = nn.PReLU(), nn.PReLU(), nn.PReLU()
        self.l2_t1_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t1_2 = nn.Linear(3 * n_hidden + n_dim_mask, out_channels)
        self.l2_t2_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t2_2 = nn.Linear(3 * n_hidden + n_dim_mask, out_channels)
        self.activate21, self.activate22, self.activate2 = nn.PReLU(), nn.PReLU(), nn.PReLU()
        self.scale_rel = scale_rel
        self.merge_edges = nn.Sequential(nn.Linear(n_hidden + ndim_proj, n_hidden), nn.PReLU())

    def _agg(self, eg, x, edge_attr):
        """propagate(A, x=x, edge_attr=...) with message = merge_edges([x_j | edge_attr]) and mean aggregation (:386-388)."""
        return eg.mean(mlp(self.merge_edges[0], self.merge_edges[1], eg.gather(x), edge_attr))

    def forward(self, tr, mask, A_in_sta, A_in_src, A_src_in_sta, pos_loc, pos_src):
        if not tr.is_cuda:
            raise capi.GenieError('genie_b200 has no CPU path: inputs must be CUDA tensors')
        n = tr.shape[0]
        eg_sta, eg_src = edge_graph(A_in_sta, n, n), edge_graph(A_in_src, n, n)
        tr = mlp(self.init_trns, self.activate, tr, mask)                                                      # :359-360
        sta_of, src_of = A_src_in_sta[0], A_src_in_sta[1]
        # edge features in the sorted edge order of the EdgeGraphs (:364-365)
        pl, ps = pos_loc[sta_of] / 1000.0, pos_src[src_of] / 1000.0
        rel_sta = ((pl[eg_sta.src] - pl[eg_sta.tgt]) / self.scale_rel).contiguous()
        rel_src = ((ps[eg_src.src] - ps[eg_src.tgt]) / self.scale_rel).contiguous()
        tr1 = mlp(self.l1_t1_2, self.activate1, tr, self._agg(eg_sta, self.activate11(tr), rel_sta), mask)     # :368
        tr2 = mlp(self.l1_t2_2, self.activate1, tr, self._agg(eg_src, self.activate12(tr), rel_src), mask)     # :369
        tr = torch.cat((tr1, tr2), dim=1)                       # one slope: activate1 of the concatenation = of the halves (:370)
        tr1 = mlp(self.l2_t1_2, self.activate2, tr, self._agg(eg_sta, mlp(self.l2_t1_1, self.activate21, tr), rel_sta), mask)
        tr2 = mlp(self.l2_t2_2, self.activate2, tr, self._agg(eg_src, mlp(self.l2_t2_1, self.activate22, tr), rel_src), mask)
        return torch.cat((tr1, tr2), dim=1)                                                                    # :372-376


class BipartiteGraphOperator(nn.Module):
    """train_double_difference_model.py:390-412: product nodes -> sources."""
    to_stations = False

    def __init__(self, ndim_in, ndim_out, ndim_mask=11, ndim_edges=3, scale_rel=30e3):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(ndim_in + ndim_edges + ndim_mask, ndim_in), nn.PReLU(), nn.Linear(ndim_in, ndim_in))
        self.fc2 = nn.Linear(ndim_in, ndim_out)
        self.activate1, self.activate2 = nn.PReLU(), nn.PReLU()
        self.scale_rel = scale_rel

    def forward(self, x, mask, A_src_in_edges, A_src_in_sta, locs_cart, src_cart):
        N = x.shape[0]
        if self.to_stations:                                     # :428-430  pos = (src_cart[A_src_in_sta[1]], locs_cart)
            pos_j, pos_i = src_cart[A_src_in_sta[1]], locs_cart
        else:                                                    # :405-407  pos = (locs_cart[A_src_in_sta[0]], src_cart)
            pos_j, pos_i = locs_cart[A_src_in_sta[0]], src_cart
        eg = edge_graph(A_src_in_edges, N, pos_i.shape[0])
        rel = ((pos_i[eg.tgt] - pos_j[eg.src]) / self.scale_rel).contiguous()                                   # :411, :434
        h = mlp(self.fc1[0], self.fc1[1], eg.gather(x), eg.gather(mask), rel)
        h = mlp(self.fc1[2], self.activate1, h)
        return self.activate2(self.fc2(eg.mean(h)))


class BipartiteGraphOperatorSta(BipartiteGraphOperator):
    """train_double_difference_model.py:414-436: product nodes -> stations."""
    to_stations = True


class GNN_Location(nn.Module):
    """train_double_difference_model.py:438-536."""

    def __init__(self, ftrns1, ftrns2, inpt_sources=True, use_sta_corr=True, use_memory=False, use_mask=False,
                 use_aggregation=True, use_attention=False, n_inpt=15, n_mask=15, n_hidden=20, n_embed=10, scale_fixed=5000.0,
                 device='cuda'):
        super().__init__()
        if inpt_sources:
            n_inpt, n_mask = n_inpt + 3, n_mask + 3
        if use_memory:
            n_read_out = 30
            self.proj_memory = nn.Sequential(nn.Linear(4, 30), nn.PReLU(), nn.Linear(30, 15))
            self.merge_data = nn.Sequential(nn.Linear(30, 30), nn.PReLU(), nn.Linear(30, n_read_out))
            n_inpt, n_mask = n_inpt + 4, n_mask + 4
        else:
            n_read_out = 15
        self.DataAggregation1 = DataAggregation(n_inpt, 15, n_dim_mask=n_embed).to(device)
        self.DataAggregation2 = DataAggregation(30, 15, n_dim_mask=n_embed).to(device)
        self.DataAggregation3 = DataAggregation(30, 15, n_dim_mask=n_embed).to(device)
        self.DataAggregation4 = DataAggregation(30, 15, n_dim_mask=n_embed).to(device)
        self.DataAggregation5 = DataAggregation(30, 15, n_dim_mask=n_embed).to(device)
        self.BipartiteReadOut1 = BipartiteGraphOperator(30, 15, ndim_mask=n_embed)
        self.BipartiteReadOut2 = BipartiteGraphOperatorSta(30, 15, ndim_mask=n_embed)
        self.embed_inpt = nn.Sequential(nn.Linear(n_mask, n_hidden), nn.PReLU(), nn.Linear(n_hidden, n_embed))
        self.proj = nn.Sequential(nn.Linear(n_read_out, 30), nn.PReLU(), nn.Linear(30, 3))
        self.proj_t = nn.Sequential(nn.Linear(n_read_out, 15), nn.PReLU(), nn.Linear(15, 1))
        self.proj_c = nn.Sequential(nn.Linear(15, 15), nn.PReLU(), nn.Linear(15, 2))
        if use_mask:
            self.proj_mask = nn.Sequential(nn.Linear(30, 15), nn.PReLU(), nn.Linear(15, 2))
        self.use_memory, self.use_sta_corr = use_memory, use_sta_corr
        self.register_buffer('scale', torch.Tensor([scale_fixed]), persistent=False)     # a plain attribute in the reference
        self.device = device
        self.ftrns1, self.ftrns2 = ftrns1, ftrns2
        self.to(device)

    def forward(self, x, mask, A_in_pick, A_in_src, A_src_in_product, A_sta_in_product, A_src_in_sta, locs_cart, srcs_cart,
                memory=False):
        note_zero_slopes(self)           # a PReLU slope of exactly 0 sends its layer through the torch ops (training.mlp)
        if self.use_memory:                                                                                     # :486-488
            mask = self.embed_inpt(torch.cat((mask, memory[A_src_in_sta[1]]), dim=1))
            x = torch.cat((x, memory[A_src_in_sta[1]]), dim=1)
        else:
            mask = self.embed_inpt(mask)
        for da in (self.DataAggregation1, self.DataAggregation2, self.DataAggregation3, self.DataAggregation4,
                   self.DataAggregation5):
            x = da(x, mask, A_in_pick, A_in_src, A_src_in_sta, locs_cart, srcs_cart)
        x1 = self.BipartiteReadOut1(x, mask, A_src_in_product, A_src_in_sta, locs_cart, srcs_cart)
        x2 = self.BipartiteReadOut2(x, mask, A_sta_in_product, A_src_in_sta, locs_cart, srcs_cart)
        if self.use_memory:
            x1 = self.merge_data(torch.cat((x1, self.proj_memory(memory)), dim=1))
        return self.scale * self.proj(x1), self.proj_t(x1), self.proj_c(x2), x

    def set_adjacencies(self, A_in_pick, A_in_src, A_src_in_product, A_sta_in_product, A_src_in_sta, locs_cart, srcs_cart):
        self.A_in_pick, self.A_in_src = A_in_pick, A_in_src
        self.A_src_in_product, self.A_sta_in_product, self.A_src_in_sta = A_src_in_product, A_sta_in_product, A_src_in_sta
        self.locs_cart, self.srcs_cart = locs_cart, srcs_cart

    def forward_fixed(self, x, mask, memory=False):
        return self.forward(x, mask, self.A_in_pick, self.A_in_src, self.A_src_in_product, self.A_sta_in_product,
                            self.A_src_in_sta, self.locs_cart, self.srcs_cart, memory=memory)
