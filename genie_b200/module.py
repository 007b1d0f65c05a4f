"""Drop-in replacement for the hot-path classes of the reference's `Code/module.py`.

`from genie_b200.module import *` gives `GCN_Detection_Network_extended` with the reference's constructor, method
signatures and state_dict key names (module.py:882-1020), so `train_GENIE_model.py:1382, 1786` / `process_continuous_days.py:
335-337, 634, 797, 1062` work against it unchanged.  Everything that scales with the product graph runs in libgenie_b200.so
(the sub-modules of the reference's names only hold the parameters):
  forward_fixed_source  front end (DataAggregation -> Bipartite_ReadIn -> SpatialAggregation1..3, >= 97 % of the reference's
                        forward time) + read-out heads (SpatialDirect, TemporalAttention, SpatialAttention)
  forward_fixed         the same + the association branch (BipartiteGraphReadOutOperator, DataAggregationAssociationPhase,
                        LocalSliceLgCollapse{P,S}); only the pick-sized source-arrival attention (Arrivals) is plain torch
  forward               under torch.no_grad() = forward_fixed with the graphs of the call; with gradients the differentiable
                        path of genie_b200/training.py (message passing and its gradient on genie_kron_spmm_fwd)
The torch `forward` methods of the head modules below are kept for head shapes the kernels are not built for and for training.

Like the reference (module.py:27-46) the module reads `config.yaml` / `train_config.yaml` from the working directory at
import time when they exist; otherwise the reference's shipped defaults are used.
"""
import math
import os

import numpy as np
import torch
from torch import nn

from . import capi, ops
from .plan import GraphPlan

# ---- import-time configuration (module.py:27-46) ----------------------------------------------------------------------
_cfg, _tcfg = {}, {}
try:
    import yaml
    if os.path.exists('config.yaml'):
        with open('config.yaml', 'r') as _f:
            _cfg = yaml.safe_load(_f) or {}
    if os.path.exists('train_config.yaml'):
        with open('train_config.yaml', 'r') as _f:
            _tcfg = yaml.safe_load(_f) or {}
except ImportError:          # pragma: no cover
    pass

use_updated_model_definition = bool(_cfg.get('use_updated_model_definition', False))
scale_rel = float(_cfg.get('scale_rel', 30000.0))
k_sta_edges = int(_cfg.get('k_sta_edges', 8))
kernel_sig_t = float(_tcfg.get('kernel_sig_t', 3.0))
scale_t = kernel_sig_t * 3.0
eps = kernel_sig_t * 5.0
use_phase_types = bool(_cfg.get('use_phase_types', True))
use_absolute_pos = bool(_cfg.get('use_absolute_pos', False))

device = torch.device('cuda')

# `from genie_b200.module import *` placed after the reference's `from module import *` overrides exactly these names
__all__ = ['DataAggregation', 'DataAggregationEdges', 'BipartiteGraphOperator', 'SpatialAggregation', 'SpatialDirect',
           'SpatialAttention', 'TemporalAttention', 'BipartiteGraphReadOutOperator', 'DataAggregationAssociationPhase',
           'DataAggregationAssociationPhaseEdges', 'LocalSliceLgCollapse', 'StationSourceAttentionMergedPhases',
           'GCN_Detection_Network_extended', 'knn_query_edges', 'use_updated_model_definition', 'scale_rel', 'k_sta_edges',
           'scale_t', 'eps', 'use_phase_types', 'use_absolute_pos', 'device']


# ---- parameter holders of the CUDA front end -------------------------------------------------------------------------

class DataAggregation(nn.Module):
    """Parameters of module.py:52-83 (the unused l1_t1_1 / l1_t2_1 stay in the state_dict).  Forward: libgenie_b200."""

    n_edge = 0      # edge-feature channels of every message (DataAggregationEdges: 4)

    def __init__(self, in_channels, out_channels, n_hidden=30, n_dim_mask=4, use_absolute_pos=use_absolute_pos):
        super().__init__()
        if (in_channels, out_channels, n_hidden, n_dim_mask) != (4, 15, 30, 4):
            raise NotImplementedError('genie_b200: DataAggregation is built for (4, 15, n_hidden=30, n_dim_mask=4)')
        self.use_absolute_pos = bool(use_absolute_pos)
        if use_absolute_pos:
            in_channels = in_channels + 3 * 2         # module.py:56-57: station and source positions join Slice
        self.in_channels, self.out_channels, self.n_hidden = in_channels, out_channels, n_hidden
        ne = self.n_edge
        self.activate = nn.PReLU()
        self.init_trns = nn.Linear(in_channels + n_dim_mask, n_hidden)
        self.l1_t1_1 = nn.Linear(n_hidden, n_hidden)
        self.l1_t1_2 = nn.Linear(2 * n_hidden + n_dim_mask + ne, n_hidden)
        self.l1_t2_1 = nn.Linear(in_channels, n_hidden)
        self.l1_t2_2 = nn.Linear(2 * n_hidden + n_dim_mask + ne, n_hidden)
        self.activate11 = nn.PReLU()
        self.activate12 = nn.PReLU()
        self.activate1 = nn.PReLU()
        self.l2_t1_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t1_2 = nn.Linear(3 * n_hidden + n_dim_mask + ne, out_channels)
        self.l2_t2_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t2_2 = nn.Linear(3 * n_hidden + n_dim_mask + ne, out_channels)
        self.activate21 = nn.PReLU()
        self.activate22 = nn.PReLU()
        self.activate2 = nn.PReLU()


class DataAggregationEdges(DataAggregation):
    """Parameters of module.py:102-140 (`use_updated_model_definition: True`): every message carries four edge-feature
    channels, so l1_t*_2 / l2_t*_2 take 68 / 98 inputs ordered [tr | mean x_j | mean pos_rel | mask]."""
    n_edge = 4


class BipartiteGraphOperator(nn.Module):
    """Parameters of module.py:214-222."""

    def __init__(self, ndim_in, ndim_out, ndim_edges=3):
        super().__init__()
        if (ndim_in, ndim_out, ndim_edges) != (30, 15, 3):
            raise NotImplementedError('genie_b200: BipartiteGraphOperator is built for (30, 15, ndim_edges=3)')
        self.fc1 = nn.Linear(ndim_in + ndim_edges, ndim_in)
        self.fc2 = nn.Linear(ndim_in, ndim_out)
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()


class SpatialAggregation(nn.Module):
    """Parameters of module.py:231-241."""

    def __init__(self, in_channels, out_channels, scale_rel=scale_rel, n_dim=3, n_global=5, n_hidden=30):
        super().__init__()
        if in_channels not in (15, 30) or (out_channels, n_dim, n_global, n_hidden) != (30, 3, 5, 30):
            raise NotImplementedError('genie_b200: SpatialAggregation is built for (15|30 -> 30)')
        self.fc1 = nn.Linear(in_channels + n_dim + n_global, n_hidden)
        self.fc2 = nn.Linear(n_hidden + in_channels, out_channels)
        self.fglobal = nn.Linear(in_channels, n_global)
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()
        self.activate3 = nn.PReLU()
        self.scale_rel = scale_rel


# ---- read-out heads (plain torch) -------------------------------------------------------------------------------------

class SpatialDirect(nn.Module):
    """module.py:251-260."""

    def __init__(self, inpt_dim, out_channels):
        super().__init__()
        self.f_direct = nn.Linear(inpt_dim, out_channels)
        self.activate = nn.PReLU()

    def forward(self, inpts):
        return self.activate(self.f_direct(inpts))


def knn_query_edges(x_context, x_query, k):
    """`knn(x_context/1000, x_query/1000, k).flip(0)` (module.py:282) without torch_cluster: libgenie_b200's brute-force
    kNN kernel (genie_knn_fwd).

    Returns int64 [2, Q*k]: row 0 = context (source) index, nearest first, row 1 = query (target) index."""
    idx = ops.knn(x_context / 1000.0, x_query / 1000.0, k)
    tgt = torch.arange(x_query.shape[0], device=x_context.device).repeat_interleave(idx.shape[1])
    return torch.stack((idx.reshape(-1), tgt), dim=0)


class SpatialAttention(nn.Module):
    """module.py:262-297 (k-nearest-context attention read-out onto arbitrary query points)."""

    def __init__(self, inpt_dim, out_channels, n_dim, n_latent, n_hidden=30, n_heads=5, scale_rel=scale_rel):
        super().__init__()
        self.param_vector = nn.Parameter(nn.init.xavier_uniform_(torch.Tensor(1, n_heads, n_latent)))
        self.f_queries = nn.Linear(n_dim, n_heads * n_latent)
        self.f_context = nn.Linear(inpt_dim + n_dim, n_heads * n_latent)
        self.f_values = nn.Linear(inpt_dim + n_dim, n_heads * n_latent)
        self.f_direct = nn.Linear(inpt_dim, out_channels)
        self.proj = nn.Linear(n_latent, out_channels)
        self.scale = np.sqrt(n_latent)
        self.n_heads, self.n_latent, self.scale_rel = n_heads, n_latent, scale_rel
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()
        self._edge_cache = None

    def _edges(self, x_query, x_context, k):
        key = (x_query.data_ptr(), x_query._version, tuple(x_query.shape), x_context.data_ptr(), x_context._version,
               tuple(x_context.shape), k)
        if self._edge_cache is None or self._edge_cache[0] != key:
            self._edge_cache = (key, knn_query_edges(x_context, x_query, k), x_query, x_context, None)
        return self._edge_cache[1]

    def _nbr_table(self, edge_index, Q):
        """[Q, k] context node of every query edge (the k consecutive edges of a query), cached with the edges."""
        c = self._edge_cache
        if len(c) < 5 or c[4] is None or c[4].shape[0] != Q:
            self._edge_cache = c[:4] + (edge_index[0].view(Q, -1).contiguous(),)
        return self._edge_cache[4]

    def forward(self, inpts, x_query, x_context, k=10, cache=True):
        """cache=False: one-off query sets (the association sources x_query_src_cart) do not evict the cached kNN edges of
        the fixed query grid."""
        edge_index = self._edges(x_query, x_context, k) if cache else knn_query_edges(x_context, x_query, k)
        Q = x_query.shape[0]
        kk = edge_index.shape[1] // max(Q, 1)
        edge_attr = (x_query[edge_index[1]] - x_context[edge_index[0]]) / self.scale_rel
        x_j = inpts[edge_index[0]]
        cat = torch.cat((x_j, edge_attr), dim=-1)
        q = self.f_queries(edge_attr).view(-1, self.n_heads, self.n_latent)
        c = self.f_context(cat).view(-1, self.n_heads, self.n_latent)
        v = self.f_values(cat).view(-1, self.n_heads, self.n_latent)
        alpha = self.activate1((q * c).sum(-1) / self.scale)
        # every query owns exactly kk consecutive edges -> the segment softmax / sum are dense over a [Q, kk] view
        alpha = torch.softmax(alpha.view(Q, kk, self.n_heads), dim=1)
        out = (alpha.unsqueeze(-1) * v.view(Q, kk, self.n_heads, self.n_latent)).sum(1)
        return self.activate2(self.proj(out.mean(1)))


class TemporalAttention(nn.Module):
    """module.py:299-331."""

    def __init__(self, inpt_dim, out_channels, n_latent, n_hidden=30, n_heads=5, scale_t=scale_t):
        super().__init__()
        self.temporal_query_1 = nn.Linear(1, n_hidden)
        self.temporal_query_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.f_context_1 = nn.Linear(inpt_dim, n_hidden)
        self.f_context_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.f_values_1 = nn.Linear(inpt_dim, n_hidden)
        self.f_values_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.proj_1 = nn.Linear(n_latent, n_hidden)
        self.proj_2 = nn.Linear(n_hidden, out_channels)
        self.scale = np.sqrt(n_latent)
        self.n_heads, self.n_latent, self.scale_t = n_heads, n_latent, scale_t
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()
        self.activate3 = nn.PReLU()
        self.activate4 = nn.PReLU()
        self.activate5 = nn.PReLU()

    def forward(self, inpts, t_query):
        c = self.f_context_2(self.activate1(self.f_context_1(inpts))).view(-1, self.n_heads, self.n_latent)
        v = self.f_values_2(self.activate2(self.f_values_1(inpts))).view(-1, self.n_heads, self.n_latent)
        q = self.temporal_query_2(self.activate3(self.temporal_query_1(t_query / self.scale_t)))
        q = q.view(-1, self.n_heads, self.n_latent)
        s = torch.einsum('nhl,thl->nth', c, q) / self.scale                       # [N, T, heads]
        z = torch.einsum('nth,nhl->ntl', s, v) / self.n_heads                     # mean over heads
        return self.proj_2(self.activate5(self.proj_1(self.activate4(z))))


# ---- association branch: parameters only (not on the forward_fixed_source path) -----------------

class BipartiteGraphReadOutOperator(nn.Module):
    """Parameters of module.py:333-343."""

    def __init__(self, ndim_in, ndim_out, ndim_edges=3):
        super().__init__()
        self.fc1 = nn.Linear(ndim_in + ndim_edges, ndim_in)
        self.fc2 = nn.Linear(ndim_in, ndim_out)
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()


class DataAggregationAssociationPhase(nn.Module):
    """Parameters of module.py:356-387."""

    n_edge = 0

    def __init__(self, in_channels, out_channels, n_hidden=30, n_dim_latent=30, n_dim_mask=5,
                 use_absolute_pos=use_absolute_pos):
        super().__init__()
        if use_absolute_pos:
            in_channels = in_channels + 2 * 3          # module.py:361-362
        ne = self.n_edge
        self.activate = nn.PReLU()
        self.init_trns = nn.Linear(in_channels + n_dim_latent + n_dim_mask, n_hidden)
        self.l1_t1_1 = nn.Linear(n_hidden, n_hidden)
        self.l1_t1_2 = nn.Linear(2 * n_hidden + n_dim_mask + ne, n_hidden)
        self.l1_t2_1 = nn.Linear(n_hidden, n_hidden)
        self.l1_t2_2 = nn.Linear(2 * n_hidden + n_dim_mask + ne, n_hidden)
        self.activate11 = nn.PReLU()
        self.activate12 = nn.PReLU()
        self.activate1 = nn.PReLU()
        self.l2_t1_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t1_2 = nn.Linear(3 * n_hidden + n_dim_mask + ne, out_channels)
        self.l2_t2_1 = nn.Linear(2 * n_hidden, n_hidden)
        self.l2_t2_2 = nn.Linear(3 * n_hidden + n_dim_mask + ne, out_channels)
        self.activate21 = nn.PReLU()
        self.activate22 = nn.PReLU()
        self.activate2 = nn.PReLU()


class DataAggregationAssociationPhaseEdges(DataAggregationAssociationPhase):
    """Parameters of module.py:406-440."""
    n_edge = 4


class LocalSliceLgCollapse(nn.Module):
    """Parameters of module.py:610-622.  Forward: libgenie_b200 (genie_assoc_collapse_fwd, P and S in one launch)."""

    def __init__(self, ndim_in, ndim_out, n_edge=2, n_hidden=30, eps=eps, use_phase_types=use_phase_types,
                 device='cuda'):
        super().__init__()
        self.fc1 = nn.Linear(ndim_in + n_edge, n_hidden)
        self.fc2 = nn.Linear(n_hidden, ndim_out)
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()
        self.eps, self.device, self.use_phase_types = eps, device, use_phase_types


class StationSourceAttentionMergedPhases(nn.Module):
    """Parameters of module.py:662-699."""

    def __init__(self, ndim_src_in, ndim_arv_in, ndim_out, n_latent, ndim_extra=1, n_heads=5, n_hidden=30,
                 scale_rel=scale_rel, k_sta_edges=k_sta_edges, eps=eps, use_neighbor_assoc_edges=False,
                 use_phase_types=use_phase_types, device=device):
        super().__init__()
        if use_neighbor_assoc_edges:
            ndim_extra = ndim_extra + 1 + 3
        self.f_arrival_query_1 = nn.Linear(2 * ndim_arv_in + 6, n_hidden)
        self.f_arrival_query_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.f_src_context_1 = nn.Linear(ndim_src_in + ndim_extra + 2, n_hidden)
        self.f_src_context_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.f_values_1 = nn.Linear(2 * ndim_arv_in + ndim_extra + 7, n_hidden)
        self.f_values_2 = nn.Linear(n_hidden, n_heads * n_latent)
        self.proj_1 = nn.Linear(n_latent, n_hidden)
        self.proj_2 = nn.Linear(n_hidden, ndim_out)
        self.activate1 = nn.PReLU()
        self.activate2 = nn.PReLU()
        self.activate3 = nn.PReLU()
        self.activate4 = nn.PReLU()
        self.n_heads, self.n_latent, self.eps = n_heads, n_latent, eps
        self.scale = np.sqrt(n_latent)
        self.use_phase_types = use_phase_types
        self.use_neighbor_assoc_edges = use_neighbor_assoc_edges
        if use_neighbor_assoc_edges:
            raise NotImplementedError('genie_b200: use_neighbor_assoc_edges=True is not supported')

    def forward(self, src, stime, src_embed, trv_src, locs_cart, arrival_p, arrival_s, tpick, ipick, phase_label):
        """module.py:698-744 with the reference's signature (`src`, `locs_cart` are unused there as well)."""
        z = arrival_p.new_zeros((1, arrival_p.shape[1]))
        arrival = torch.cat((torch.cat((arrival_p, z), dim=0), torch.cat((arrival_s, z), dim=0)), dim=1)     # :709-711
        return self.forward_merged(stime, src_embed, trv_src, arrival, tpick, ipick, phase_label)

    def forward_merged(self, stime, src_embed, trv_src, arrival, tpick, ipick, phase_label):
        """Source-arrival attention (module.py:698-781) on `arrival` [n_arv + 1, 2 * ndim_arv_in] = [P embedding | S embedding]
        with the null arrival as last row (what genie_assoc_collapse_fwd writes).  Pick-sized work (n_src x sum over
        stations of n_s (n_s + 1) edges): plain torch on the tensors' device, no torch_geometric / cKDTree / Python loop —
        the same-station pairs come from one comparison matrix, the segment softmax from scatter-amax / index_add."""
        dev = arrival.device
        n_src, n_sta, n_arv = int(trv_src.shape[0]), int(trv_src.shape[1]), int(tpick.shape[0])
        eps = float(self.eps)
        tpick = tpick.reshape(-1).float()
        ipick = ipick.reshape(-1).long()
        stime = stime.reshape(-1).float()
        ph = phase_label.reshape(-1, 1).float()
        if not self.use_phase_types:
            ph = ph * 0.0
        # edges of one source: every pick a (target) <- every pick b of the same station, and <- the null arrival (:714)
        tgt, srcn = torch.nonzero(ipick.view(-1, 1) == ipick.view(1, -1), as_tuple=True)
        ar = torch.arange(n_arv, device=dev)
        e0 = torch.cat((srcn, torch.full((n_arv,), n_arv, dtype=torch.long, device=dev)))
        e1 = torch.cat((tgt, ar))
        n_edge = e0.numel()
        sindex = torch.arange(n_src, device=dev).repeat_interleave(n_edge)                                   # :719
        e0 = e0.repeat(n_src)
        e1 = e1.repeat(n_src) + sindex * n_arv
        atime = torch.cat((tpick, tpick.new_full((1,), -eps)))
        stindex = torch.cat((ipick, ipick.new_full((1,), n_sta)))
        pad = trv_src.new_full((n_src, 1), -eps)
        tsrc_p = torch.cat((trv_src[:, :, 0], pad), dim=1)
        tsrc_s = torch.cat((trv_src[:, :, 1], pad), dim=1)
        phase = torch.cat((ph, ph.new_full((1, 1), -1.0)), dim=0)
        t_kernel_sq = torch.tensor([eps], dtype=torch.float32, device=dev) ** 2
        rel_p = (atime[e0] - (tsrc_p[sindex, stindex[e0]] + stime[sindex])).reshape(-1, 1)
        rel_s = (atime[e0] - (tsrc_s[sindex, stindex[e0]] + stime[sindex])).reshape(-1, 1)
        thr = 2.0 * torch.sqrt(t_kernel_sq)
        keep = torch.nonzero(((rel_p.abs() < thr) | (rel_s.abs() < thr)).reshape(-1), as_tuple=True)[0]      # :727
        e0, e1, sindex, rel_p, rel_s = e0[keep], e1[keep], sindex[keep], rel_p[keep], rel_s[keep]
        M, H, L = n_arv * n_src, self.n_heads, self.n_latent
        if e0.numel() > 0:
            e0max = int(e0.max())           # taken over the kept edges, as the reference's message() does (:762-763)
            f_p = torch.cat((torch.exp(-0.5 * (rel_p ** 2) / t_kernel_sq), torch.sign(rel_p), phase[e0]), dim=1)
            f_s = torch.cat((torch.exp(-0.5 * (rel_s ** 2) / t_kernel_sq), torch.sign(rel_s), phase[e0]), dim=1)
            self_link = (e0 == torch.remainder(e1, e0max)).reshape(-1, 1).float()
            null_link = (e0 == e0max).reshape(-1, 1).float()
            xj = arrival[e0]
            ctx = self.f_src_context_2(self.activate1(self.f_src_context_1(torch.cat(
                (src_embed[sindex], stime[sindex].reshape(-1, 1), self_link, null_link), dim=1)))).view(-1, H, L)
            qry = self.f_arrival_query_2(self.activate2(self.f_arrival_query_1(torch.cat((xj, f_p, f_s), dim=1)))).view(-1, H, L)
            val = self.f_values_2(self.activate3(self.f_values_1(torch.cat(
                (xj, f_p, f_s, self_link, null_link), dim=1)))).view(-1, H, L)
            scores = (qry * ctx).sum(-1) / self.scale
            idx = e1.view(-1, 1).expand(-1, H)
            mx = scores.new_full((M, H), float('-inf')).scatter_reduce(0, idx, scores, reduce='amax', include_self=True)
            ex = (scores - mx.gather(0, idx)).exp()
            den = scores.new_zeros((M, H)).scatter_add_(0, idx, ex)
            alpha = ex / (den.gather(0, idx) + 1e-16)
            agg = val.new_zeros((M, H, L)).index_add_(0, e1, alpha.unsqueeze(-1) * val)
        else:
            agg = arrival.new_zeros((M, H, L))
        out = self.proj_2(self.activate4(self.proj_1(agg.mean(1))))                                          # :742
        return out.view(n_src, n_arv, out.shape[-1])


# ---- the model ----------------------------------------------------------------------------------------------------------

def edge_feature_means(pos, rowptr, col):
    """Mean over every node's in-edges of the reference's edge embedding (module.py:1102-1111):
    v = [pos_j - pos_i, |pos_j - pos_i|], pos_rel = sign(v) * exp(-0.5 v^2 / scale_rel^2).  `rowptr`, `col`: CSR by target
    node with the SOURCE node of every edge; returns a function of scale_rel -> [n, 4] (0 for nodes without in-edges)."""
    n = rowptr.numel() - 1
    deg = (rowptr[1:] - rowptr[:-1])
    tgt = torch.repeat_interleave(torch.arange(n, device=pos.device), deg)
    d = pos[col.long()] - pos[tgt]
    d = torch.cat((d, torch.norm(d, dim=1, keepdim=True)), dim=1)

    def means(scale):
        e = torch.sign(d) * torch.exp(-0.5 * (d ** 2) / (scale ** 2))
        out = torch.zeros((n, 4), dtype=e.dtype, device=e.device).index_add_(0, tgt, e)
        return out / deg.clamp(min=1).to(e.dtype).unsqueeze(1)
    return means


class GCN_Detection_Network_extended(nn.Module):
    """module.py:882-1020, or — `use_updated_model_definition: True` in config.yaml, or `updated_model=True` — the
    DataAggregationEdges definition of module.py:1024-1186.  In the latter the four edge-feature channels of every message
    reduce, by linearity of the mean and of l*_t*_2, to per-node additive terms that do not depend on the window
    (include/genie_b200.h genie_plan_set_edge_terms); the kernels are the same."""

    def __init__(self, ftrns1, ftrns2, scale_rel=scale_rel, use_absolute_pos=use_absolute_pos, device='cuda',
                 updated_model=None):
        super().__init__()
        self.updated_model = use_updated_model_definition if updated_model is None else bool(updated_model)
        self.DataAggregation = (DataAggregationEdges if self.updated_model else DataAggregation)(
            4, 15, use_absolute_pos=use_absolute_pos).to(device)
        self.Bipartite_ReadIn = BipartiteGraphOperator(30, 15, ndim_edges=3).to(device)
        self.SpatialAggregation1 = SpatialAggregation(15, 30, scale_rel=scale_rel).to(device)
        self.SpatialAggregation2 = SpatialAggregation(30, 30, scale_rel=scale_rel).to(device)
        self.SpatialAggregation3 = SpatialAggregation(30, 30, scale_rel=scale_rel).to(device)
        self.SpatialDirect = SpatialDirect(30, 30).to(device)
        self.SpatialAttention = SpatialAttention(30, 30, 3, 15, scale_rel=scale_rel).to(device)
        self.TemporalAttention = TemporalAttention(30, 1, 15).to(device)
        self.BipartiteGraphReadOutOperator = BipartiteGraphReadOutOperator(30, 15).to(device)
        self.DataAggregationAssociationPhase = (DataAggregationAssociationPhaseEdges if self.updated_model
                                                else DataAggregationAssociationPhase)(
            15, 15, use_absolute_pos=use_absolute_pos).to(device)
        self.LocalSliceLgCollapseP = LocalSliceLgCollapse(30, 15, device=device).to(device)
        self.LocalSliceLgCollapseS = LocalSliceLgCollapse(30, 15, device=device).to(device)
        self.Arrivals = StationSourceAttentionMergedPhases(30, 15, 2, 15, n_heads=3, device=device).to(device)
        self.use_absolute_pos = use_absolute_pos
        self.scale_rel = scale_rel
        self.ftrns1, self.ftrns2 = ftrns1, ftrns2
        self._plan = None
        self._plan_key = None
        self._plan_refs = None
        self._packed = None
        self._read_in_attr = None
        self._heads_w = None
        self._assoc_w = None
        self._read_out_key, self._read_out_attr = None, None
        self._edge_means = None       # updated model: (means_sta(scale), means_src(scale)) of the current plan
        self._edge_terms = None       # (key, t_sta, t_src, re-laid weight tensors)
        self._init_terms = None       # use_absolute_pos: (key, re-laid init_trns weight [30,8])
        self._assoc_terms = None      # association branch of the model variants: (key, re-laid weights)

    # -- graph plans ---------------------------------------------------------------------------------------------------
    def set_adjacencies(self, A_in_sta, A_in_src, A_src_in_edges, A_Lg_in_src, A_src_in_sta, A_src, A_edges_p, A_edges_s,
                        dt_partition, tlatent, pos_loc, pos_src):
        """module.py:941-961: cache the graphs; additionally build the GraphPlan the CUDA kernels use."""
        self.A_in_sta, self.A_in_src, self.A_src_in_edges = A_in_sta, A_in_src, A_src_in_edges
        self.A_Lg_in_src, self.A_src_in_sta, self.A_src = A_Lg_in_src, A_src_in_sta, A_src
        self.A_edges_p, self.A_edges_s, self.dt_partition, self.tlatent = A_edges_p, A_edges_s, dt_partition, tlatent
        n_sta, n_grid = int(pos_loc.shape[0]), int(pos_src.shape[0])
        self._plan = GraphPlan.from_edge_lists(A_in_sta, A_in_src, A_src_in_edges.edge_index, A_src, n_sta, n_grid,
                                               device=pos_src.device)
        self._read_in_attr = A_src_in_edges.x.to(pos_src.device).float().contiguous()
        self._plan_key, self._plan_refs = None, None
        self._init_terms = self._assoc_terms = None          # keyed on id(plan): a recycled id must not match
        self._set_edge_means(pos_loc, pos_src, A_src_in_sta)

    def set_adjacencies_cartesian(self, A_sta_sta, A_src_src, read_in_attr, n_sta, n_grid, device=None, pos_loc=None,
                                  pos_src=None, A_edges_p=None, A_edges_s=None, dt_partition=None, tlatent=None):
        """Dense mode without ever materialising the product edge lists (needed beyond a few 10^7 product nodes):
        the two kNN graphs of process_utils.py:718-719 and the read-in edge features `A_src_in_edges.x` [P,3]
        (+ the Cartesian station / grid positions for the updated model's edge features, and the time-pointer tables of
        the association branch when forward_fixed is going to be called)."""
        device = device if device is not None else read_in_attr.device
        self._plan = GraphPlan.cartesian(A_sta_sta, A_src_src, n_sta, n_grid, device=device)
        self._read_in_attr = read_in_attr.to(device).float().contiguous()
        self.A_src = A_src_src
        self._plan_key, self._plan_refs = None, None
        self._init_terms = self._assoc_terms = None
        # association inputs (forward_fixed): the read-out graph A_Lg_in_src is the implicit [g(i); i] with the read-in features
        self.A_Lg_in_src, self.A_edges_p, self.A_edges_s = None, A_edges_p, A_edges_s
        self.dt_partition, self.tlatent = dt_partition, tlatent
        self._set_edge_means(pos_loc, pos_src, None)

    def _set_edge_means(self, pos_loc, pos_src, A_src_in_sta):
        """Updated model: pos_rel_sta / pos_rel_src of module.py:1102-1111, already averaged over every node's in-edges.
        CARTESIAN plans: one row per station / grid node (the product edge (s',g) -> (s,g) has the offset of s' - s);
        EXPLICIT plans: one row per product node."""
        self._edge_means, self._edge_terms = None, None
        if not self.updated_model:
            return
        if pos_loc is None or pos_src is None:
            raise capi.GenieError('the updated model definition needs the station and grid positions (pos_loc, pos_src)')
        plan = self._plan
        dev = plan.device
        pos_loc, pos_src = pos_loc.to(dev).float(), pos_src.to(dev).float()
        if plan.mode == capi.GRAPH_CARTESIAN:
            self._edge_means = (edge_feature_means(pos_loc, plan.sta_rowptr, plan.sta_col),
                                edge_feature_means(pos_src, plan.src_rowptr, plan.src_col))
        else:
            if A_src_in_sta is None:
                raise capi.GenieError('explicit product graphs need A_src_in_sta for the edge features')
            idx = A_src_in_sta.to(dev).long()
            self._edge_means = (edge_feature_means(pos_loc[idx[0]], plan.sta_rowptr, plan.sta_col),
                                edge_feature_means(pos_src[idx[1]], plan.src_rowptr, plan.src_col))

    def _update_edge_terms(self):
        """Tables of genie_plan_set_edge_terms + the l*_t*_2 weights without their four edge-feature columns."""
        da = self.DataAggregation
        ws = (da.l1_t1_2.weight, da.l1_t2_2.weight, da.l2_t1_2.weight, da.l2_t2_2.weight)
        key = (id(self._edge_means), float(self.scale_rel)) + tuple((w.data_ptr(), w._version) for w in ws)
        if self._edge_terms is not None and self._edge_terms[0] == key:
            return self._edge_terms
        m_sta, m_src = self._edge_means[0](float(self.scale_rel)), self._edge_means[1](float(self.scale_rel))
        w11, w12, w21, w22 = (w.detach() for w in ws)

        def table(m, wa, wb):
            t = torch.zeros((m.shape[0], capi.EDGE_TERM_LD), dtype=torch.float32, device=m.device)
            t[:, 0:30] = m @ wa[:, 60:64].t()
            t[:, 32:47] = m @ wb[:, 90:94].t()
            return t.contiguous()
        t_sta, t_src = table(m_sta, w11, w21), table(m_src, w12, w22)
        relaid = (torch.cat((w11[:, :60], w11[:, 64:]), dim=1).contiguous(), torch.cat((w12[:, :60], w12[:, 64:]), dim=1).contiguous(),
                  torch.cat((w21[:, :90], w21[:, 94:]), dim=1).contiguous(), torch.cat((w22[:, :90], w22[:, 94:]), dim=1).contiguous())
        self._edge_terms = (key, t_sta, t_src, relaid)
        self._plan.set_edge_terms(t_sta, t_src)
        return self._edge_terms

    def _plan_for(self, A_in_sta, A_in_src, A_src_in_edges, A_src, n_sta, n_grid, dev, pos=None, A_src_in_sta=None):
        """`forward` receives the graphs on every call (module.py:908): plans are cached on tensor identity."""
        # The key tensors are kept referenced next to the plan (self._plan_refs): a freed tensor's address can be handed out
        # again by the caching allocator for the next sample's same-sized edge lists, which would match a stale key.
        ei = A_src_in_edges.edge_index
        keyed = (A_in_sta, A_in_src, ei, A_src, A_src_in_edges.x)
        key = tuple((t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for t in keyed) + (n_sta, n_grid, str(dev))
        if self._plan is None or self._plan_key != key:
            self._plan = GraphPlan.from_edge_lists(A_in_sta, A_in_src, ei, A_src, n_sta, n_grid, device=dev)
            self._read_in_attr = A_src_in_edges.x.to(dev).float().contiguous()
            self._plan_key, self._plan_refs = key, keyed
            self._init_terms = self._assoc_terms = None
            if self.updated_model:        # module.py:1059-1072: the edge features come from the positions of the call
                if pos is None:
                    raise RuntimeError('the updated model needs the station / grid positions to build its edge features')
                self._set_edge_means(pos[0], pos[1], A_src_in_sta)
        return self._plan

    def set_storage(self, storage):
        """'fp32' (default) or 'bf16': storage of the gathered intermediate rows of the current plan (GraphPlan.set_storage).
        bf16 is the fast inference mode (BASELINE.json configs[1]: ~1e-3 relative to the fp32 reference, NOT within the 1e-4
        parity bar); it runs only on the tensor-core kernel family, which needs PReLU slopes of activate11 / activate12 in
        (1e-3, 1e3) — checked here."""
        if self._plan is None:
            raise RuntimeError('set_adjacencies must be called before set_storage')
        if storage == 'bf16':
            da = self.DataAggregation
            a11, a12 = float(da.activate11.weight.detach().reshape(-1)[0]), float(da.activate12.weight.detach().reshape(-1)[0])
            if not (1e-3 < a11 < 1e3 and 1e-3 < a12 < 1e3):
                raise capi.GenieError('bf16 storage needs activate11 / activate12 slopes in (1e-3, 1e3)')
        self._plan.set_storage(storage)

    # -- CUDA front end ------------------------------------------------------------------------------------------------
    def _packed_weights(self, dev, init_relaid=None):
        if self._packed is None or self._packed.device != torch.device(dev):
            self._packed = ops.PackedWeights(dev)
        relaid = None
        if self.updated_model:
            if self._edge_means is None:
                raise RuntimeError('set_adjacencies must be called before the weights of the updated model are packed')
            relaid = self._update_edge_terms()[3]
        return self._packed.update(self, relaid, init_relaid)

    def _update_init_terms(self, locs_use_cart, x_temp_cuda_cart):
        """use_absolute_pos (module.py:913-914): tables of genie_plan_set_init_terms + init_trns without its six position
        columns.  Rebuilt when the weight, the positions, scale_rel or the plan change."""
        w = self.DataAggregation.init_trns.weight
        plan = self._plan
        key = (w.data_ptr(), w._version, locs_use_cart.data_ptr(), locs_use_cart._version, x_temp_cuda_cart.data_ptr(),
               x_temp_cuda_cart._version, float(self.scale_rel), id(plan))
        if self._init_terms is not None and self._init_terms[0] == key:
            return self._init_terms[1]
        wd = w.detach()
        sc = 3.0 * float(self.scale_rel)

        def table(pos, cols):
            t = torch.zeros((pos.shape[0], 32), dtype=torch.float32, device=wd.device)
            t[:, :30] = (pos.to(wd.device).float() / sc) @ wd[:, cols].t()
            return t
        t_sta, t_src = table(locs_use_cart, slice(4, 7)), table(x_temp_cuda_cart, slice(7, 10))
        if plan.mode == capi.GRAPH_CARTESIAN:
            if t_sta.shape[0] != plan.n_sta or t_src.shape[0] != plan.n_grid:
                raise capi.GenieError('use_absolute_pos: locs_use_cart / x_temp_cuda_cart do not match the plan')
            plan.set_init_terms(t_sta.contiguous(), t_src.contiguous())
        else:
            idx = getattr(self, 'A_src_in_sta', None)
            if idx is None:
                raise capi.GenieError('use_absolute_pos on an explicit product graph needs A_src_in_sta (set_adjacencies)')
            idx = idx.to(wd.device).long()
            plan.set_init_terms((t_sta[idx[0]] + t_src[idx[1]]).contiguous(), None)
        relaid = torch.cat((wd[:, 0:4], wd[:, 10:14]), dim=1).contiguous()
        self._init_terms = (key, relaid)
        return relaid

    def front_end(self, Slice, Mask, x_temp_cuda_cart, want_latent=False, want_readin=False, locs_use_cart=None):
        """DataAggregation -> Bipartite_ReadIn -> SpatialAggregation1..3 in libgenie_b200 (module.py:1010-1014)."""
        if self._plan is None:
            raise RuntimeError('set_adjacencies must be called before forward_fixed*')
        if not Slice.is_cuda:
            raise capi.GenieError('genie_b200 has no CPU path: inputs must be CUDA tensors')
        if torch.is_grad_enabled() and (Slice.requires_grad or any(p.requires_grad for p in
                                                                   self.DataAggregation.parameters())) \
                and self.training:
            raise NotImplementedError('genie_b200: the fused inference kernels keep no activations; for gradients call '
                                      'forward(...) (the differentiable path), else use torch.no_grad() / model.eval()')
        init_relaid = None
        if self.use_absolute_pos:
            if locs_use_cart is None:
                raise capi.GenieError('use_absolute_pos: the front end needs locs_use_cart')
            init_relaid = self._update_init_terms(locs_use_cart, x_temp_cuda_cart)
        packed = self._packed_weights(Slice.device, init_relaid)
        return ops.frontend_fwd(self._plan, packed, Slice, Mask, self._read_in_attr, x_temp_cuda_cart,
                                float(self.scale_rel), want_latent=want_latent, want_readin=want_readin)

    def _heads(self, x_spatial, x_temp_cuda_cart, x_query_cart, t_query):
        """SpatialDirect -> TemporalAttention and SpatialAttention -> TemporalAttention (module.py:1015-1020)."""
        if ops.HeadsWeights.supported(self) and x_query_cart.shape[0] > 0:
            # read-out heads in libgenie_b200 (two kernels); the torch restatement below is kept for other head shapes
            if self._heads_w is None or self._heads_w.device != x_spatial.device:
                self._heads_w = ops.HeadsWeights(x_spatial.device)
            hp, fold, T = self._heads_w.update(self, t_query)
            edges = self.SpatialAttention._edges(x_query_cart, x_temp_cuda_cart, 10)
            nbr = self.SpatialAttention._nbr_table(edges, x_query_cart.shape[0])
            return ops.heads_fwd(self._heads_w, hp, fold, T, x_spatial, x_temp_cuda_cart, x_query_cart, nbr,
                                 float(self.SpatialAttention.scale_rel))
        y_latent = self.SpatialDirect(x_spatial)
        y = self.TemporalAttention(y_latent, t_query)
        x = self.SpatialAttention(x_spatial, x_query_cart, x_temp_cuda_cart)
        x = self.TemporalAttention(x, t_query)
        return y, x

    def forward_fixed_source(self, Slice, Mask, tpick, ipick, phase_label, locs_use_cart, x_temp_cuda_cart,
                             x_query_cart, t_query):
        """module.py:999-1020 -> (y [G,T,1], x [Q,T,1])."""
        with torch.no_grad():
            x_spatial = self.front_end(Slice, Mask, x_temp_cuda_cart, locs_use_cart=locs_use_cart)[0]
            return self._heads(x_spatial, x_temp_cuda_cart, x_query_cart, t_query)

    # -- association branch -----------------------------------------------------------------------
    def _check_read_out_graph(self, A_Lg_in_src):
        """The read-out kernel assumes edge e of A_Lg_in_src runs from grid node g(e) to product node e — the flipped read-in
        list of process_continuous_days.py:632 / :645.  Checked once per tensor."""
        ei = A_Lg_in_src.edge_index
        key = (ei.data_ptr(), ei._version, tuple(ei.shape))
        if self._read_out_key != key:
            plan = self._plan
            ei = ei.to(plan.device)
            ok = ei.shape[1] == plan.n_prod and bool((ei[1] == torch.arange(plan.n_prod, device=plan.device)).all())
            if ok:
                ok = bool((ei[0] == plan.node_grid_index()).all())
            if not ok:
                raise capi.GenieError('A_Lg_in_src.edge_index must be the flipped A_src_in_edges.edge_index '
                                      '([g(i); i], process_continuous_days.py:632)')
            self._read_out_key = key
            self._read_out_attr = A_Lg_in_src.x.to(plan.device).float().contiguous()
        return self._read_out_attr

    def _update_assoc_terms(self, locs_use_cart, x_temp_cuda_cart):
        """Model variants of the association branch (genie_assoc_set_terms): tables + re-laid weights, rebuilt when a weight,
        the positions, scale_rel or the plan change.  Returns the dict AssocWeights.update takes (empty for the default model)."""
        if not (self.updated_model or self.use_absolute_pos):
            return None
        da, plan = self.DataAggregationAssociationPhase, self._plan
        ws = (da.init_trns.weight, da.l1_t1_2.weight, da.l1_t2_2.weight, da.l2_t1_2.weight, da.l2_t2_2.weight)
        key = tuple((w.data_ptr(), w._version) for w in ws) + (
            locs_use_cart.data_ptr(), locs_use_cart._version, x_temp_cuda_cart.data_ptr(), x_temp_cuda_cart._version,
            float(self.scale_rel), id(plan), id(self._edge_means))
        if self._assoc_terms is not None and self._assoc_terms[0] == key:
            return self._assoc_terms[1]
        w0, w11, w12, w21, w22 = (w.detach() for w in ws)
        dev = w0.device
        relaid, init_sta, init_src, edge_sta, edge_src = {}, None, None, None, None
        if self.use_absolute_pos:                     # init_trns input order: [s0 (15) | locs (3) | x_temp (3) | x_latent | mask5]
            sc = 3.0 * float(self.scale_rel)

            def table(pos, cols):
                t = torch.zeros((pos.shape[0], 32), dtype=torch.float32, device=dev)
                t[:, :30] = (pos.to(dev).float() / sc) @ w0[:, cols].t()
                return t
            t_sta, t_src = table(locs_use_cart, slice(15, 18)), table(x_temp_cuda_cart, slice(18, 21))
            if plan.mode == capi.GRAPH_CARTESIAN:
                init_sta, init_src = t_sta.contiguous(), t_src.contiguous()
            else:
                idx = getattr(self, 'A_src_in_sta', None)
                if idx is None:
                    raise capi.GenieError('use_absolute_pos on an explicit product graph needs A_src_in_sta (set_adjacencies)')
                idx = idx.to(dev).long()
                init_sta = (t_sta[idx[0]] + t_src[idx[1]]).contiguous()
            relaid['init_trns'] = torch.cat((w0[:, 0:15], w0[:, 21:56]), dim=1).contiguous()
        if self.updated_model:                        # l1_t*_2: [tr | mean x_j | mean pos_rel (4) | mask5]; l2_t*_2 likewise
            if self._edge_means is None:
                raise RuntimeError('set_adjacencies must be called before forward_fixed of the updated model')
            m_sta, m_src = self._edge_means[0](float(self.scale_rel)), self._edge_means[1](float(self.scale_rel))

            def etable(m, wa, wb):
                t = torch.zeros((m.shape[0], capi.EDGE_TERM_LD), dtype=torch.float32, device=dev)
                t[:, 0:30] = m @ wa[:, 60:64].t()
                t[:, 32:47] = m @ wb[:, 90:94].t()
                return t.contiguous()
            edge_sta, edge_src = etable(m_sta, w11, w21), etable(m_src, w12, w22)
            cut = lambda w, a: torch.cat((w[:, :a], w[:, a + 4:]), dim=1).contiguous()
            relaid.update(l1_t1_2=cut(w11, 60), l1_t2_2=cut(w12, 60), l2_t1_2=cut(w21, 90), l2_t2_2=cut(w22, 90))
        plan.set_assoc_terms(init_sta, init_src, edge_sta, edge_src)
        self._assoc_terms = (key, relaid)
        return relaid

    def _association(self, Slice, Mask, A_Lg_in_src, A_edges_p, A_edges_s, dt_partition, tlatent, tpick, ipick, phase_label,
                     locs_use_cart, x_temp_cuda_cart, x_query_cart, x_query_src_cart, t_query, tq_sample, trv_out_q):
        """module.py:974-997: front end + heads + association branch -> (y, x, arv_p, arv_s)."""
        if not ops.AssocWeights.supported(self):
            raise NotImplementedError('genie_b200: the association kernels are built for the reference\'s module shapes')
        with torch.no_grad():
            x_spatial, x_latent, _ = self.front_end(Slice, Mask, x_temp_cuda_cart, want_latent=True,
                                                    locs_use_cart=locs_use_cart)
            y, x = self._heads(x_spatial, x_temp_cuda_cart, x_query_cart, t_query)
            x_src = self.SpatialAttention(x_spatial, x_query_src_cart, x_temp_cuda_cart, cache=False)         # :980
            if A_Lg_in_src is None and self._plan.mode == capi.GRAPH_CARTESIAN:
                attr = self._read_in_attr                      # set_adjacencies_cartesian: implicit read-out graph
            else:
                attr = self._check_read_out_graph(A_Lg_in_src)
            if A_edges_p is None or A_edges_s is None or dt_partition is None or tlatent is None:
                raise RuntimeError('forward_fixed needs A_edges_p, A_edges_s, dt_partition and tlatent (set_adjacencies)')
            dev = x_spatial.device
            if self._assoc_w is None or self._assoc_w.device != dev:
                self._assoc_w = ops.AssocWeights(dev)
            packed = self._assoc_w.update(self, self._update_assoc_terms(locs_use_cart, x_temp_cuda_cart))
            s_rows = ops.assoc_product_fwd(self._plan, packed, x_spatial, y.reshape(y.shape[0], -1), attr, x_latent, Mask,
                                           mask_thresh=0.01)                                                   # :983-987
            cp = self.LocalSliceLgCollapseP
            ph = phase_label.reshape(-1).float()
            if not cp.use_phase_types:
                ph = ph * 0.0
            arrival = ops.assoc_collapse_fwd(packed, s_rows, A_edges_p, A_edges_s, dt_partition, tlatent, tpick, ipick, ph,
                                             int(locs_use_cart.shape[0]), float(cp.eps))                       # :988-989
            arv = self.Arrivals.forward_merged(tq_sample, x_src, trv_out_q, arrival, tpick, ipick, phase_label)  # :990
        return y, x, arv[:, :, 0].unsqueeze(-1), arv[:, :, 1].unsqueeze(-1)

    def forward_fixed(self, Slice, Mask, tpick, ipick, phase_label, locs_use_cart, x_temp_cuda_cart, x_query_cart,
                      x_query_src_cart, t_query, tq_sample, trv_out_q):
        """module.py:963-997 -> (y [G,T,1], x [Q,T,1], arv_p [n_src,n_arv,1], arv_s [n_src,n_arv,1])."""
        if self._plan is None:
            raise RuntimeError('set_adjacencies must be called before forward_fixed*')
        return self._association(Slice, Mask, self.A_Lg_in_src, self.A_edges_p, self.A_edges_s, self.dt_partition,
                                 self.tlatent, tpick, ipick, phase_label, locs_use_cart, x_temp_cuda_cart, x_query_cart,
                                 x_query_src_cart, t_query, tq_sample, trv_out_q)

    def forward(self, Slice, Mask, A_in_sta, A_in_src, A_src_in_edges, A_Lg_in_src, A_src_in_sta, A_src, A_edges_p,
                A_edges_s, dt_partition, tlatent, tpick, ipick, phase_label, locs_use_cart, x_temp_cuda_cart,
                x_query_cart, x_query_src_cart, t_query, tq_sample, trv_out_q):
        """module.py:908-939: forward_fixed with the adjacencies passed on every call (plans are cached on tensor identity).
        Under torch.no_grad() the fused inference kernels run; with gradients enabled (training) the differentiable path of
        genie_b200/training.py does."""
        n_sta, n_grid = int(locs_use_cart.shape[0]), int(x_temp_cuda_cart.shape[0])
        self._plan_for(A_in_sta, A_in_src, A_src_in_edges, A_src, n_sta, n_grid, Slice.device,
                       pos=(locs_use_cart, x_temp_cuda_cart), A_src_in_sta=A_src_in_sta)
        self.A_src_in_sta = A_src_in_sta
        if torch.is_grad_enabled():
            # training (train_GENIE_model.py:1786): the differentiable path — the reference's operator graph with the
            # product-graph message passing (forward and backward) on libgenie_b200's gather kernel (genie_b200/training.py)
            from . import training
            return training.forward_train(self, Slice, Mask, A_Lg_in_src, A_src, A_edges_p, A_edges_s, dt_partition, tlatent,
                                          tpick, ipick, phase_label, locs_use_cart, x_temp_cuda_cart, x_query_cart,
                                          x_query_src_cart, t_query, tq_sample, trv_out_q)
        return self._association(Slice, Mask, A_Lg_in_src, A_edges_p, A_edges_s, dt_partition, tlatent, tpick, ipick,
                                 phase_label, locs_use_cart, x_temp_cuda_cart, x_query_cart, x_query_src_cart, t_query,
                                 tq_sample, trv_out_q)
