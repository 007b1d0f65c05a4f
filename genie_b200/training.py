"""Differentiable forward of `GCN_Detection_Network_extended` for training (module.py:908-939, `mz(*input_tensors)` at
train_GENIE_model.py:1786; BASELINE.json configs[2]).

The inference path (`forward_fixed_source`, `forward_fixed`) runs fused kernels that keep no activations.  Training needs
gradients with respect to every parameter, so this path keeps the reference's operator graph — and replaces the part of it
that costs the reference >= 85 % of its time: `MessagePassing.propagate` over the product graph, i.e. an
`index_select` that materialises an [E, 30] tensor (E = 15 P) followed by `scatter_add_`, in the forward AND in the backward
pass.  Here that pair is ONE gather kernel per direction (`genie_kron_spmm_fwd`: the product graph as the Kronecker product of
a small CSR matrix with an identity, forward by target, backward by source, no atomics), wrapped in `MeanAggregate`.  The dense
per-node Linear / PReLU layers stay torch ops (plain library GEMMs with autograd); the grid-sized SpatialAggregation layers and
the pick-sized heads / association modules are index ops on small tensors.

Nothing here touches the CPU: `ops.kron_spmm` raises on host tensors.
"""
import numpy as np
import torch

from . import capi, ops


class KronGraph(object):
    """One edge type of the product graph: forward CSR (by target, val = 1/in-degree) and its transpose (by source)."""

    def __init__(self, mode, n_sta, n_grid, n_prod, rowptr, col, device):
        self.mode, self.n_sta, self.n_grid, self.n_prod = int(mode), int(n_sta), int(n_grid), int(n_prod)
        rowptr = rowptr.to(device).long().contiguous()
        col = col.to(device).to(torch.int32).contiguous()
        n = rowptr.numel() - 1
        deg = rowptr[1:] - rowptr[:-1]
        tgt = torch.repeat_interleave(torch.arange(n, device=device), deg)
        w = 1.0 / deg.clamp(min=1).float()
        self.fwd = (rowptr, col, w[tgt].contiguous())
        order = torch.sort(col.long(), stable=True)[1]                  # edges grouped by source
        rrow = torch.zeros(n + 1, dtype=torch.long, device=device)
        rrow[1:] = torch.cumsum(torch.bincount(col.long(), minlength=n), 0)
        self.rev = (rrow.contiguous(), tgt[order].to(torch.int32).contiguous(), w[tgt][order].contiguous())


def build_kron_graphs(plan):
    """(station-edge graph, source-edge graph) of a GraphPlan: the small kNN graphs for CARTESIAN plans, the explicit
    product-level CSR otherwise."""
    dev = plan.device
    if plan.mode == capi.GRAPH_CARTESIAN:
        return (KronGraph(0, plan.n_sta, plan.n_grid, plan.n_prod, plan.sta_rowptr, plan.sta_col, dev),
                KronGraph(1, plan.n_sta, plan.n_grid, plan.n_prod, plan.src_rowptr, plan.src_col, dev))
    return (KronGraph(2, 0, 0, plan.n_prod, plan.sta_rowptr, plan.sta_col, dev),
            KronGraph(2, 0, 0, plan.n_prod, plan.src_rowptr, plan.src_col, dev))


class MeanAggregate(torch.autograd.Function):
    """`propagate(A, x=x)` with aggr='mean' and message = x_j (module.py:90-95), forward and backward on libgenie_b200."""

    @staticmethod
    def forward(ctx, x, kg):
        ctx.kg = kg
        return ops.kron_spmm(kg, kg.fwd, x)

    @staticmethod
    def backward(ctx, grad):
        return ops.kron_spmm(ctx.kg, ctx.kg.rev, grad), None


def mean_aggregate(x, kg):
    return MeanAggregate.apply(x.contiguous(), kg)


class _SplitKLinear(torch.autograd.Function):
    """nn.Linear on a product-node-sized input [P, n_in] with a weight gradient the GPU can parallelise.  The gradient
    dW = dY^T X is a [n_out x n_in] (at most 30 x 95) result reduced over P = 10^5..10^7 rows: as ONE GEMM the library
    maps it to a single output tile, i.e. one SM walks all P rows (measured: 409 us per layer at P = 5 * 10^5, 27 % of a training
    sample, profiles/r3j_train_profile.log).  Here the rows are cut into SPLIT slabs, the slabs' partial products run as one
    batched GEMM over all SMs and are summed (a fixed order: bit-reproducible)."""
    SPLIT = 256

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        n = x.shape[0]
        c = n // _SplitKLinear.SPLIT
        main = c * _SplitKLinear.SPLIT
        gw = None
        if ctx.needs_input_grad[1]:
            gw = torch.bmm(gy[:main].view(_SplitKLinear.SPLIT, c, -1).transpose(1, 2),
                           x[:main].view(_SplitKLinear.SPLIT, c, -1)).sum(0)
            if main < n:
                gw = gw + gy[main:].t() @ x[main:]
        gx = gy @ weight if ctx.needs_input_grad[0] else None
        gb = gy.sum(0) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


LIN_SPLIT_MIN_ROWS = 16384


def lin(layer, x):
    """`layer(x)` for an nn.Linear; product-node-sized inputs take the split-K weight gradient."""
    if x.dim() == 2 and x.shape[0] >= LIN_SPLIT_MIN_ROWS:
        return _SplitKLinear.apply(x.contiguous(), layer.weight, layer.bias)
    return layer(x)


class NodeMLP(torch.autograd.Function):
    """`activate(Linear(torch.cat(parts, dim=1)))` over product-node-sized tensors, forward and backward in ONE kernel each
    (genie_node_mlp_fwd / genie_node_mlp_bwd, csrc/mlp_kernels.cu): no concatenated tensor, no separate addmm / prelu /
    weight-gradient GEMM.  `slope` is the 1-element nn.PReLU weight, or None for a bare Linear."""

    @staticmethod
    def forward(ctx, weight, bias, slope, *parts):
        parts = tuple(p if (p.dtype == torch.float32 and p.stride(1) == 1) else p.float().contiguous() for p in parts)
        y, neg = ops.node_mlp_fwd(parts, weight, bias, slope)
        ctx.has_slope = slope is not None
        ctx.has_bias = bias is not None
        ctx.save_for_backward(weight, bias if bias is not None else weight.new_zeros(1),
                              slope if slope is not None else weight.new_zeros(1), y,
                              neg if neg is not None else weight.new_zeros(1, dtype=torch.int32), *parts)
        return y

    @staticmethod
    def backward(ctx, gy):
        weight, bias, slope, y, neg = ctx.saved_tensors[:5]
        parts = ctx.saved_tensors[5:]
        need_gx = list(ctx.needs_input_grad[3:])
        gx, gW, gb, ga = ops.node_mlp_bwd(parts, weight, bias if ctx.has_bias else None, slope if ctx.has_slope else None, y, neg,
                                          gy, need_gx)
        return (gW, gb if ctx.has_bias else None, ga.view_as(slope) if ctx.has_slope else None) + tuple(gx)


MLP_MIN_ROWS = 4096


def mlp(layer, act, *parts):
    """`act(layer(torch.cat(parts, dim=1)))` (act None: no activation).  Product-node-sized CUDA inputs of the module shapes
    go through the fused kernels; anything else (tiny inputs, oversize layers, a zero PReLU slope, whose gradient the fused
    backward cannot recover from the activation's output) through the torch ops."""
    w = layer.weight
    n = parts[0].shape[0]
    if n >= MLP_MIN_ROWS and w.is_cuda and ops.node_mlp_supported(parts, w) and not _ZERO_SLOPE.get(id(act), False):
        return NodeMLP.apply(w, layer.bias, act.weight if act is not None else None, *parts)
    x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    y = lin(layer, x)
    return act(y) if act is not None else y


_ZERO_SLOPE = {}


def note_zero_slopes(model):
    """One device read per call of forward_train: which nn.PReLU modules currently have slope 0 (fused backward not usable)."""
    acts = [m for m in model.modules() if isinstance(m, torch.nn.PReLU)]
    z = (torch.cat([a.weight.detach().reshape(-1)[:1] for a in acts]) == 0).cpu().numpy()
    _ZERO_SLOPE.clear()
    for a, f in zip(acts, z):
        if f:
            _ZERO_SLOPE[id(a)] = True


# ---- the modules of module.py as differentiable functions of the parameter holders ----------------------------------------

def _expand_edge_means(model, kg_sta):
    """Updated model: mean over in-edges of pos_rel for every product node ([P,4] each), from the per-station / per-grid-node
    (CARTESIAN) or per-node (EXPLICIT) tables of `model._edge_means`."""
    m_sta, m_src = model._edge_means[0](float(model.scale_rel)), model._edge_means[1](float(model.scale_rel))
    if kg_sta.mode == 2:
        return m_sta, m_src
    return m_sta.repeat(kg_sta.n_grid, 1), m_src.repeat_interleave(kg_sta.n_sta, dim=0)


def _agg_layer(da, l_a, l_b, act, tr, msg_a, msg_b, kg_sta, kg_src, tail_sta, tail_src):
    """[l_a([tr | mean_sta msg_a | tail]) | l_b([tr | mean_src msg_b | tail])] with `act` applied (one slope: applying it to
    the halves is applying it to the concatenation, module.py:92, :96)."""
    t1 = mlp(l_a, act, tr, mean_aggregate(msg_a, kg_sta), tail_sta)
    t2 = mlp(l_b, act, tr, mean_aggregate(msg_b, kg_src), tail_src)
    return torch.cat((t1, t2), dim=1)


def data_aggregation(da, Slice, Mask, kg_sta, kg_src, edge_means=None):
    """DataAggregation.forward (module.py:85-98) / DataAggregationEdges.forward (:143-157)."""
    tr = mlp(da.init_trns, da.activate, Slice, Mask)
    if edge_means is not None:          # the edge-feature channels sit between the aggregate and the mask (:150-156)
        tail_sta, tail_src = torch.cat((edge_means[0], Mask), dim=1), torch.cat((edge_means[1], Mask), dim=1)
    else:
        tail_sta = tail_src = Mask
    tr = _agg_layer(da, da.l1_t1_2, da.l1_t2_2, da.activate1, tr, da.activate11(tr), da.activate12(tr), kg_sta, kg_src,
                    tail_sta, tail_src)
    return _agg_layer(da, da.l2_t1_2, da.l2_t2_2, da.activate2, tr, mlp(da.l2_t1_1, da.activate21, tr),
                      mlp(da.l2_t2_1, da.activate22, tr), kg_sta, kg_src, tail_sta, tail_src)


def data_aggregation_association(da, s, latent, mask1, mask2, kg_sta, kg_src, edge_means=None):
    """DataAggregationAssociationPhase.forward (module.py:387-403) / ...Edges.forward (:442-467)."""
    mask = torch.cat((mask1, mask2), dim=-1)
    tr = mlp(da.init_trns, da.activate, s, latent, mask)
    if edge_means is not None:
        tail_sta, tail_src = torch.cat((edge_means[0], mask), dim=1), torch.cat((edge_means[1], mask), dim=1)
    else:
        tail_sta = tail_src = mask
    tr = _agg_layer(da, da.l1_t1_2, da.l1_t2_2, da.activate1, tr, mlp(da.l1_t1_1, da.activate11, tr),
                    mlp(da.l1_t2_1, da.activate12, tr), kg_sta, kg_src, tail_sta, tail_src)
    return _agg_layer(da, da.l2_t1_2, da.l2_t2_2, da.activate2, tr, mlp(da.l2_t1_1, da.activate21, tr),
                      mlp(da.l2_t2_1, da.activate22, tr), kg_sta, kg_src, tail_sta, tail_src)


def bipartite_read_in(ri, x_latent, attr, node_grid, n_grid, Mask):
    """BipartiteGraphOperator.forward (module.py:224-229): masked per-node MLP summed onto the node's grid node."""
    h = Mask.max(1, keepdim=True)[0] * mlp(ri.fc1, ri.activate1, x_latent, attr)
    xg = h.new_zeros((n_grid, h.shape[1])).index_add_(0, node_grid, h)
    return ri.activate2(ri.fc2(xg))


def spatial_aggregation(sa, x, A_src, pos, scale_rel):
    """SpatialAggregation.forward / message (module.py:243-249); the global feature is a mean over EDGES (:249)."""
    n = x.shape[0]
    p = pos / scale_rel
    src, tgt = A_src[0], A_src[1]
    xj = x[src]
    glob = sa.activate3(sa.fglobal(xj)).mean(0, keepdim=True)
    msg = sa.activate1(sa.fc1(torch.cat((xj, p[tgt] - p[src], glob.expand(xj.shape[0], -1)), dim=-1)))
    cnt = torch.bincount(tgt, minlength=n).clamp(min=1).to(msg.dtype).unsqueeze(1)
    agg = msg.new_zeros((n, msg.shape[1])).index_add_(0, tgt, msg) / cnt
    return sa.activate2(sa.fc2(torch.cat((x, agg), dim=-1)))


def bipartite_read_out(ro, y_latent, attr, node_grid, mask_out):
    """BipartiteGraphReadOutOperator.forward (module.py:344-352) for the read-out graph [g(i); i]."""
    mj = mask_out[node_grid]
    G = y_latent.shape[0]
    if node_grid.numel() % max(G, 1) == 0 and node_grid.numel() and getattr(node_grid, '_genie_regular', False):
        # dense mode: node_grid = repeat_interleave(arange(G), S) — an expand, whose gradient is a reshape + sum over the stations
        # (the fancy-index gather's backward is an index_put with accumulation: 1.7 ms per sample at 100 x 5000)
        yl = y_latent.unsqueeze(1).expand(G, node_grid.numel() // G, y_latent.shape[1]).reshape(-1, y_latent.shape[1])
    else:
        yl = y_latent[node_grid]
    h = mj * mlp(ro.fc1, ro.activate1, yl, attr)
    return mlp(ro.fc2, ro.activate2, h), mj


def local_slice_collapse(cm, A_edges, dt_partition, tpick, ipick, phase_label, s, tlatent, k_infer=10):
    """LocalSliceLgCollapse.forward / message (module.py:624-653)."""
    n_arv, l_dt = tpick.shape[0], dt_partition.shape[0]
    dev = s.device
    dt = dt_partition[1] - dt_partition[0]
    ph = phase_label.reshape(-1, 1).float()
    if not cm.use_phase_types:
        ph = ph * 0.0
    t_index = torch.floor((tpick - dt_partition[0]) / dt).long()
    t_index = ((ipick * l_dt * k_infer + t_index * k_infer).view(-1, 1) + torch.arange(k_infer, device=dev).view(1, -1)).reshape(-1)
    e1 = torch.arange(n_arv, device=dev).repeat_interleave(k_infer)
    e0 = A_edges[t_index].long()
    keep = torch.nonzero((tpick[e1] - tlatent[e0, 0]).abs() < 2.0 * cm.eps, as_tuple=True)[0]
    e0, e1 = e0[keep], e1[keep]
    msg = cm.activate1(cm.fc1(torch.cat((s[e0], (tpick.view(-1, 1)[e1] - tlatent[e0]) / cm.eps, ph[e1]), dim=-1)))
    cnt = torch.bincount(e1, minlength=n_arv).clamp(min=1).to(msg.dtype).unsqueeze(1)
    agg = msg.new_zeros((n_arv, msg.shape[1])).index_add_(0, e1, msg) / cnt
    return cm.activate2(cm.fc2(agg))


def forward_train(model, Slice, Mask, A_Lg_in_src, A_src, A_edges_p, A_edges_s, dt_partition, tlatent, tpick, ipick,
                  phase_label, locs_use_cart, x_temp_cuda_cart, x_query_cart, x_query_src_cart, t_query, tq_sample, trv_out_q):
    """module.py:908-939 with autograd: returns (y, x, arv_p, arv_s).  `model._plan` must be the plan of the call's graphs."""
    plan = model._plan
    if not Slice.is_cuda:
        raise capi.GenieError('genie_b200 has no CPU path: inputs must be CUDA tensors')
    key = id(plan)
    if getattr(model, '_kron', None) is None or model._kron[0] != key:
        model._kron = (key,) + build_kron_graphs(plan)
    kg_sta, kg_src = model._kron[1], model._kron[2]
    node_grid = plan.node_grid_index()
    if plan.mode == capi.GRAPH_CARTESIAN:
        node_grid._genie_regular = True
    note_zero_slopes(model)
    Slice, Mask = Slice.float(), Mask.float()
    scale = float(model.scale_rel)
    edge_means = _expand_edge_means(model, kg_sta) if model.updated_model else None
    abs_ch = None
    if model.use_absolute_pos:                                                                               # :913-914
        if plan.mode == capi.GRAPH_CARTESIAN:
            abs_ch = torch.cat((locs_use_cart.repeat(plan.n_grid, 1), x_temp_cuda_cart.repeat_interleave(plan.n_sta, dim=0)),
                               dim=1) / (3.0 * scale)
        else:
            idx = model.A_src_in_sta.to(Slice.device).long()
            abs_ch = torch.cat((locs_use_cart[idx[0]], x_temp_cuda_cart[idx[1]]), dim=1) / (3.0 * scale)
        Slice = torch.cat((Slice, abs_ch), dim=1)
    attr = model._read_in_attr
    x_latent = data_aggregation(model.DataAggregation, Slice, Mask, kg_sta, kg_src, edge_means)
    x = bipartite_read_in(model.Bipartite_ReadIn, x_latent, attr, node_grid, plan.n_grid, Mask)
    A_src = A_src.to(x.device).long()
    x = spatial_aggregation(model.SpatialAggregation1, x, A_src, x_temp_cuda_cart, scale)
    x = spatial_aggregation(model.SpatialAggregation2, x, A_src, x_temp_cuda_cart, scale)
    x_spatial = spatial_aggregation(model.SpatialAggregation3, x, A_src, x_temp_cuda_cart, scale)
    y_latent = model.SpatialDirect(x_spatial)
    y = model.TemporalAttention(y_latent, t_query)
    x = model.SpatialAttention(x_spatial, x_query_cart, x_temp_cuda_cart)
    x_src = model.SpatialAttention(x_spatial, x_query_src_cart, x_temp_cuda_cart, cache=False)
    x = model.TemporalAttention(x, t_query)
    mask_out = 1.0 * (y[:, :, 0].detach().max(1, keepdim=True)[0] > 0.01).detach()                           # :926
    ro_attr = attr if A_Lg_in_src is None else model._check_read_out_graph(A_Lg_in_src)
    s, mask_out_1 = bipartite_read_out(model.BipartiteGraphReadOutOperator, y_latent, ro_attr, node_grid, mask_out)
    if abs_ch is not None:
        s = torch.cat((s, abs_ch), dim=1)                                                                    # :930-931
    s = data_aggregation_association(model.DataAggregationAssociationPhase, s, x_latent.detach(), mask_out_1, Mask, kg_sta,
                                     kg_src, edge_means)                                                     # :932
    dtp, tl = dt_partition.to(s.device).float(), tlatent.to(s.device).float()
    arv_p = local_slice_collapse(model.LocalSliceLgCollapseP, A_edges_p.to(s.device), dtp, tpick, ipick, phase_label, s,
                                 tl[:, 0].reshape(-1, 1))
    arv_s = local_slice_collapse(model.LocalSliceLgCollapseS, A_edges_s.to(s.device), dtp, tpick, ipick, phase_label, s,
                                 tl[:, 1].reshape(-1, 1))
    arv = model.Arrivals(x_query_src_cart, tq_sample, x_src, trv_out_q, locs_use_cart, arv_p, arv_s, tpick, ipick, phase_label)
    return y, x, arv[:, :, 0].unsqueeze(-1), arv[:, :, 1].unsqueeze(-1)
