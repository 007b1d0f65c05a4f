"""Oracle check of a FULL-SIZE window on a sampled sub-network (TEST INFRASTRUCTURE — never on the product path).

At BASELINE.json's headline size (1000 stations x 50000 grid nodes, P = 5e7 product nodes) the CPU oracle cannot run the
whole window (one message tensor is 90 GB).  But the value of DataAggregation at a product node (g, s) depends only on
  * all stations of grid node g                                      (station edges keep the grid node, process_utils.py:720)
  * the 2-hop in-neighbourhood of g in the source graph              (source edges keep the station, :721; two aggregations,
                                                                      module.py:90-95)
so for a sample T of grid nodes the oracle is run on the sub-network  (all stations) x C2,
      C1 = T u N_src(T),   C2 = C1 u N_src(C1),
with the full-size source graph restricted to the edges whose TARGET lies in C1 (their sources are in C2 by construction).
Rows of T then equal the full-size result of the reference algorithm: time_bin (integer, exact), Slice / Mask, x_latent
(module.py:85-98) and the Bipartite_ReadIn rows (module.py:224-229).  Given a full-size read-in table the rest of the
window — SpatialAggregation x3 and the read-out heads, G x 15 edges — is cheap enough to run on the full grid.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (outside its timed regions) import this file.
"""
import numpy as np
import torch

from . import genie_oracle as go


def sample_targets(G, n_clusters=5, cluster=4):
    """Grid nodes to check: `n_clusters` runs of `cluster` consecutive ids (consecutive ids are spatial neighbours when the
    grid is sorted along a space-filling curve, so their closures overlap) spread over the whole id range."""
    starts = np.unique(np.linspace(0, max(G - cluster, 0), n_clusters).astype(np.int64))
    t = np.unique(np.concatenate([np.arange(s, min(G, s + cluster)) for s in starts]))
    return t.astype(np.int64)


def in_neighbours(A_src_src, nodes, G):
    """Union of `nodes` and the sources of every edge of A_src_src ([2,E], row 0 source, row 1 target) ending in `nodes`."""
    A = A_src_src.cpu().numpy() if torch.is_tensor(A_src_src) else np.asarray(A_src_src)
    mark = np.zeros(G, dtype=bool)
    mark[nodes] = True
    return np.unique(np.concatenate((np.asarray(nodes, dtype=np.int64), A[0][mark[A[1]]].astype(np.int64))))


def closure(A_src_src, targets, G):
    """(C1, C2, sub-graph [2,E'] in C2-local ids with every in-edge of C1, positions of the targets inside C2)."""
    A = A_src_src.cpu().numpy() if torch.is_tensor(A_src_src) else np.asarray(A_src_src)
    c1 = in_neighbours(A, targets, G)
    c2 = in_neighbours(A, c1, G)
    local = -np.ones(G, dtype=np.int64)
    local[c2] = np.arange(len(c2))
    in_c1 = np.zeros(G, dtype=bool)
    in_c1[c1] = True
    keep = in_c1[A[1]]                                 # edge order of the full list is kept: same summation order
    sub = np.stack((local[A[0][keep]], local[A[1][keep]]), axis=0)
    assert sub.min() >= 0
    return c1, c2, torch.from_numpy(sub).long(), local[np.asarray(targets, dtype=np.int64)]


def rowwise_rel(a, b):
    """Element-wise relative metric: max over rows of  max_j |a_ij - b_ij| / max_j |b_ij|  (every element is held to the
    tolerance times the larger of its own magnitude and its row's scale) — stricter than conftest.rel_err, which divides by
    the maximum of the whole tensor."""
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    scale = np.maximum(np.abs(b).max(axis=1), 1e-30)
    return float((np.abs(a - b).max(axis=1) / scale).max())


def global_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def oracle_on_closure(sd, A_sta_sta, A_src_src, S, G, targets, picks, t0, trv_of, attr_of, max_t, kernel_sig_t, dt):
    """Runs the oracle's a1 + a2 + a3 on the closure of `targets`.

    picks: float64 [n,5] (any superset of the window's picks); trv_of(nodes) -> fp32 [len(nodes), S, 2] travel times and
    attr_of(nodes) -> fp32 [len(nodes) * S, 3] read-in edge features of the given grid nodes (full-size ids).
    Returns dict(targets, nodes (product-node ids g*S+s of the target rows, full-size numbering), time_bin [n,2] int64,
    slice, mask [n,4], x_latent [n,30], read_in [len(targets),15], n_closure)."""
    targets = np.asarray(targets, dtype=np.int64)
    c1, c2, A_sub, t_loc = closure(A_src_src, targets, G)
    n2 = len(c2)
    A_sta = A_sta_sta.cpu() if torch.is_tensor(A_sta_sta) else torch.from_numpy(np.asarray(A_sta_sta))
    A_sta = A_sta.long()
    # product lists of the sub-network by the reference's own patterns (process_utils.py:720-722)
    A_prod_sta = (A_sta.repeat(1, n2) + S * torch.arange(n2).repeat_interleave(A_sta.shape[1]).view(1, -1)).contiguous()
    A_prod_src = (S * A_sub.repeat(1, S) + torch.arange(S).repeat_interleave(A_sub.shape[1]).view(1, -1)).contiguous()
    A_src_in_prod = torch.stack((torch.arange(S * n2), torch.arange(n2).repeat_interleave(S)), dim=0)
    A_src_in_sta = np.stack((np.tile(np.arange(S), n2), np.repeat(np.arange(n2), S)), axis=0)
    trv = np.ascontiguousarray(trv_of(c2), dtype=np.float32)
    Sl, Mk, parts = go.input_scatter(picks, t0, np.arange(S), S, A_src_in_sta, trv, max_t, kernel_sig_t, dt,
                                     return_parts=True)
    Slice, Mask = torch.from_numpy(Sl), torch.from_numpy(Mk)
    with torch.no_grad():
        x_latent = go.data_aggregation(sd, 'DataAggregation.', Slice, Mask, A_prod_sta, A_prod_src)
        attr = torch.from_numpy(np.ascontiguousarray(attr_of(c2), dtype=np.float32))
        r = go.bipartite_read_in(sd, 'Bipartite_ReadIn.', x_latent, attr, A_src_in_prod, Mask)
    rows = (t_loc[:, None] * S + np.arange(S)[None, :]).reshape(-1)
    nodes = (targets[:, None] * S + np.arange(S)[None, :]).reshape(-1)
    return dict(targets=targets, nodes=nodes, time_bin=parts['time_bin'][rows], slice=Sl[rows], mask=Mk[rows],
                x_latent=x_latent.numpy()[rows], read_in=r.numpy()[t_loc], n_closure=int(n2), n_c1=int(len(c1)))


def compare(want, got_time_bin, got_slice, got_mask, got_latent, got_read_in):
    """`got_*`: the CUDA path's rows for want['nodes'] / want['targets'] (numpy).  Returns the report bench.py prints."""
    rep = dict(nodes=int(len(want['targets'])), product_nodes=int(len(want['nodes'])), closure_grid_nodes=want['n_closure'])
    if got_time_bin is not None:
        rep['time_bin_equal'] = bool(np.array_equal(np.asarray(got_time_bin, dtype=np.int64), want['time_bin']))
    if got_slice is not None:
        rep['slice_max_abs'] = float(np.abs(got_slice - want['slice']).max())
        rep['mask_equal'] = bool(np.array_equal(got_mask, want['mask']))
    rep['x_latent_rel'] = global_rel(got_latent, want['x_latent'])
    rep['x_latent_rowwise_rel'] = rowwise_rel(got_latent, want['x_latent'])
    rep['read_in_rel'] = global_rel(got_read_in, want['read_in'])
    rep['read_in_rowwise_rel'] = rowwise_rel(got_read_in, want['read_in'])
    rep['max_rel'] = max(rep['x_latent_rowwise_rel'], rep['read_in_rowwise_rel'])
    return rep


def oracle_tail(sd, read_in, A_src, grid_cart, x_query_cart, t_query, scale_rel, scale_t):
    """SpatialAggregation x3 + read-out heads (module.py:1012-1020) on the FULL grid from a full-size read-in table [G,15]:
    G x 15 edges, cheap on the CPU.  Returns y [G,T,1], x [Q,T,1]."""
    with torch.no_grad():
        r = torch.as_tensor(read_in).float()
        A_src = A_src.cpu().long() if torch.is_tensor(A_src) else torch.from_numpy(np.asarray(A_src)).long()
        x1 = go.spatial_aggregation(sd, 'SpatialAggregation1.', r, A_src, grid_cart, scale_rel)
        x2 = go.spatial_aggregation(sd, 'SpatialAggregation2.', x1, A_src, grid_cart, scale_rel)
        x3 = go.spatial_aggregation(sd, 'SpatialAggregation3.', x2, A_src, grid_cart, scale_rel)
        y = go.temporal_attention(sd, 'TemporalAttention.', go.spatial_direct(sd, 'SpatialDirect.', x3), t_query, scale_t)
        xq = go.spatial_attention(sd, 'SpatialAttention.', x3, x_query_cart, grid_cart, scale_rel)
        x = go.temporal_attention(sd, 'TemporalAttention.', xq, t_query, scale_t)
    return y.numpy(), x.numpy()
