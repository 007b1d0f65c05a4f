"""CPU oracle for the GENIE product-graph front end (TEST INFRASTRUCTURE — never on the product path).

A functional restatement, in plain torch-CPU fp32 / numpy fp64, of the algorithm of the reference's hot path
(SURVEY.md §8a rows a1–a5).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this file; `genie_b200/` never does.

Pinning: every function below is checked (tests/test_oracle_golden.py) against `tests/golden/*.npz`, which were produced
by `oracle/gen_golden.py` running the UNMODIFIED reference classes (`/root/reference/Code/module.py`,
`process_utils.py`) in the build container through the import stubs in `oracle/refshim/`.  The reference ships no tests
or golden vectors of its own for this path, and its arithmetic lives in un-vendored, un-pinned
third-party wheels (torch_geometric `MessagePassing.propagate`, torch_scatter `scatter`, torch_cluster `knn`;
`Code/install_dependencies.txt:9-17`), whose documented semantics are restated in `propagate_*` / `scatter_max` / `knn`
below.  So: parity is pinned against the reference's own Python classes executed here, with the third-party message
passing semantics restated — not against a reference-held golden vector (none exists).

Conventions: edge lists are int64 `[2, E]`, row 0 = message source j, row 1 = target i; product node id
= g*S + s in dense mode; all floating point tensors fp32; the pick → time-bin index map is computed in fp64 and
truncated toward zero exactly as numpy does.

State is passed as a flat dict with the reference's own state_dict key names
(`DataAggregation.init_trns.weight`, ..., every `nn.PReLU` a 1-element `weight`).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------------------------------


def _lin(sd, name, x):
    """nn.Linear: x @ W^T + b."""
    return torch.nn.functional.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def _prelu(sd, name, x):
    """nn.PReLU with a single slope."""
    a = sd[name + '.weight'].reshape(())
    return torch.where(x >= 0, x, a * x)


def propagate_sum(msg, index, n_out):
    """PyG aggr='add': out[i] = sum_{e: index[e]=i} msg[e]  (index_add_, the op PyG-CPU dispatches to)."""
    out = torch.zeros((n_out,) + tuple(msg.shape[1:]), dtype=msg.dtype)
    return out.index_add_(0, index, msg)


def propagate_mean(msg, index, n_out):
    """PyG aggr='mean': sum / max(count, 1); a node with no in-edges gets 0."""
    s = propagate_sum(msg, index, n_out)
    cnt = torch.zeros(n_out, dtype=msg.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=msg.dtype))
    return s / cnt.clamp(min=1).view((-1,) + (1,) * (msg.dim() - 1))


def segment_softmax(src, index, n_seg):
    """torch_geometric.utils.softmax over groups of equal `index` (dim 0)."""
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    mx = torch.full((n_seg,) + tuple(src.shape[1:]), float('-inf'), dtype=src.dtype)
    mx = mx.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    ex = (src - mx.gather(0, idx)).exp()
    den = torch.zeros((n_seg,) + tuple(src.shape[1:]), dtype=src.dtype).scatter_add_(0, idx, ex)
    return ex / (den.gather(0, idx) + 1e-16)


def knn(x, y, k):
    """torch_cluster.knn(x, y, k): for each row of y the k nearest rows of x; [2, |y|k], row0 = y idx, row1 = x idx."""
    from scipy.spatial import cKDTree
    xn = np.asarray(x, dtype=np.float64)
    yn = np.asarray(y, dtype=np.float64)
    k = int(min(k, xn.shape[0]))
    ind = cKDTree(xn).query(yn, k=k)[1].reshape(yn.shape[0], k)
    return torch.from_numpy(np.stack((np.repeat(np.arange(yn.shape[0]), k), ind.reshape(-1)), axis=0)).long()


def knn_graph_no_self(pos_km, k):
    """`remove_self_loops(knn(x, x, k + 1).flip(0))` — process_utils.py:718-719.  Row 0 = source j, row 1 = target i."""
    e = knn(pos_km, pos_km, k + 1).flip(0).contiguous()
    return e[:, e[0] != e[1]]


# --------------------------------------------------------------------------------------------------------------------
# graph assembly (process_utils.py:701-742, dense Cartesian-product mode)
# --------------------------------------------------------------------------------------------------------------------


def build_adjacencies_dense(sta_cart_m, grid_cart_m, k_sta, k_spc):
    """Restates extract_inputs_adjacencies (process_utils.py:712-722) for Cartesian coordinates in metres.

    Returns A_sta_sta [2,S*k_s], A_src_src [2,G*k_g], A_prod_sta_sta, A_prod_src_src, A_src_in_prod, A_src_in_sta.
    """
    S, G = sta_cart_m.shape[0], grid_cart_m.shape[0]
    k_sta = int(min(k_sta, S - 2))                                              # :712
    A_sta = knn_graph_no_self((np.asarray(sta_cart_m, dtype=np.float64) / 1000.0).astype(np.float32), k_sta)   # :718
    A_src = knn_graph_no_self((np.asarray(grid_cart_m, dtype=np.float64) / 1000.0).astype(np.float32), k_spc)  # :719
    # :720  every grid node g carries a copy of the station graph, shifted by S*g
    A_prod_sta = (A_sta.repeat(1, G) + S * torch.arange(G).repeat_interleave(A_sta.shape[1]).view(1, -1)).contiguous()
    # :721  every station s carries a copy of the grid graph, node (g, s) = S*g + s
    A_prod_src = (S * A_src.repeat(1, S) + torch.arange(S).repeat_interleave(A_src.shape[1]).view(1, -1)).contiguous()
    # :722  product node -> its grid node
    A_src_in_prod = torch.stack((torch.arange(S * G), torch.arange(G).repeat_interleave(S)), dim=0).contiguous()
    # process_continuous_days.py:629  (station, grid) of every product node
    A_src_in_sta = torch.stack((torch.arange(S).repeat(G), torch.arange(G).repeat_interleave(S)), dim=0).contiguous()
    return A_sta, A_src, A_prod_sta, A_prod_src, A_src_in_prod, A_src_in_sta


# --------------------------------------------------------------------------------------------------------------------
# a1 — pick window -> Slice / Mask (process_utils.py:460-642, use_sign_input False, trv_times given)
# --------------------------------------------------------------------------------------------------------------------


def input_time_axis(t0, max_t, kernel_sig_t, dt):
    """abs_time_ref of process_utils.py:502 and its length; fp64, numpy arange semantics (start + i*dt)."""
    t0 = float(t0)
    t_offset = 3.0 * kernel_sig_t                                              # :500
    ref = np.arange(t0 - t_offset, t0 + max_t + t_offset + dt, dt)             # :502
    return ref, int(len(ref))


def input_scatter(P, t0, ind_use, n_locs, A_src_in_sta, trv_times, max_t, kernel_sig_t, dt, return_parts=False,
                  use_sign_input=False, trv_node=None):
    """Restates extract_input_from_data (process_utils.py:460-629).

    P [n,5] float64 (time, station, amp, prob, phase 0/1); ind_use int array of used (absolute) station ids;
    A_src_in_sta int [2,P] (row 0 = index into ind_use, row 1 = grid node); trv_times fp32 [G, n_locs, 2].
    Returns Slice [P,4] fp32, Mask [P,4] fp32 (and, with return_parts, the per-station series and the integer maps).

    Stations without a pick in the window read zeros in the reference (they are left out of `ind_unique`, :490, and of
    `ifind`, :586); here every used station owns a (possibly all-zero) series, which gives the same values.
    """
    P = np.asarray(P, dtype=np.float64)
    t0 = float(t0)
    ind_use = np.asarray(ind_use).astype('int')
    A = np.asarray(A_src_in_sta).astype('int')
    S_use = len(ind_use)
    ref, n_ts = input_time_axis(t0, max_t, kernel_sig_t, dt)
    # use_sign_input (:610-614): every feature times sign(-diff) of the flattened [station][bin] series it is read from.
    # trv_node [P,2]: per-node travel times as `trv_pairwise` returns them when trv_times is None (:594-596).

    keep = (P[:, 0] > (t0 - 2.0 * kernel_sig_t)) & (P[:, 0] < (t0 + max_t + 2.0 * kernel_sig_t))   # :476
    Pw = P[keep]
    perm = -1 * np.ones(n_locs, dtype='int')
    perm[ind_use] = np.arange(S_use)                                                               # :485-486
    sta_loc = perm[Pw[:, 1].astype('int')]
    Pw, sta_loc = Pw[sta_loc > -1], sta_loc[sta_loc > -1]                                          # :480-482

    n_extra = np.ceil(3 * kernel_sig_t / dt)                                                       # :520
    offs = np.arange(-n_extra, n_extra + 1).astype('int')                                          # :521
    series = np.zeros((2, S_use * n_ts), dtype=np.float32)
    for ph in (0, 1):
        sel = np.where(Pw[:, 4] == ph)[0]                                                          # :508-509
        near = ((Pw[sel, 0] - ref[0]) / dt).astype('int')                                          # :515-516
        idx = near.reshape(-1, 1) + offs.reshape(1, -1)                                            # :535
        ok = (idx >= 0) & (idx < n_ts)                                                             # :538
        idx = np.minimum(np.maximum(0, idx), n_ts - 1)                                             # :541
        dtv = Pw[sel, 0].reshape(-1, 1).repeat(len(offs), axis=1) - ref[idx]                       # :544
        vals = (ok * np.exp(-0.5 * (dtv ** 2) / (kernel_sig_t ** 2))).reshape(-1)                  # :546
        w = (idx + sta_loc[sel].reshape(-1, 1) * n_ts).reshape(-1)                                 # :557
        np.maximum.at(series[ph], w, vals.astype(np.float32))                                      # :563 scatter-max, 0 fill
    series = series.reshape(2, S_use, n_ts)
    series[:, :, 0] = 0.0                                                                          # :565-568
    series[:, :, n_ts - 1] = 0.0
    either = series.max(axis=0)                                                                    # :569

    # :599  fp32 travel time + fp64 t0 - fp64 ref[0], divided by fp64 dt, truncated toward zero
    tt = trv_times[A[1], ind_use[A[0]], :] if trv_node is None else np.asarray(trv_node)
    tb = ((tt + np.array([t0]) - ref[0]) / dt).astype('int')
    inb = (tb >= 0) & (tb < n_ts)
    tbc = np.clip(tb, 0, n_ts - 1)
    s = A[0]
    f0 = np.where(inb[:, 0], either[s, tbc[:, 0]], 0.0)                                            # :605
    f1 = np.where(inb[:, 1], either[s, tbc[:, 1]], 0.0)                                            # :606
    f2 = np.where(inb[:, 0], series[0, s, tbc[:, 0]], 0.0)                                         # :607
    f3 = np.where(inb[:, 1], series[1, s, tbc[:, 1]], 0.0)                                         # :608
    if use_sign_input:
        def slope_sign(e):                           # torch.sign(-diff(e, append = e[-1] + (e[-1] - e[-2]))) on the flat array
            flat = e.reshape(-1)
            d = np.diff(flat, append=flat[-1:] + (flat[-1:] - flat[-2:-1]))
            return np.sign(-1.0 * d).reshape(e.shape)
        sg_e, sg_p, sg_s = slope_sign(either), slope_sign(series[0]), slope_sign(series[1])
        f0 = f0 * np.where(inb[:, 0], sg_e[s, tbc[:, 0]], 0.0)                                     # :611
        f1 = f1 * np.where(inb[:, 1], sg_e[s, tbc[:, 1]], 0.0)                                     # :612
        f2 = f2 * np.where(inb[:, 0], sg_p[s, tbc[:, 0]], 0.0)                                     # :613
        f3 = f3 * np.where(inb[:, 1], sg_s[s, tbc[:, 1]], 0.0)                                     # :614
    Slice = np.stack((f0, f1, f2, f3), axis=1).astype(np.float32)                                  # :627-628
    Mask = (np.abs(Slice) > 0.01).astype(np.float32)                                               # :629
    if return_parts:
        return Slice, Mask, dict(series=series, time_bin=tb.astype(np.int64), n_ts=n_ts, ref0=float(ref[0]),
                                 n_picks=int(len(Pw)))
    return Slice, Mask


def legacy_pick_window(arrivals, time_samples, max_t, t_win):
    """process_utils.py:137-139: picks within t_win + max_t/2 of the centre of every sample's move-out window (the reference
    asks a cKDTree on the pick times; a closed ball of radius r is |t - c| <= r)."""
    return [np.where(np.abs(arrivals[:, 0] - (ts + max_t / 2.0)) <= t_win + max_t / 2.0)[0] for ts in time_samples]


def legacy_input_features(arrivals, phase_labels, ind_use, time_samples, x_grid_trv, max_t, t_win, kernel_sig_t,
                          return_parts=False):
    """a1' — extract_inputs_from_data_fixed_grids_with_phase_type (process_utils.py:102-308): per product node and phase the
    distance from the predicted arrival time to the NEAREST pick (any phase: channels 0, 1; same phase: channels 2, 3),
    through exp(-0.5 d^2 / sigma^2).  The reference lays all stations and samples on ONE sorted time axis (offsets of
    1.5 max_t per sample and 1.5 n_batch 1.5 max_t per station, :177-183) and looks at the two searchsorted neighbours
    (:197-207); that construction is kept literally, including what it does at the ends of the axis.
    Returns ([Inpts], [Masks]) as float64 [G * n_sta_use, 4] per sample (grid-major, stations in np.unique(ind_use) order)."""
    n_batch = len(time_samples)
    n_spc = x_grid_trv.shape[0]
    lp = legacy_pick_window(arrivals, time_samples, max_t, t_win)
    ind_sta_select = np.unique(ind_use)                                                                   # :152
    n_node = n_spc * len(ind_sta_select)
    offset_per_batch = 1.5 * max_t                                                                        # :177
    offset_per_station = 1.5 * n_batch * offset_per_batch                                                 # :178
    arrivals_offset = np.hstack([-time_samples[i] + i * offset_per_batch + offset_per_station * arrivals[lp[i], 1]
                                 for i in range(n_batch)])                                                # :180
    t_sel = np.hstack([arrivals[lp[i], 0] for i in range(n_batch)]) + arrivals_offset                     # :182
    ph_sel = np.hstack([phase_labels[lp[i]] for i in range(n_batch)])
    order = np.argsort(t_sel)                                                                             # :188
    t_sel, ph_sel = t_sel[order], ph_sel[order]
    n_arvs = len(t_sel)
    sta_col = np.tile(ind_sta_select, n_spc).astype(np.float64)                                           # :156
    batch = np.repeat(np.arange(n_batch), n_node).astype(np.float64)
    q = [np.tile(x_grid_trv[:, ind_sta_select, ph].reshape(-1).astype(np.float64), n_batch) + batch * offset_per_batch
         + np.tile(sta_col, n_batch) * offset_per_station for ph in (0, 1)]                               # :194-195

    def nearest(times, query):                                                                            # :197-207
        ip = np.searchsorted(times, query)
        pad = np.minimum(np.maximum(ip.reshape(-1, 1) + np.array([-1, 0]).reshape(1, -1), 0), len(times) - 1)
        return np.abs(query[:, np.newaxis] - times[pad]).min(1)

    feats = np.zeros((n_batch * n_node, 4))
    if n_arvs > 0:
        feats[:, 0] = np.exp(-0.5 * (nearest(t_sel, q[0]) ** 2) / (kernel_sig_t ** 2))                    # :254
        feats[:, 1] = np.exp(-0.5 * (nearest(t_sel, q[1]) ** 2) / (kernel_sig_t ** 2))
    for ph in (0, 1):                                                                                     # :209-222
        tp = t_sel[ph_sel == ph]
        if len(tp) > 0:
            feats[:, 2 + ph] = np.exp(-0.5 * (nearest(tp, q[ph]) ** 2) / (kernel_sig_t ** 2))             # :258-261
    Inpts = [feats[i * n_node:(i + 1) * n_node] for i in range(n_batch)]
    Masks = [1.0 * (x > 0.01) for x in Inpts]                                                             # :268
    if return_parts:
        return Inpts, Masks, dict(lp=lp, t_sel=t_sel, ph_sel=ph_sel, offset_per_batch=offset_per_batch,
                                  offset_per_station=offset_per_station)
    return Inpts, Masks


def legacy_pick_lists(arrivals, phase_labels, ind_use, time_samples, lp, n_sta):
    """process_utils.py:270-291: the per-sample pick lists (times relative to the sample, station slot, phase, meta),
    sorted by (station slot, time)."""
    sta_select = np.unique(ind_use)
    out = ([], [], [], [])
    for i in range(len(time_samples)):
        perm_vec = -1 * np.ones(n_sta)
        perm_vec[sta_select] = np.arange(len(sta_select))
        meta = arrivals[lp[i], :]
        phase_vals = phase_labels[lp[i]]
        times = meta[:, 0]
        indices = perm_vec[meta[:, 1].astype('int')]
        ineed = np.where(indices > -1)[0]
        times, indices, phase_vals, meta = times[ineed], indices[ineed], phase_vals[ineed], meta[ineed]
        lex_sort = np.lexsort((times, indices))
        out[0].append(times[lex_sort] - time_samples[i])
        out[1].append(indices[lex_sort])
        out[2].append(phase_vals[lex_sort])
        out[3].append(meta[lex_sort])
    return out


# --------------------------------------------------------------------------------------------------------------------
# a2 — DataAggregation (module.py:52-98)
# --------------------------------------------------------------------------------------------------------------------


def data_aggregation(sd, pre, Slice, Mask, A_in_sta, A_in_src, return_parts=False):
    """module.py:85-98.  `pre` is the state_dict prefix ('DataAggregation.')."""
    n = Slice.shape[0]

    def agg(edges, x):
        return propagate_mean(x.index_select(0, edges[0]), edges[1], n)

    tr0 = _prelu(sd, pre + 'activate', _lin(sd, pre + 'init_trns', torch.cat((Slice, Mask), dim=-1)))        # :87-88
    tr1 = _lin(sd, pre + 'l1_t1_2', torch.cat((tr0, agg(A_in_sta, _prelu(sd, pre + 'activate11', tr0)), Mask), dim=1))
    tr2 = _lin(sd, pre + 'l1_t2_2', torch.cat((tr0, agg(A_in_src, _prelu(sd, pre + 'activate12', tr0)), Mask), dim=1))
    tr = _prelu(sd, pre + 'activate1', torch.cat((tr1, tr2), dim=1))                                          # :90-92
    a = _prelu(sd, pre + 'activate21', _lin(sd, pre + 'l2_t1_1', tr))
    b = _prelu(sd, pre + 'activate22', _lin(sd, pre + 'l2_t2_1', tr))
    o1 = _lin(sd, pre + 'l2_t1_2', torch.cat((tr, agg(A_in_sta, a), Mask), dim=1))                            # :94
    o2 = _lin(sd, pre + 'l2_t2_2', torch.cat((tr, agg(A_in_src, b), Mask), dim=1))                            # :95
    out = _prelu(sd, pre + 'activate2', torch.cat((o1, o2), dim=1))                                           # :96
    if return_parts:
        return out, dict(tr0=tr0, tr=tr)
    return out


def edge_features(pos_loc, pos_src, A_src_in_sta, A_in_sta, A_in_src, scale_rel):
    """pos_rel_sta / pos_rel_src of the `use_updated_model_definition: True` model (module.py:1102-1111): per product edge
    [dx, dy, dz, |d|] of (source node - target node), embedded as sign(v) * exp(-0.5 v^2 / scale_rel^2)."""
    def emb(pos, idx, edges):
        d = pos[idx[edges[0]]] - pos[idx[edges[1]]]
        d = torch.cat((d, torch.norm(d, dim=1, keepdim=True)), dim=1)
        return torch.sign(d) * torch.exp(-0.5 * (d ** 2) / (scale_rel ** 2))
    return emb(pos_loc, A_src_in_sta[0], A_in_sta), emb(pos_src, A_src_in_sta[1], A_in_src)


def data_aggregation_edges(sd, pre, Slice, Mask, A_in_sta, A_in_src, pos_rel_sta, pos_rel_src, return_parts=False):
    """a2' — DataAggregationEdges.forward / message (module.py:143-174): as data_aggregation, but every message is
    [x_j | pos_rel(edge)] (4 more channels), so l*_t*_2 take 68 / 98 inputs ordered [tr | mean x_j | mean pos_rel | mask]."""
    n = Slice.shape[0]

    def agg(edges, x, pos_rel):
        return propagate_mean(torch.cat((x.index_select(0, edges[0]), pos_rel), dim=1), edges[1], n)        # :164-174

    tr0 = _prelu(sd, pre + 'activate', _lin(sd, pre + 'init_trns', torch.cat((Slice, Mask), dim=-1)))        # :145-146
    tr1 = _lin(sd, pre + 'l1_t1_2', torch.cat((tr0, agg(A_in_sta, _prelu(sd, pre + 'activate11', tr0), pos_rel_sta), Mask), dim=1))
    tr2 = _lin(sd, pre + 'l1_t2_2', torch.cat((tr0, agg(A_in_src, _prelu(sd, pre + 'activate12', tr0), pos_rel_src), Mask), dim=1))
    tr = _prelu(sd, pre + 'activate1', torch.cat((tr1, tr2), dim=1))                                          # :161-163
    a = _prelu(sd, pre + 'activate21', _lin(sd, pre + 'l2_t1_1', tr))
    b = _prelu(sd, pre + 'activate22', _lin(sd, pre + 'l2_t2_1', tr))
    o1 = _lin(sd, pre + 'l2_t1_2', torch.cat((tr, agg(A_in_sta, a, pos_rel_sta), Mask), dim=1))               # :165
    o2 = _lin(sd, pre + 'l2_t2_2', torch.cat((tr, agg(A_in_src, b, pos_rel_src), Mask), dim=1))               # :166
    out = _prelu(sd, pre + 'activate2', torch.cat((o1, o2), dim=1))                                           # :167
    if return_parts:
        return out, dict(tr0=tr0, tr=tr)
    return out


# --------------------------------------------------------------------------------------------------------------------
# a3 — BipartiteGraphOperator (module.py:214-229)
# --------------------------------------------------------------------------------------------------------------------


def bipartite_read_in(sd, pre, x_latent, edge_attr, edge_index, Mask):
    """module.py:224-229: per product node MLP, masked, summed onto its grid node, then a per-grid-node MLP."""
    M = int(edge_index[1].max()) + 1                                                                          # :227
    h = Mask.max(1, keepdim=True)[0] * _prelu(sd, pre + 'activate1',
                                              _lin(sd, pre + 'fc1', torch.cat((x_latent, edge_attr), dim=-1)))
    xg = propagate_sum(h.index_select(0, edge_index[0]), edge_index[1], M)
    return _prelu(sd, pre + 'activate2', _lin(sd, pre + 'fc2', xg))


# --------------------------------------------------------------------------------------------------------------------
# a4 — SpatialAggregation (module.py:231-249)
# --------------------------------------------------------------------------------------------------------------------


def spatial_aggregation(sd, pre, x, A_src, pos, scale_rel):
    """module.py:243-249.  The 'global' feature is a mean over EDGES of PReLU3(fglobal(x_j)) (:249)."""
    n = x.shape[0]
    p = pos / scale_rel
    xj = x.index_select(0, A_src[0])
    glob = _prelu(sd, pre + 'activate3', _lin(sd, pre + 'fglobal', xj)).mean(0, keepdim=True)
    msg = _prelu(sd, pre + 'activate1', _lin(sd, pre + 'fc1', torch.cat(
        (xj, p.index_select(0, A_src[1]) - p.index_select(0, A_src[0]), glob.repeat(xj.shape[0], 1)), dim=-1)))
    agg = propagate_mean(msg, A_src[1], n)
    return _prelu(sd, pre + 'activate2', _lin(sd, pre + 'fc2', torch.cat((x, agg), dim=-1)))


# --------------------------------------------------------------------------------------------------------------------
# a5 — read-out heads (module.py:251-331) and the whole forward_fixed_source (module.py:999-1020)
# --------------------------------------------------------------------------------------------------------------------


def spatial_direct(sd, pre, x):
    """module.py:251-260."""
    return _prelu(sd, pre + 'activate', _lin(sd, pre + 'f_direct', x))


def spatial_attention(sd, pre, x, x_query, x_context, scale_rel, k=10, n_heads=5, n_latent=15, edge_index=None):
    """module.py:280-297 (the f_queries variant of the current code)."""
    if edge_index is None:
        edge_index = knn(x_context / 1000.0, x_query / 1000.0, k).flip(0)                                     # :282
    ea = (x_query.index_select(0, edge_index[1]) - x_context.index_select(0, edge_index[0])) / scale_rel     # :283
    xj = x.index_select(0, edge_index[0])
    q = _lin(sd, pre + 'f_queries', ea).view(-1, n_heads, n_latent)
    c = _lin(sd, pre + 'f_context', torch.cat((xj, ea), dim=-1)).view(-1, n_heads, n_latent)
    v = _lin(sd, pre + 'f_values', torch.cat((xj, ea), dim=-1)).view(-1, n_heads, n_latent)
    alpha = _prelu(sd, pre + 'activate1', (q * c).sum(-1) / math.sqrt(n_latent))                              # :293
    alpha = segment_softmax(alpha, edge_index[1], x_query.shape[0])
    out = propagate_sum(alpha.unsqueeze(-1) * v, edge_index[1], x_query.shape[0])
    return _prelu(sd, pre + 'activate2', _lin(sd, pre + 'proj', out.mean(1)))                                 # :285


def temporal_attention(sd, pre, x, t_query, scale_t, n_heads=5, n_latent=15):
    """module.py:325-331."""
    c = _lin(sd, pre + 'f_context_2', _prelu(sd, pre + 'activate1', _lin(sd, pre + 'f_context_1', x)))
    v = _lin(sd, pre + 'f_values_2', _prelu(sd, pre + 'activate2', _lin(sd, pre + 'f_values_1', x)))
    q = _lin(sd, pre + 'temporal_query_2', _prelu(sd, pre + 'activate3',
                                                   _lin(sd, pre + 'temporal_query_1', t_query / scale_t)))
    c = c.view(-1, n_heads, n_latent)
    v = v.view(-1, n_heads, n_latent)
    q = q.view(-1, n_heads, n_latent)
    s = (c.unsqueeze(1) * q.unsqueeze(0)).sum(-1, keepdim=True) / math.sqrt(n_latent)      # [N, T, heads, 1]
    z = _prelu(sd, pre + 'activate4', (s * v.unsqueeze(1)).mean(2))                        # [N, T, latent]
    return _lin(sd, pre + 'proj_2', _prelu(sd, pre + 'activate5', _lin(sd, pre + 'proj_1', z)))


def absolute_pos_channels(Slice, locs_cart, grid_cart, A_src_in_sta, scale_rel):
    """`use_absolute_pos: True`, module.py:913-914: Slice gains [station position | source position] / (3 scale_rel)."""
    return torch.cat((Slice, locs_cart[A_src_in_sta[0]] / (3.0 * scale_rel), grid_cart[A_src_in_sta[1]] / (3.0 * scale_rel)),
                     dim=1)


def front_end(sd, Slice, Mask, A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart, scale_rel,
              return_parts=False, pos_rel=None, abs_pos=None):
    """a2 -> a3 -> a4 x3: the product-graph front end of module.py:1010-1014.  `pos_rel` = (pos_rel_sta, pos_rel_src) selects
    the `use_updated_model_definition: True` DataAggregationEdges (module.py:1176)."""
    if abs_pos is not None:              # (locs_cart, A_src_in_sta): the `use_absolute_pos: True` input channels
        Slice = absolute_pos_channels(Slice, abs_pos[0], grid_cart, abs_pos[1], scale_rel)
    if pos_rel is not None:
        x_latent = data_aggregation_edges(sd, 'DataAggregation.', Slice, Mask, A_in_sta, A_in_src, pos_rel[0], pos_rel[1])
    else:
        x_latent = data_aggregation(sd, 'DataAggregation.', Slice, Mask, A_in_sta, A_in_src)
    r = bipartite_read_in(sd, 'Bipartite_ReadIn.', x_latent, read_in_attr, read_in_index, Mask)
    x1 = spatial_aggregation(sd, 'SpatialAggregation1.', r, A_src, grid_cart, scale_rel)
    x2 = spatial_aggregation(sd, 'SpatialAggregation2.', x1, A_src, grid_cart, scale_rel)
    x3 = spatial_aggregation(sd, 'SpatialAggregation3.', x2, A_src, grid_cart, scale_rel)
    if return_parts:
        return x3, dict(x_latent=x_latent, read_in=r, sa1=x1, sa2=x2)
    return x3


def forward_fixed_source(sd, Slice, Mask, A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart,
                         x_query_cart, t_query, scale_rel, scale_t, return_parts=False, query_edges=None, pos_rel=None,
                         abs_pos=None):
    """module.py:999-1020 / 1165-1186 (use_absolute_pos False): returns y [G,T,1] and x [Q,T,1]."""
    x_spatial, parts = front_end(sd, Slice, Mask, A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart,
                                 scale_rel, return_parts=True, pos_rel=pos_rel, abs_pos=abs_pos)
    y_latent = spatial_direct(sd, 'SpatialDirect.', x_spatial)
    y = temporal_attention(sd, 'TemporalAttention.', y_latent, t_query, scale_t)
    xq = spatial_attention(sd, 'SpatialAttention.', x_spatial, x_query_cart, grid_cart, scale_rel,
                           edge_index=query_edges)
    x = temporal_attention(sd, 'TemporalAttention.', xq, t_query, scale_t)
    if return_parts:
        parts.update(x_spatial=x_spatial, y_latent=y_latent, x_query_embed=xq)
        return y, x, parts
    return y, x


def init_state(seed=2, scale=1.0, edges=False):
    """Random-init weights with the reference's key names / shapes (nn.Linear & nn.PReLU defaults are NOT reproduced —
    tests that need the reference's own init load a golden state_dict instead).  Used for seeded synthetic parity."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, n_in, n_out):
        bound = scale / math.sqrt(n_in)
        sd[name + '.weight'] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound
        sd[name + '.bias'] = (torch.rand(n_out, generator=g) * 2 - 1) * bound

    def prelu(name):
        sd[name + '.weight'] = torch.full((1,), 0.25) + 0.1 * (torch.rand(1, generator=g) - 0.5)

    p = 'DataAggregation.'
    e4 = 4 if edges else 0            # DataAggregationEdges: four edge-feature channels more (module.py:118-130)
    lin(p + 'init_trns', 8, 30); lin(p + 'l1_t1_1', 30, 30); lin(p + 'l1_t1_2', 64 + e4, 30)
    lin(p + 'l1_t2_1', 4, 30); lin(p + 'l1_t2_2', 64 + e4, 30); lin(p + 'l2_t1_1', 60, 30); lin(p + 'l2_t1_2', 94 + e4, 15)
    lin(p + 'l2_t2_1', 60, 30); lin(p + 'l2_t2_2', 94 + e4, 15)
    for a in ('activate', 'activate11', 'activate12', 'activate1', 'activate21', 'activate22', 'activate2'):
        prelu(p + a)
    p = 'Bipartite_ReadIn.'
    lin(p + 'fc1', 33, 30); lin(p + 'fc2', 30, 15); prelu(p + 'activate1'); prelu(p + 'activate2')
    for i, cin in ((1, 15), (2, 30), (3, 30)):
        p = 'SpatialAggregation%d.' % i
        lin(p + 'fc1', cin + 8, 30); lin(p + 'fc2', 30 + cin, 30); lin(p + 'fglobal', cin, 5)
        prelu(p + 'activate1'); prelu(p + 'activate2'); prelu(p + 'activate3')
    p = 'SpatialDirect.'
    lin(p + 'f_direct', 30, 30); prelu(p + 'activate')
    p = 'SpatialAttention.'
    sd[p + 'param_vector'] = torch.rand(1, 5, 15, generator=g) - 0.5
    lin(p + 'f_queries', 3, 75); lin(p + 'f_context', 33, 75); lin(p + 'f_values', 33, 75)
    lin(p + 'f_direct', 30, 30); lin(p + 'proj', 15, 30); prelu(p + 'activate1'); prelu(p + 'activate2')
    p = 'TemporalAttention.'
    lin(p + 'temporal_query_1', 1, 30); lin(p + 'temporal_query_2', 30, 75)
    lin(p + 'f_context_1', 30, 30); lin(p + 'f_context_2', 30, 75)
    lin(p + 'f_values_1', 30, 30); lin(p + 'f_values_2', 30, 75)
    lin(p + 'proj_1', 15, 30); lin(p + 'proj_2', 30, 1)
    for a in ('activate1', 'activate2', 'activate3', 'activate4', 'activate5'):
        prelu(p + a)
    return sd


# --------------------------------------------------------------------------------------------------------------------
# association branch: module.py:333-352, 356-403, 604-653, 656-781, 963-997
# --------------------------------------------------------------------------------------------------------------------


def bipartite_read_out(sd, pre, y_latent, edge_attr, edge_index, mask_out):
    """BipartiteGraphReadOutOperator, module.py:333-352: messages grid node -> product node (`A_Lg_in_src`, edge_index =
    [g(i); i], process_continuous_days.py:632), aggr='add', mask_j = source-prediction mask of the grid node.  Returns
    (s [P,15], mask_out_1 [E,1] = mask_out[edge_index[0]])."""
    M = int(edge_index[1].max()) + 1
    xj = y_latent.index_select(0, edge_index[0])
    mj = mask_out.index_select(0, edge_index[0])
    msg = mj * _prelu(sd, pre + 'activate1', _lin(sd, pre + 'fc1', torch.cat((xj, edge_attr), dim=-1)))
    out = _prelu(sd, pre + 'activate2', _lin(sd, pre + 'fc2', propagate_sum(msg, edge_index[1], M)))
    return out, mj


def data_aggregation_association(sd, pre, s, x_latent, mask1, mask2, A_in_sta, A_in_src, return_parts=False, pos_rel=None):
    """DataAggregationAssociationPhase.forward, module.py:387-403 (`use_updated_model_definition: False`).  Unlike
    DataAggregation, the layer-1 messages pass through l1_t*_1 first and the mask has five channels.
    `pos_rel` = (pos_rel_sta, pos_rel_src): DataAggregationAssociationPhaseEdges, module.py:442-481 — every message is
    [x_j | pos_rel(edge)] (message_type 1 / 2)."""
    n = s.shape[0]
    mask = torch.cat((mask1, mask2), dim=-1)
    tr = _prelu(sd, pre + 'activate', _lin(sd, pre + 'init_trns', torch.cat((s, x_latent, mask), dim=-1)))

    def agg(x, A):
        msg = x.index_select(0, A[0])
        if pos_rel is not None:
            msg = torch.cat((msg, pos_rel[0] if A is A_in_sta else pos_rel[1]), dim=1)
        return propagate_mean(msg, A[1], n)

    a1 = _prelu(sd, pre + 'activate11', _lin(sd, pre + 'l1_t1_1', tr))
    a2 = _prelu(sd, pre + 'activate12', _lin(sd, pre + 'l1_t2_1', tr))
    tr1 = _lin(sd, pre + 'l1_t1_2', torch.cat((tr, agg(a1, A_in_sta), mask), dim=1))
    tr2 = _lin(sd, pre + 'l1_t2_2', torch.cat((tr, agg(a2, A_in_src), mask), dim=1))
    trb = _prelu(sd, pre + 'activate1', torch.cat((tr1, tr2), dim=1))
    b1 = _prelu(sd, pre + 'activate21', _lin(sd, pre + 'l2_t1_1', trb))
    b2 = _prelu(sd, pre + 'activate22', _lin(sd, pre + 'l2_t2_1', trb))
    o1 = _lin(sd, pre + 'l2_t1_2', torch.cat((trb, agg(b1, A_in_sta), mask), dim=1))
    o2 = _lin(sd, pre + 'l2_t2_2', torch.cat((trb, agg(b2, A_in_src), mask), dim=1))
    out = _prelu(sd, pre + 'activate2', torch.cat((o1, o2), dim=1))
    if return_parts:
        return out, dict(assoc_tr0=tr, assoc_tr1=trb)
    return out


def local_slice_collapse(sd, pre, A_edges, dt_partition, tpick, ipick, phase_label, inpt, tlatent, eps,
                         use_phase_types=True, k_infer=10):
    """LocalSliceLgCollapse.forward/message, module.py:617-653: every pick selects the k_infer product nodes its
    (station, time bin) row of the pointer table names, keeps those whose theoretical time is within 2*eps, and averages
    PReLU(fc1([s_j ‖ (t_pick - t_j)/eps ‖ phase])) over them."""
    n_arv, l_dt = tpick.shape[0], dt_partition.shape[0]
    dt = dt_partition[1] - dt_partition[0]
    ph = phase_label.float() if use_phase_types else phase_label.float() * 0.0
    t_index = torch.floor((tpick - dt_partition[0]) / dt).long()                                              # :630
    t_index = ((ipick * l_dt * k_infer + t_index * k_infer).view(-1, 1) + torch.arange(k_infer).view(1, -1)).reshape(-1)
    src_index = torch.arange(n_arv).view(-1, 1).repeat(1, k_infer).view(-1)
    e0 = A_edges[t_index].long()
    t_rel = tpick[src_index] - tlatent[e0, 0]                                                                 # :637
    keep = torch.where(t_rel.abs() < 2.0 * eps)[0]
    e0, e1 = e0[keep], src_index[keep]
    msg = torch.cat((inpt.index_select(0, e0), (tpick.view(-1, 1)[e1] - tlatent[e0]) / eps, ph[e1]), dim=-1)
    msg = _prelu(sd, pre + 'activate1', _lin(sd, pre + 'fc1', msg))
    return _prelu(sd, pre + 'activate2', _lin(sd, pre + 'fc2', propagate_mean(msg, e1, n_arv)))


def arrival_edges(ipick, n_arv):
    """module.py:714: for every pick a (target) all picks b on the same station plus the null arrival n_arv (sources),
    in the reference's order (stations ascending; itertools.product(list, list + [null])).  Returns [2,E]: row 0 = b, row 1 = a.
    The per-station lists come from cKDTree.query_ball_point(r=0) in the reference; SORTED here — the edge order only
    permutes the terms of the per-target sums."""
    ip = ipick.numpy()
    src, dst = [], []
    for u in np.unique(ip):
        lst = np.where(ip == u)[0]
        ext = np.concatenate((lst, np.array([n_arv])))
        dst.append(np.repeat(lst, len(ext)))
        src.append(np.tile(ext, len(lst)))
    if not src:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.from_numpy(np.stack((np.concatenate(src), np.concatenate(dst)))).long()


def station_source_attention(sd, pre, stime, src_embed, trv_src, arrival_p, arrival_s, tpick, ipick, phase_label, eps,
                             use_phase_types=True, n_heads=3, n_latent=15):
    """StationSourceAttentionMergedPhases.forward/message, module.py:698-781 (use_neighbor_assoc_edges False)."""
    n_src, n_sta, n_arv = trv_src.shape[0], trv_src.shape[1], tpick.shape[0]
    ph = phase_label.float() if use_phase_types else phase_label.float() * 0.0
    arrival = torch.cat((torch.cat((arrival_p, torch.zeros(1, arrival_p.shape[1])), dim=0),
                         torch.cat((arrival_s, torch.zeros(1, arrival_s.shape[1])), dim=0)), dim=1)           # :709-711
    edges = arrival_edges(ipick, n_arv)
    n_edge = edges.shape[1]
    e0 = edges[0].repeat(n_src)                                                                               # :718
    e1 = edges[1].repeat(n_src) + (torch.arange(n_src) * n_arv).repeat_interleave(n_edge)
    sindex = torch.arange(n_src).repeat_interleave(n_edge)
    atime = torch.cat((tpick, torch.tensor([-eps], dtype=tpick.dtype)))
    stindex = torch.cat((ipick, torch.tensor([n_sta], dtype=torch.long)))
    tsrc_p = torch.cat((trv_src[:, :, 0], -eps * torch.ones(n_src, 1)), dim=1)
    tsrc_s = torch.cat((trv_src[:, :, 1], -eps * torch.ones(n_src, 1)), dim=1)
    phase = torch.cat((ph, torch.tensor([[-1.0]])), dim=0)
    t_kernel_sq = torch.tensor([eps], dtype=torch.float32) ** 2
    rel_p = (atime[e0] - (tsrc_p[sindex, stindex[e0]] + stime[sindex])).reshape(-1, 1)
    rel_s = (atime[e0] - (tsrc_s[sindex, stindex[e0]] + stime[sindex])).reshape(-1, 1)
    thr = 2.0 * torch.sqrt(t_kernel_sq)
    keep = torch.where((((rel_p.abs() < thr) + (rel_s.abs() < thr)).reshape(-1)) > 0)[0]                      # :727
    e0, e1, sindex, rel_p, rel_s = e0[keep], e1[keep], sindex[keep], rel_p[keep], rel_s[keep]
    M = n_arv * n_src
    if e0.numel() > 0:
        # message (:746-781); `edge_index[0].max()` is taken over the KEPT edges, as in the reference
        e0max = int(e0.max())
        f_p = torch.cat((torch.exp(-0.5 * (rel_p ** 2) / t_kernel_sq), torch.sign(rel_p), phase[e0]), dim=1)
        f_s = torch.cat((torch.exp(-0.5 * (rel_s ** 2) / t_kernel_sq), torch.sign(rel_s), phase[e0]), dim=1)
        self_link = (e0 == torch.remainder(e1, e0max)).reshape(-1, 1).float()                                # :762
        null_link = (e0 == e0max).reshape(-1, 1).float()
        xj = arrival.index_select(0, e0)
        ctx = _lin(sd, pre + 'f_src_context_2', _prelu(sd, pre + 'activate1', _lin(sd, pre + 'f_src_context_1', torch.cat(
            (src_embed[sindex], stime[sindex].reshape(-1, 1), self_link, null_link), dim=1)))).view(-1, n_heads, n_latent)
        qry = _lin(sd, pre + 'f_arrival_query_2', _prelu(sd, pre + 'activate2', _lin(sd, pre + 'f_arrival_query_1', torch.cat(
            (xj, f_p, f_s), dim=1)))).view(-1, n_heads, n_latent)
        val = _lin(sd, pre + 'f_values_2', _prelu(sd, pre + 'activate3', _lin(sd, pre + 'f_values_1', torch.cat(
            (xj, f_p, f_s, self_link, null_link), dim=1)))).view(-1, n_heads, n_latent)
        scores = (qry * ctx).sum(-1) / math.sqrt(n_latent)
        alpha = segment_softmax(scores, e1, M)
        agg = propagate_sum(alpha.unsqueeze(-1) * val, e1, M)
    else:
        agg = torch.zeros(M, n_heads, n_latent)
    out = _lin(sd, pre + 'proj_2', _prelu(sd, pre + 'activate4', _lin(sd, pre + 'proj_1', agg.mean(1))))     # :742
    return out.view(n_src, n_arv, -1)


def forward_fixed(sd, Slice, Mask, A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart, A_edges_p,
                  A_edges_s, dt_partition, tlatent, tpick, ipick, phase_label, x_query_cart, x_query_src_cart, t_query,
                  tq_sample, trv_out_q, scale_rel, scale_t, eps, return_parts=False, query_edges=None,
                  query_src_edges=None, pos_rel=None, abs_pos=None):
    """module.py:963-997 (= forward, :908-939, with the adjacencies passed per call): returns y, x, arv_p, arv_s."""
    x_spatial, parts = front_end(sd, Slice, Mask, A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart,
                                 scale_rel, return_parts=True, pos_rel=pos_rel, abs_pos=abs_pos)
    x_latent = parts['x_latent']
    y_latent = spatial_direct(sd, 'SpatialDirect.', x_spatial)
    y = temporal_attention(sd, 'TemporalAttention.', y_latent, t_query, scale_t)
    xq = spatial_attention(sd, 'SpatialAttention.', x_spatial, x_query_cart, grid_cart, scale_rel, edge_index=query_edges)
    x_src = spatial_attention(sd, 'SpatialAttention.', x_spatial, x_query_src_cart, grid_cart, scale_rel,
                              edge_index=query_src_edges)
    x = temporal_attention(sd, 'TemporalAttention.', xq, t_query, scale_t)
    mask_out = 1.0 * (y[:, :, 0].max(1, keepdim=True)[0] > 0.01)                                              # :983
    A_Lg = torch.stack((read_in_index[1], read_in_index[0]))                     # process_continuous_days.py:632
    s0, mask_out_1 = bipartite_read_out(sd, 'BipartiteGraphReadOutOperator.', y_latent, read_in_attr, A_Lg, mask_out)
    s_in = s0
    if abs_pos is not None:                                                                                   # :987-988
        s_in = absolute_pos_channels(s0, abs_pos[0], grid_cart, abs_pos[1], scale_rel)
    # x_latent enters detached, as in the reference (`x_latent.detach()`, :932): no gradient reaches DataAggregation this way
    s = data_aggregation_association(sd, 'DataAggregationAssociationPhase.', s_in, x_latent.detach(), mask_out_1, Mask, A_in_sta,
                                     A_in_src, pos_rel=pos_rel)
    arv_p = local_slice_collapse(sd, 'LocalSliceLgCollapseP.', A_edges_p, dt_partition, tpick, ipick, phase_label, s,
                                 tlatent[:, 0].reshape(-1, 1), eps)
    arv_s = local_slice_collapse(sd, 'LocalSliceLgCollapseS.', A_edges_s, dt_partition, tpick, ipick, phase_label, s,
                                 tlatent[:, 1].reshape(-1, 1), eps)
    arv = station_source_attention(sd, 'Arrivals.', tq_sample, x_src, trv_out_q, arv_p, arv_s, tpick, ipick,
                                   phase_label, eps)
    out = (y, x, arv[:, :, 0].unsqueeze(-1), arv[:, :, 1].unsqueeze(-1))
    if return_parts:
        parts.update(x_spatial=x_spatial, y_latent=y_latent, x_src=x_src, mask_out=mask_out, assoc_s0=s0, assoc_s=s,
                     arv_p_embed=arv_p, arv_s_embed=arv_s)
        return out + (parts,)
    return out


def init_state_association(sd, seed=7, scale=1.0):
    """Adds seeded weights for the association modules (module.py:900-904) to a state made by init_state."""
    g = torch.Generator().manual_seed(seed)

    def lin(name, n_in, n_out):
        bound = scale / math.sqrt(n_in)
        sd[name + '.weight'] = (torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound
        sd[name + '.bias'] = (torch.rand(n_out, generator=g) * 2 - 1) * bound

    def prelu(name):
        sd[name + '.weight'] = torch.full((1,), 0.25) + 0.1 * (torch.rand(1, generator=g) - 0.5)

    p = 'BipartiteGraphReadOutOperator.'
    lin(p + 'fc1', 33, 30); lin(p + 'fc2', 30, 15); prelu(p + 'activate1'); prelu(p + 'activate2')
    p = 'DataAggregationAssociationPhase.'
    lin(p + 'init_trns', 50, 30); lin(p + 'l1_t1_1', 30, 30); lin(p + 'l1_t1_2', 65, 30)
    lin(p + 'l1_t2_1', 30, 30); lin(p + 'l1_t2_2', 65, 30); lin(p + 'l2_t1_1', 60, 30); lin(p + 'l2_t1_2', 95, 15)
    lin(p + 'l2_t2_1', 60, 30); lin(p + 'l2_t2_2', 95, 15)
    for a in ('activate', 'activate11', 'activate12', 'activate1', 'activate21', 'activate22', 'activate2'):
        prelu(p + a)
    for p in ('LocalSliceLgCollapseP.', 'LocalSliceLgCollapseS.'):
        lin(p + 'fc1', 32, 30); lin(p + 'fc2', 30, 15); prelu(p + 'activate1'); prelu(p + 'activate2')
    p = 'Arrivals.'
    lin(p + 'f_arrival_query_1', 36, 30); lin(p + 'f_arrival_query_2', 30, 45)
    lin(p + 'f_src_context_1', 33, 30); lin(p + 'f_src_context_2', 30, 45)
    lin(p + 'f_values_1', 38, 30); lin(p + 'f_values_2', 30, 45)
    lin(p + 'proj_1', 15, 30); lin(p + 'proj_2', 30, 2)
    for a in ('activate1', 'activate2', 'activate3', 'activate4'):
        prelu(p + a)
    return sd


# --------------------------------------------------------------------------------------------------------------------
# caller-side streaming loop: process_continuous_days.py:757-813, one source grid,
# `use_updated_input: True`.  The loop lives in the reference's script body (not importable); it is PINNED by
# tests/golden/streaming_10x100.npz, which oracle/gen_golden.py `streaming` produced by executing those source lines verbatim
# (read from the reference file at generation time) over the unmodified reference module / process_utils
# (tests/test_oracle_golden.py::test_streaming_loop_matches_reference).
# --------------------------------------------------------------------------------------------------------------------


def continuous_day_stack(sd, P, tsteps, tsteps_abs, ind_use, n_locs, A_src_in_sta, trv_times, max_t, kernel_sig_t, dt,
                         A_in_sta, A_in_src, read_in_attr, read_in_index, A_src, grid_cart, x_query_cart, scale_rel, scale_t,
                         t_win=6.0, dt_win=0.75, step_size='half', n_scale_x_grid=1, pick_t_win=10.0):
    from scipy.spatial import cKDTree
    n_overlap = {'full': 1.0, 'partial': 3.0, 'half': 2.0}[step_size]                        # :369-379
    tsteps_abs = np.asarray(tsteps_abs, dtype=np.float64)
    tree_tsteps = cKDTree(tsteps_abs.reshape(-1, 1))                                          # :412
    Out_2 = np.zeros((x_query_cart.shape[0], len(tsteps_abs)))                                # :758
    t_rel = np.arange(-t_win / 2.0, t_win / 2.0 + dt_win, dt_win)
    tq = torch.from_numpy(t_rel).reshape(-1, 1).float()                                       # :534
    idx = tree_tsteps.query(np.asarray(tsteps, dtype=np.float64).reshape(-1, 1))[1]           # :765
    n_done = 0
    for i0, t0 in enumerate(np.asarray(tsteps, dtype=np.float64)):
        # pick list of the window (process_utils.py:476-481, 665): only its length matters here (:787)
        sel = (P[:, 0] > (t0 - 2.0 * kernel_sig_t)) * (P[:, 0] < (t0 + max_t + 2.0 * kernel_sig_t))
        sel = sel * np.isin(P[:, 1].astype('int'), np.asarray(ind_use))
        sel = sel * (np.abs(P[:, 0] - (t0 + max_t / 2.0)) <= (pick_t_win + max_t / 2.0))
        if sel.sum() == 0:
            continue
        Slice, Mask = input_scatter(P, t0, ind_use, n_locs, A_src_in_sta, trv_times, max_t, kernel_sig_t, dt)
        ip_need = tree_tsteps.query(tsteps_abs[idx[i0]] + t_rel.reshape(-1, 1))               # :793
        _, x = forward_fixed_source(sd, torch.from_numpy(Slice), torch.from_numpy(Mask), A_in_sta, A_in_src, read_in_attr,
                                    read_in_index, A_src, grid_cart, x_query_cart, tq, scale_rel, scale_t)
        if step_size == 'half':                                                                # :802-805
            Out_2[:, ip_need[1][0:-1]] += x[:, 0:-1, 0].numpy() / n_overlap / n_scale_x_grid
        else:
            Out_2[:, ip_need[1]] += x[:, :, 0].numpy() / n_overlap / n_scale_x_grid
        n_done += 1
    return Out_2, n_done
