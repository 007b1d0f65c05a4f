"""Host-side mirrors of the reference's `Code/process_utils.py` entry points that feed the hot path.

* `extract_input_from_data`  — same signature and return structure as process_utils.py:460, but the per-station series,
  the travel-time-shifted gather and Slice/Mask live on the GPU (libgenie_b200 `genie_input_scatter_fwd`).
* `extract_inputs_adjacencies` — the dense-mode graph builder of process_utils.py:701-742.  kNN graphs are static per
  station set; `device=` runs the searches in libgenie_b200's kNN kernel (genie_knn_fwd), otherwise a host k-d tree is used.
* `InputExtractor` — the streaming form used by bench.py: travel times, station tables and (optionally) a whole day of
  picks stay resident in HBM; one call per window.
"""
import numpy as np
import torch
from scipy.spatial import cKDTree

from . import capi, ops
from .plan import GraphPlan


# `from genie_b200.process_utils import *` after the reference's `from process_utils import *` overrides exactly these names
__all__ = ['extract_input_from_data', 'extract_inputs_from_data_fixed_grids_with_phase_type', 'extract_pick_inputs_from_data',
           'extract_inputs_adjacencies', 'extract_inputs_adjacencies_subgraph', 'compute_time_embedding_vectors', 'extract_inputs_adjacencies_cartesian',
           'product_edge_lists', 'knn_graph', 'knn_graph_device', 'InputExtractor']

# ---- graphs -------------------------------------------------------------------------------------------------------------

def knn_graph(pos_km, k):
    """`remove_self_loops(knn(x, x, k+1).flip(0))` (process_utils.py:718-719): int64 [2,E], row 0 source, row 1 target."""
    pos_km = np.asarray(pos_km, dtype=np.float32).astype(np.float64)
    n = pos_km.shape[0]
    kk = int(min(k + 1, n))
    ind = cKDTree(pos_km).query(pos_km, k=kk)[1].reshape(n, kk)
    tgt = np.repeat(np.arange(n), kk)
    src = ind.reshape(-1)
    keep = src != tgt
    return torch.from_numpy(np.stack((src[keep], tgt[keep]), axis=0)).long()


def knn_graph_device(pos_km, k):
    """knn_graph on the device (genie_knn_fwd): pos_km fp32 CUDA tensor [n,3] -> int64 [2,E] on the same device, edges
    grouped by target, nearest source first — the order `knn(x, x, k+1).flip(0)` + remove_self_loops produces."""
    n = pos_km.shape[0]
    idx = ops.knn(pos_km, pos_km, min(k + 1, n))                       # [n, k+1]
    tgt = torch.arange(n, device=pos_km.device).view(-1, 1).expand_as(idx)
    keep = idx != tgt
    return torch.stack((idx[keep], tgt[keep]), dim=0)


def extract_inputs_adjacencies_cartesian(locs_cart, grid_cart, k_sta_edges, k_spc_edges, device=None):
    """The two small graphs of process_utils.py:712-719 for Cartesian coordinates in metres.  With `device` (a CUDA device)
    the searches run in libgenie_b200's kNN kernel and the edge lists stay on the device; otherwise a host k-d tree is used
    (set-up code, once per station set)."""
    k_sta = int(min(k_sta_edges, locs_cart.shape[0] - 2))
    if device is not None:
        km = lambda a: torch.from_numpy((np.asarray(a, dtype=np.float64) / 1000.0).astype(np.float32)).to(device)
        return knn_graph_device(km(locs_cart), k_sta), knn_graph_device(km(grid_cart), k_spc_edges)
    A_sta_sta = knn_graph((np.asarray(locs_cart, dtype=np.float64) / 1000.0).astype(np.float32), k_sta)
    A_src_src = knn_graph((np.asarray(grid_cart, dtype=np.float64) / 1000.0).astype(np.float32), k_spc_edges)
    return A_sta_sta, A_src_src


def extract_inputs_adjacencies(trv, locs, ind_use, x_grid, x_grid_trv, x_grid_trv_ref, x_grid_trv_pointers_p,
                               x_grid_trv_pointers_s, ftrns1, graph_params, device=None, verbose=False):
    """Dense-mode graph builder with the reference's signature and return list (process_utils.py:701-742):
    [A_sta_sta, A_src_src, A_prod_sta_sta, A_prod_src_src, A_src_in_prod, A_edges_time_p, A_edges_time_s, A_edges_ref].
    kNN through genie_knn_fwd when `device` is a CUDA device.  The explicit product lists are returned because the reference's
    callers expect them (`set_adjacencies` recognises their pattern and keeps only the two small graphs); beyond a few 10^7
    product nodes use extract_inputs_adjacencies_cartesian + set_adjacencies_cartesian instead."""
    k_sta_edges, k_spc_edges, k_time_edges = graph_params
    ind_use = np.asarray(ind_use).astype('int')
    n_sta, n_spc, n_sta_slice = locs.shape[0], x_grid.shape[0], len(ind_use)
    A_sta_sta, A_src_src = extract_inputs_adjacencies_cartesian(ftrns1(locs[ind_use]), ftrns1(x_grid), k_sta_edges,
                                                                k_spc_edges, device=device)
    dev = A_sta_sta.device
    A_prod_sta, A_prod_src, A_src_in_prod, _ = (a.to(dev) for a in product_edge_lists(A_sta_sta.cpu(), A_src_src.cpu(),
                                                                                       n_sta_slice, n_spc))
    # Time-pointer tables of the used stations (:723-734).  A pointer is a product-node id over ALL stations,
    # g * n_sta + station; the block of station u (k_time_edges * len_dt entries) moves to slot u of the subset and its
    # entries become g * n_sta_slice + u.  The division is carried out in floating point as the reference's is.
    blk = int(k_time_edges) * len(x_grid_trv_ref)
    slot = np.arange(n_sta_slice, dtype=np.float64).reshape(-1, 1)

    def subset_pointers(table):
        rows = np.asarray(table).reshape(n_sta, blk)[ind_use]                           # [n_sta_slice, blk]
        return ((n_sta_slice * (rows - ind_use.reshape(-1, 1))) / n_sta + slot).reshape(-1)
    A_edges_time_p, A_edges_time_s = subset_pointers(x_grid_trv_pointers_p), subset_pointers(x_grid_trv_pointers_s)
    A_edges_ref = np.array(x_grid_trv_ref, copy=True)
    if max(A_edges_time_p.max(), A_edges_time_s.max()) >= n_spc * n_sta_slice:
        raise ValueError('time pointers reference product nodes outside the station subset')
    return [A_sta_sta, A_src_src, A_prod_sta, A_prod_src, A_src_in_prod, A_edges_time_p, A_edges_time_s, A_edges_ref]


def extract_inputs_adjacencies_subgraph(locs, x_grid, ftrns1, ftrns2=None, max_deg_offset=5.0, k_nearest_pairs=30,
                                        k_sta_edges=10, k_spc_edges=15, verbose=False,
                                        scale_pairwise_sta_in_src_distances=100e3, scale_deg=110e3, device=None):
    """Sub-graph mode of the reference (process_utils.py:744-849), same signature and return list
    [A_sta_sta, A_src_src, A_prod_sta_sta, A_prod_src_src, A_src_in_prod, A_src_in_sta]: the product graph only over the
    (station, source) pairs that are within `scale_deg * max_deg_offset` metres or among the source's `k_nearest_pairs`
    nearest stations.  The reference loops over all grid nodes and all stations in Python (:798-841); here the pair set is
    one sorted key list, membership a dense [G*S] index table, and the two product edge lists are expansions of the small
    graphs' CSR rows — torch ops on `device` (CUDA: kNN through genie_knn_fwd), no Python loop.  Edge order as the reference's."""
    dev = torch.device('cpu') if device is None else torch.device(device)
    on_gpu = dev.type == 'cuda'
    sta_m = np.asarray(ftrns1(locs), dtype=np.float64)
    grid_m = np.asarray(ftrns1(x_grid), dtype=np.float64)
    S, G = sta_m.shape[0], grid_m.shape[0]
    k_sta = int(min(k_sta_edges, S - 2))                                                              # :764
    km = lambda a: torch.from_numpy((a / 1000.0).astype(np.float32))
    if on_gpu:
        A_sta_sta, A_src_src = knn_graph_device(km(sta_m).to(dev), k_sta), knn_graph_device(km(grid_m).to(dev), k_spc_edges)
        nn_idx = ops.knn(km(sta_m).to(dev), km(grid_m).to(dev), int(k_nearest_pairs))             # [G, k]  (:784)
    else:
        A_sta_sta, A_src_src = knn_graph(km(sta_m).numpy(), k_sta), knn_graph(km(grid_m).numpy(), k_spc_edges)
        kk = int(min(k_nearest_pairs, S))
        nn_idx = torch.from_numpy(cKDTree(km(sta_m).numpy().astype(np.float64)).query(
            km(grid_m).numpy().astype(np.float64), k=kk)[1].reshape(G, kk)).long()
    # pair set as sorted keys g*S + s  (:776-794: epsilon pairs + kNN pairs, unique, ordered by source then station)
    g_t, s_t = torch.from_numpy(grid_m).to(dev), torch.from_numpy(sta_m).to(dev)
    keys = [(torch.arange(G, device=dev).view(-1, 1) * S + nn_idx.long()).reshape(-1)]
    thr = float(scale_deg) * float(max_deg_offset)
    for g0 in range(0, G, 4096):                                    # fp64 distances as numpy's in the reference (:777-779)
        d = torch.linalg.norm(g_t[g0:g0 + 4096, None, :] - s_t[None, :, :], dim=2)
        gi, si = torch.nonzero(d < thr, as_tuple=True)
        keys.append((gi + g0) * S + si)
    keys = torch.unique(torch.cat(keys))                             # sorted
    n_prod = keys.numel()
    node_grid, node_sta = keys // S, keys % S
    A_src_in_sta = torch.stack((node_sta, node_grid), dim=0)
    A_src_in_prod = torch.stack((torch.arange(n_prod, device=dev), node_grid), dim=0)
    table = torch.full((G * S,), -1, dtype=torch.long, device=dev)
    table[keys] = torch.arange(n_prod, device=dev)

    def expand(A_small, n_small, node_row):
        """For every product node, the in-edges of its row in the small graph (CSR by target, the small graph's own order)."""
        rowptr, col = csr_by_destination_stable(A_small, n_small)
        deg = (rowptr[1:] - rowptr[:-1])[node_row]
        tgt = torch.repeat_interleave(torch.arange(n_prod, device=dev), deg)
        start = torch.repeat_interleave(rowptr[:-1][node_row], deg)
        off = torch.arange(tgt.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg)
        return col[start + off], tgt

    # station edges (:823-826): per source, the station graph restricted to the source's station list, in A_sta_sta's order
    sj, tgt = expand(A_sta_sta.to(dev), S, node_sta)
    src = table[node_grid[tgt] * S + sj]
    keep = src >= 0
    A_prod_sta_sta = torch.stack((src[keep], tgt[keep]), dim=0)
    # source edges (:828-846): per station, the source graph restricted to the sources linked to it; sorted by (target, source)
    gj, tgt = expand(A_src_src.to(dev), G, node_grid)
    src = table[gj * S + node_sta[tgt]]
    keep = src >= 0
    src, tgt = src[keep], tgt[keep]
    order = torch.sort(tgt * n_prod + src)[1]
    A_prod_src_src = torch.stack((src[order], tgt[order]), dim=0)
    return [A_sta_sta.to(dev), A_src_src.to(dev), A_prod_sta_sta, A_prod_src_src, A_src_in_prod, A_src_in_sta]


def compute_time_embedding_vectors(trv_pairwise, locs, x_grid, A_src_in_sta, max_t, dt_res=1.0, k_times=10, t_win=10,
                                   device=None, trv_out=None):
    """Time-pointer tables of the association branch with the reference's signature (process_utils.py:851-877): for every
    station and every step of `dt_partition = arange(-t_win, max_t + t_win + dt_res, dt_res)`, the `k_times` product nodes of
    that station whose P (resp. S) travel time is nearest to the step, nearest first.  The reference queries two 2-D k-d
    trees whose first coordinate (station index x 2 (max_t + t_win + dt_res)) only separates the stations; here every
    station's travel times are sorted once and the 2 k_times candidates around each step's insertion point are ranked —
    torch ops on `device`, no tree.  `trv_out` [P,2] may be given instead of calling `trv_pairwise`.
    Returns (edges_time_p, edges_time_s, dt_partition) as the reference does (int64 arrays, float64 array)."""
    dev = torch.device('cpu') if device is None else torch.device(device)
    A = torch.as_tensor(A_src_in_sta).to(dev).long()
    if trv_out is None:
        trv_out = trv_pairwise(torch.Tensor(locs).to(dev)[A[0]], torch.Tensor(x_grid).to(dev)[A[1]])
    trv_out = torch.as_tensor(trv_out).to(dev).double()
    n_sta, n_prod, k = int(len(locs)), int(A.shape[1]), int(k_times)
    dt_partition = np.arange(-t_win, max_t + t_win + dt_res, dt_res)
    q = torch.from_numpy(dt_partition).to(dev)                                        # [L]
    L = q.numel()
    sta = A[0]
    cnt = torch.bincount(sta, minlength=n_sta)
    if int(cnt.min()) < k:
        raise capi.GenieError('compute_time_embedding_vectors: every station needs at least k_times product nodes')
    start = torch.cumsum(cnt, 0) - cnt
    big = 4.0 * float(max_t + t_win + dt_res) + 1.0        # separates the stations inside one global sort key
    out = []
    for ph in (0, 1):
        t = trv_out[:, ph]
        order = torch.sort(sta.double() * big + t)[1]                                 # nodes grouped by station, by time
        ts = t[order]
        # insertion point of every (station, step) in that station's sorted times
        keys = (torch.arange(n_sta, device=dev).double() * big).view(-1, 1) + q.view(1, -1)      # [S, L]
        pos = torch.searchsorted((sta[order].double() * big + ts).contiguous(), keys.reshape(-1).contiguous()).view(n_sta, L)
        cand = pos.unsqueeze(-1) + torch.arange(-k, k, device=dev).view(1, 1, -1)                # [S, L, 2k] global positions
        lo, hi = start.view(-1, 1, 1), (start + cnt).view(-1, 1, 1)
        valid = (cand >= lo) & (cand < hi)
        candc = cand.clamp(0, n_prod - 1)
        dist = (ts[candc] - q.view(1, -1, 1)).abs()
        dist = torch.where(valid, dist, torch.full_like(dist, float('inf')))
        sel = torch.topk(dist, k, dim=-1, largest=False, sorted=True)[1]
        out.append(order[torch.gather(candc, -1, sel)].reshape(-1).cpu().numpy())
    return out[0], out[1], dt_partition


def csr_by_destination_stable(A, n):
    """CSR by target node of an int64 [2,E] edge list, keeping the list's order inside every row: (rowptr [n+1], source [E])."""
    order = torch.sort(A[1], stable=True)[1]
    rowptr = torch.zeros(n + 1, dtype=torch.long, device=A.device)
    rowptr[1:] = torch.cumsum(torch.bincount(A[1], minlength=n), 0)
    return rowptr, A[0][order]


def product_edge_lists(A_sta_sta, A_src_src, n_sta, n_grid):
    """A_prod_sta_sta, A_prod_src_src, A_src_in_prod, A_src_in_sta of process_utils.py:720-722 /
    process_continuous_days.py:629 (small networks only: these lists are what the CARTESIAN plan avoids)."""
    S, G = n_sta, n_grid
    A_prod_sta = (A_sta_sta.repeat(1, G) + S * torch.arange(G).repeat_interleave(A_sta_sta.shape[1]).view(1, -1))
    A_prod_src = (S * A_src_src.repeat(1, S) + torch.arange(S).repeat_interleave(A_src_src.shape[1]).view(1, -1))
    A_src_in_prod = torch.stack((torch.arange(S * G), torch.arange(G).repeat_interleave(S)), dim=0)
    A_src_in_sta = torch.stack((torch.arange(S).repeat(G), torch.arange(G).repeat_interleave(S)), dim=0)
    return A_prod_sta.contiguous(), A_prod_src.contiguous(), A_src_in_prod.contiguous(), A_src_in_sta.contiguous()


# ---- a1 -----------------------------------------------------------------------------------------------------------------

class InputExtractor(object):
    """Device-resident state of `extract_input_from_data` for one (station set, source grid)."""

    def __init__(self, plan, trv_times, ind_use, n_locs, max_t, kernel_sig_t, dt, node_sta=None, node_grid=None,
                 use_sign_input=False):
        dev = plan.device
        self.use_sign_input = bool(use_sign_input)
        self.plan, self.max_t, self.kernel_sig_t, self.dt = plan, float(max_t), float(kernel_sig_t), float(dt)
        self.n_locs = int(n_locs)
        ind_use = np.asarray(ind_use).astype('int')
        self.n_sta_use = len(ind_use)
        perm = -1 * np.ones(self.n_locs, dtype=np.int32)
        perm[ind_use] = np.arange(self.n_sta_use, dtype=np.int32)                       # process_utils.py:485-486
        self.sta_perm_host = perm                         # host copy: no device sync per window
        self.sta_perm = torch.from_numpy(perm).to(dev)
        self.ind_use = torch.from_numpy(ind_use.astype(np.int32)).to(dev)
        self.trv_times = trv_times if torch.is_tensor(trv_times) else torch.from_numpy(np.ascontiguousarray(trv_times))
        self.trv_times = self.trv_times.to(dev, torch.float32).contiguous()
        self.node_sta = None if node_sta is None else torch.as_tensor(node_sta).to(dev, torch.int32).contiguous()
        self.node_grid = None if node_grid is None else torch.as_tensor(node_grid).to(dev, torch.int32).contiguous()
        self._series = None
        self._day = None

    def params(self, t0):
        return ops.input_params(t0, self.max_t, self.kernel_sig_t, self.dt, self.n_locs, self.n_sta_use, self.use_sign_input)

    def set_day(self, P):
        """Keep a whole pick table [n,5] (sorted by time here) resident on the device."""
        P = np.asarray(P, dtype=np.float64)
        P = P[np.argsort(P[:, 0], kind='stable')]
        self._day = (P[:, 0].copy(), torch.from_numpy(np.ascontiguousarray(P)).to(self.plan.device))
        used = self.sta_perm_host[P[:, 1].astype('int')] >= 0
        self._used_cum = np.concatenate(([0], np.cumsum(used))).astype(np.int64)

    def used_cum(self):
        """Prefix counts, over the resident (time-sorted) pick table, of the picks on used stations."""
        return self._used_cum

    def window_rows(self, t0):
        """Row range of the resident pick table that can touch window t0 (process_utils.py:476)."""
        times = self._day[0]
        lo = np.searchsorted(times, t0 - 2.0 * self.kernel_sig_t, side='left')
        hi = np.searchsorted(times, t0 + self.max_t + 2.0 * self.kernel_sig_t, side='right')
        return int(lo), int(hi)

    def __call__(self, t0, picks=None, want_time_bin=False):
        """picks: CUDA float64 [n,5] (any superset of the window's picks); None = use the resident day table."""
        prm = self.params(t0)
        if picks is None:
            lo, hi = self.window_rows(float(t0))
            picks = self._day[1][lo:hi]
        if self._series is None or self._series.shape[2] != prm.n_ts:
            self._series = torch.empty((2, self.n_sta_use, prm.n_ts), dtype=torch.float32, device=self.plan.device)
        Slice, Mask, tb, _ = ops.input_scatter_fwd(self.plan, prm, picks, self.sta_perm, self.ind_use, self.trv_times,
                                                   self.node_sta, self.node_grid, want_time_bin, self._series)
        return (Slice, Mask, tb) if want_time_bin else (Slice, Mask)


def extract_pick_inputs_from_data(P_slice, locs, ind_use, time_samples, max_t, t_win=10.0):
    """process_utils.py:644-697 (use_batch False): picks of the window as (relative time, station index, phase, row)."""
    n_sta = len(locs)
    perm = -1 * np.ones(n_sta).astype('int')
    perm[ind_use] = np.arange(len(ind_use))
    ts = float(np.asarray(time_samples).reshape(-1)[0])
    sel = np.where(np.abs(P_slice[:, 0] - (ts + max_t / 2.0)) <= (t_win + max_t / 2.0))[0]
    meta = P_slice[sel, :]
    idx = perm[meta[:, 1].astype('int')]
    keep = np.where(idx > -1)[0]
    meta, idx = meta[keep], idx[keep]
    order = np.lexsort((meta[:, 0], idx))
    return [[meta[order, 0] - ts], [idx[order]], [meta[order, 4]], [meta[order]]]


_extractors = {}
_legacy_cache = {}


def _legacy_host_tables(locs, ind_use, arrivals, phase_labels, arrivals_tree, time_samples, max_t, pred_params):
    """Host half of extract_inputs_from_data_fixed_grids_with_phase_type: the pick selection of every sample, the merged
    offset time axes and the per-sample pick lists — a few thousand numbers, numpy."""
    arrivals = np.asarray(arrivals, dtype=np.float64)
    phase_labels = np.asarray(phase_labels)
    time_samples = np.asarray(time_samples, dtype=np.float64).reshape(-1)
    n_batch, n_sta = len(time_samples), int(locs.shape[0])
    t_win, kernel_sig_t = float(pred_params[0]), float(pred_params[1])
    if arrivals_tree is not None:                                                                          # :138
        lp = arrivals_tree.query_ball_point(time_samples.reshape(-1, 1) + max_t / 2.0, r=t_win + max_t / 2.0)
        lp = [np.array(list(l)).astype('int') for l in lp]
    else:
        lp = [np.where(np.abs(arrivals[:, 0] - (ts + max_t / 2.0)) <= t_win + max_t / 2.0)[0] for ts in time_samples]
    ind_sta_select = np.unique(ind_use)                                                                    # :152
    # One merged, sorted time axis for all samples and stations (:177-189): sample i and absolute station a are moved to
    # their own disjoint stretch of the axis, offset i * 1.5 max_t + a * 1.5 n_batch * 1.5 max_t, relative to the sample time.
    offset_per_batch = 1.5 * max_t
    offset_per_station = 1.5 * n_batch * offset_per_batch
    shifted, labels = [], []
    for i, rows in enumerate(lp):
        shift = -time_samples[i] + i * offset_per_batch + offset_per_station * arrivals[rows, 1]
        shifted.append(arrivals[rows, 0] + shift)
        labels.append(phase_labels[rows])
    shifted, labels = np.concatenate(shifted), np.concatenate(labels)
    by_time = np.argsort(shifted)                                                                          # :188
    shifted, labels = np.ascontiguousarray(shifted[by_time]), labels[by_time]
    axes = [shifted] + [np.ascontiguousarray(shifted[labels == ph]) for ph in (0, 1)]                      # :209-210
    # per-sample pick lists (:270-291): picks on selected stations, ordered by (station slot, time)
    slot_of = np.full(n_sta, -1.0)
    slot_of[ind_sta_select] = np.arange(len(ind_sta_select))
    lp_times, lp_stations, lp_phases, lp_meta = [], [], [], []
    for i, rows in enumerate(lp):
        slots = slot_of[arrivals[rows, 1].astype('int')]
        rows = rows[slots > -1]
        slots = slots[slots > -1]
        order = np.lexsort((arrivals[rows, 0], slots))
        rows = rows[order]
        lp_times.append(arrivals[rows, 0] - time_samples[i])
        lp_stations.append(slots[order])
        lp_phases.append(phase_labels[rows])
        lp_meta.append(arrivals[rows, :])
    return dict(axes=axes, lists=[lp_times, lp_stations, lp_phases, lp_meta], ind_sta_select=ind_sta_select,
                offset_per_batch=offset_per_batch, offset_per_station=offset_per_station, kernel_sig_t=kernel_sig_t)


def extract_inputs_from_data_fixed_grids_with_phase_type(trv, locs, ind_use, arrivals, phase_labels, arrivals_tree,
                                                          time_samples, x_grid, x_grid_trv, lat_range, lon_range,
                                                          depth_range, max_t, training_params, graph_params, pred_params,
                                                          ftrns1, ftrns2, verbose=False, device='cuda'):
    """Same call as process_utils.py:102 (the input features of `use_updated_input: False` and of training): returns
    `[Inpts, Masks], [lp_times, lp_stations, lp_phases, lp_meta]`, Inpts / Masks being fp32 CUDA tensors [G * n_sta, 4] per
    time sample.  The pick selection and the merged, offset time axis of :137-189 are a few thousand numbers and are built on
    the host with the reference's own numpy expressions; the G x S x 4 nearest-pick searches run in libgenie_b200."""
    import ctypes
    time_samples = np.asarray(time_samples, dtype=np.float64).reshape(-1)
    n_batch, n_spc, n_sta = len(time_samples), int(x_grid.shape[0]), int(locs.shape[0])
    h = _legacy_host_tables(locs, ind_use, arrivals, phase_labels, arrivals_tree, time_samples, max_t, pred_params)
    axes, ind_sta_select, kernel_sig_t = h['axes'], h['ind_sta_select'], h['kernel_sig_t']
    offset_per_batch, offset_per_station = h['offset_per_batch'], h['offset_per_station']
    dev = torch.device(device)
    if not dev.type == 'cuda':
        raise capi.GenieError('genie_b200 has no CPU path: device must be a CUDA device')
    key = (id(x_grid_trv), tuple(x_grid_trv.shape), str(dev))
    if _legacy_cache.get('key') != key:
        trv_dev = (x_grid_trv if torch.is_tensor(x_grid_trv) else torch.from_numpy(np.ascontiguousarray(x_grid_trv)))
        _legacy_cache.update(key=key, trv=trv_dev.to(dev).float().contiguous(), ref=x_grid_trv)
    trv_dev = _legacy_cache['trv']
    ind_dev = torch.from_numpy(ind_sta_select.astype(np.int32)).to(dev)
    axes_dev = [torch.from_numpy(a).to(dev) if len(a) else None for a in axes]
    n_node = n_spc * len(ind_sta_select)
    Slice = torch.empty((n_batch, n_node, 4), dtype=torch.float32, device=dev)
    Mask = torch.empty((n_batch, n_node, 4), dtype=torch.float32, device=dev)
    prm = capi.NearestParams(offset_per_batch, offset_per_station, kernel_sig_t, len(axes[0]), len(axes[1]), len(axes[2]),
                             n_batch, n_spc, n_sta, len(ind_sta_select))
    ptr = lambda t: capi.dptr(t, torch.float64) if t is not None else None
    with torch.cuda.device(dev):
        capi.check(capi.load().genie_input_nearest_fwd(
            ctypes.byref(prm), ptr(axes_dev[0]), ptr(axes_dev[1]), ptr(axes_dev[2]), capi.dptr(ind_dev, torch.int32),
            capi.dptr(trv_dev, torch.float32), capi.dptr(Slice), capi.dptr(Mask), capi.stream_ptr(dev)))
    lp_times, lp_stations, lp_phases, lp_meta = h['lists']
    return [[Slice[i] for i in range(n_batch)], [Mask[i] for i in range(n_batch)]], \
        [lp_times, lp_stations, lp_phases, lp_meta]


def extract_input_from_data(trv_pairwise, P, t0, ind_use, locs, x_grid, A_src_in_sta, trv_times=None, max_t=300.0,
                            kernel_sig_t=5.0, dt=0.2, batch_grids=False, use_asserts=True, verbose=False,
                            use_sign_input=False, return_embedding=False, device='cuda', plan=None):
    """Same call as process_utils.py:460; returns `[Inpts, Masks], [lp_times, lp_stations, lp_phases, lp_meta]` with
    Inpts/Masks CUDA tensors.  `trv_times` is the [G, n_locs, 2] travel-time table; without it (:594-596) the table is
    filled once from `trv_pairwise` for the pairs of `A_src_in_sta`.  `use_sign_input` (:610-614; the flag the
    day-processing script passes, process_continuous_days.py:776) and `return_embedding` (:571-572) behave as in the
    reference; `batch_grids` is 'Not implemented' there as well (:574-576).  Per-(grid, station set) state is cached."""
    if batch_grids:
        raise NotImplementedError('extract_input_from_data: batch_grids is not implemented (neither is it in the reference, '
                                  'process_utils.py:574-576)')
    t0v = float(np.asarray(t0).reshape(-1)[0])
    ind_use = np.asarray(ind_use).astype('int')
    A = A_src_in_sta.cpu().numpy() if torch.is_tensor(A_src_in_sta) else np.asarray(A_src_in_sta)
    # One extractor is cached.  The key holds the station subset itself and the sizes; the entry keeps `trv_times` and
    # `A_src_in_sta` (and `trv_pairwise`) referenced, so their id()s cannot be recycled by other objects while the entry lives.
    key = (id(trv_times) if trv_times is not None else ('pairwise', id(trv_pairwise)), id(A_src_in_sta), ind_use.tobytes(),
           int(locs.shape[0]), int(x_grid.shape[0]), float(max_t), float(kernel_sig_t), float(dt), str(device),
           id(plan) if plan is not None else None, bool(use_sign_input))
    ent = _extractors.get(key)
    if ent is None:
        G, S = int(x_grid.shape[0]), len(ind_use)
        if plan is None:
            # a node -> (station, grid) table is enough for a1; the plan only carries sizes here
            dense = A.shape[1] == S * G and np.array_equal(A[0], np.tile(np.arange(S), G)) and \
                np.array_equal(A[1], np.repeat(np.arange(G), S))
            empty = torch.zeros((2, 0), dtype=torch.long)
            if dense:
                plan = GraphPlan.cartesian(empty, empty, S, G, device=device)
            else:
                plan = GraphPlan.explicit(empty, empty, torch.from_numpy(A[1].astype(np.int64)), empty, A.shape[1], G,
                                          device=device)
        nodes = (None, None) if plan.mode == 0 else (A[0], A[1])
        table = trv_times
        if table is None:
            # :594-596: travel times of exactly the (station, source) pairs of A_src_in_sta from the caller's calculator
            dev = torch.device(device)
            tt = trv_pairwise(torch.Tensor(locs[ind_use]).to(dev)[torch.from_numpy(A[0]).to(dev)],
                              torch.Tensor(x_grid).to(dev)[torch.from_numpy(A[1]).to(dev)]).detach().float()
            table = torch.zeros((G, int(locs.shape[0]), 2), dtype=torch.float32, device=tt.device)
            table[torch.from_numpy(A[1]).to(tt.device), torch.from_numpy(ind_use[A[0]]).to(tt.device)] = tt
        ex = InputExtractor(plan, table, ind_use, locs.shape[0], max_t, kernel_sig_t, dt, nodes[0], nodes[1],
                            use_sign_input=use_sign_input)
        _extractors.clear()
        _extractors[key] = ent = (ex, trv_times, A_src_in_sta, plan, trv_pairwise)
    ex = ent[0]
    P = np.asarray(P, dtype=np.float64)
    keep = (P[:, 0] > (t0v - 2.0 * kernel_sig_t)) & (P[:, 0] < (t0v + max_t + 2.0 * kernel_sig_t))      # :476
    P_slice = P[keep]
    P_slice = P_slice[ex.sta_perm_host[P_slice[:, 1].astype('int')] > -1]                               # :480-482
    picks_dev = torch.from_numpy(np.ascontiguousarray(P_slice)).to(ex.plan.device)
    Slice, Mask = ex(t0v, picks_dev)
    if return_embedding:
        # :571-572: per-phase series of the stations that have a pick in the window, flattened [station][bin]
        ind_unique = np.sort(np.unique(P_slice[:, 1]).astype('int'))
        prm = ex.params(t0v)
        ser = ex._series.view(ex.n_sta_use, prm.n_ts, 2)[torch.from_numpy(ex.sta_perm_host[ind_unique].astype(np.int64)).to(Slice.device)]
        t_offset = 3.0 * kernel_sig_t
        abs_time_ref = np.arange(t0v - t_offset, t0v + max_t + t_offset + dt, dt)
        return (ser[:, :, 0].reshape(-1).clone(), ser[:, :, 1].reshape(-1).clone(), ind_unique, abs_time_ref, int(prm.n_ts),
                len(ind_unique))
    lp = extract_pick_inputs_from_data(P_slice, locs, ind_use, np.array([t0v]), max_t)
    return [[Slice], [Mask]], lp
