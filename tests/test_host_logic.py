"""CPU-only tests: the C-ABI library loads and exports what include/genie_b200.h declares; host-side plan logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden
from oracle import genie_oracle as go


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'genie_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(genie_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from genie_b200 import capi
    lib = capi.load()                                  # builds with nvcc if missing
    names = _declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(capi.SIGNATURES)           # the ctypes binding covers the whole header
    assert lib.genie_abi_version() == capi.ABI_VERSION
    assert lib.genie_frontend_packed_floats() > 20000
    assert lib.genie_launch_count() == 0                # nothing was launched (no GPU here)


def test_plan_create_rejects_bad_descriptors():
    from genie_b200 import capi
    lib = capi.load()
    h = ctypes.c_void_p()
    d = capi.GraphDesc()
    d.mode = 7
    assert lib.genie_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b'mode' in lib.genie_last_error()
    d.mode, d.n_sta, d.n_grid, d.n_prod = 0, 3, 4, 13
    assert lib.genie_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b'n_sta * n_grid' in lib.genie_last_error()
    assert lib.genie_plan_create(None, ctypes.byref(h)) != 0


def test_ops_fail_loudly_without_cuda():
    from genie_b200 import capi
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(capi.GenieError):
        capi.dptr(torch.zeros(4), torch.float32, 'x')


def test_csr_by_destination():
    from genie_b200.plan import csr_by_destination
    e = torch.tensor([[5, 1, 2, 0, 3], [2, 0, 2, 3, 0]])
    rowptr, col = csr_by_destination(e, 5)
    assert rowptr.tolist() == [0, 2, 2, 4, 5, 5]
    assert col.tolist() == [1, 3, 5, 2, 0]              # stable: source order kept inside each row
    with pytest.raises(ValueError):
        csr_by_destination(torch.tensor([[0], [9]]), 5)


@pytest.mark.parametrize('name', ['c1_10x100', 'small_6x40'])
def test_cartesian_pattern_detection(name):
    from genie_b200.plan import _is_cartesian
    d, _ = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, _ = go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']),
                                                                    int(d['k_spc']))
    got = _is_cartesian(A_ps, A_pg, A_sip[1], S, G)
    assert got is not None and torch.equal(got[0], A_sta) and torch.equal(got[1], A_src)
    bad = A_ps.clone()
    bad[0, -1] = (bad[0, -1] + 1) % (S * G)
    assert _is_cartesian(bad, A_pg, A_sip[1], S, G) is None
    bad = A_pg.clone()
    bad[1, 3] = (bad[1, 3] + 1) % (S * G)
    assert _is_cartesian(A_ps, bad, A_sip[1], S, G) is None


def test_knn_graph_matches_oracle_and_reference():
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    d, _ = load_golden('mid_36of40x300')
    A_sta, A_src = extract_inputs_adjacencies_cartesian(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']), int(d['k_spc']))
    assert np.array_equal(A_sta.numpy(), d['A_sta_sta']) and np.array_equal(A_src.numpy(), d['A_src_src'])


@pytest.mark.parametrize('t0,max_t,sig', [(200.0, 60.0, 3.0), (38940.0, 83.0, 3.5), (86399.7, 300.0, 3.0),
                                          (1234.5678, 127.0, 2.5)])
def test_input_params_follow_numpy_arange(t0, max_t, sig):
    """ref0 + i*ref_step and n_ts reproduce numpy.arange bit for bit (process_utils.py:502)."""
    from genie_b200 import ops
    dt = float(np.round(sig / 10.0, 2))
    prm = ops.input_params(t0, max_t, sig, dt, 10, 10)
    ref, n_ts = go.input_time_axis(t0, max_t, sig, dt)
    assert prm.n_ts == n_ts and prm.ref0 == ref[0]
    i = np.arange(n_ts, dtype=np.float64)
    assert np.array_equal(prm.ref0 + i * prm.ref_step, ref)
    assert prm.n_extra == int(np.ceil(3 * sig / dt))


def test_extract_pick_inputs_matches_reference_semantics():
    from genie_b200.process_utils import extract_pick_inputs_from_data
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    P = np.stack((rng.uniform(0, 500, 400), rng.integers(0, 12, 400).astype(float), np.ones(400), np.ones(400),
                  rng.integers(0, 2, 400).astype(float)), 1)
    ind_use = np.array([0, 2, 3, 5, 7, 8, 11])
    t, max_t = np.array([120.0]), 60.0
    lp_t, lp_s, lp_p, lp_m = extract_pick_inputs_from_data(P, np.zeros((12, 3)), ind_use, t, max_t)
    # process_utils.py:660-691 restated with the k-d tree the reference uses
    lp = cKDTree(P[:, [0]]).query_ball_point(t.reshape(-1, 1) + max_t / 2.0, r=10.0 + max_t / 2.0)[0]
    perm = -np.ones(12, dtype=int)
    perm[ind_use] = np.arange(len(ind_use))
    meta = P[sorted(lp)]
    idx = perm[meta[:, 1].astype(int)]
    meta, idx = meta[idx > -1], idx[idx > -1]
    order = np.lexsort((meta[:, 0], idx))
    assert np.array_equal(lp_m[0], meta[order]) and np.array_equal(lp_s[0], idx[order])
    assert np.array_equal(lp_t[0], meta[order, 0] - 120.0) and np.array_equal(lp_p[0], meta[order, 4])


def test_synthetic_network_is_seeded():
    from genie_b200 import synth
    a, b = synth.Network(20, 50, seed=3), synth.Network(20, 50, seed=3)
    assert np.array_equal(a.sta, b.sta) and np.array_equal(a.grid, b.grid)
    P = synth.make_picks(a, 0.0, 300.0, seed=1)
    assert P.shape[1] == 5 and np.all(np.diff(P[:, 0]) >= 0) and set(np.unique(P[:, 4])) <= {0.0, 1.0}
    assert a.travel_times().shape == (50, 20, 2) and a.travel_times().max() < a.max_moveout()


def test_locality_order_is_a_permutation_with_short_edges():
    """plan.locality_order: a permutation of the grid nodes (torch-convertible) that keeps kNN neighbours close."""
    import torch
    from genie_b200 import synth
    from genie_b200.plan import csr_by_destination, locality_order
    from genie_b200.process_utils import knn_graph
    net = synth.Network(20, 3000, seed=5)
    A = knn_graph(net.grid / 1000.0, 15)
    rp, col = csr_by_destination(A, 3000)
    order = locality_order(rp, col, 3000)
    assert order.dtype == np.int32 and order.flags['C_CONTIGUOUS']
    assert np.array_equal(np.sort(order), np.arange(3000))
    torch.from_numpy(order)                                   # must be convertible (no negative strides)
    pos = np.empty(3000, dtype=np.int64)
    pos[order] = np.arange(3000)
    d = np.abs(pos[A[0].numpy()] - pos[A[1].numpy()])
    assert d.max() < 1000                                     # graph bandwidth far below the node count
    assert np.array_equal(locality_order(rp[:1], col[:0], 0), np.arange(0))


@pytest.mark.parametrize('S,k', [(300, 15), (64, 8), (1000, 15), (33, 4)])
def test_station_tiles_reproduce_the_station_mean(S, k):
    """The tiling tables of the station-pass kernels (plan.station_tiles): every station in exactly one tile, staged
    rows cover every in-neighbour, and gathering through the tables equals the direct mean over in-edges."""
    from genie_b200 import synth
    from genie_b200.plan import ROWS_MAX, TILE_M, csr_by_destination, station_tiles
    from genie_b200.process_utils import knn_graph
    rng = np.random.default_rng(S)
    net = synth.Network(S, 10, seed=S, morton=False)            # unordered stations: the tiler must find the locality
    A = knn_graph(net.sta / 1000.0, k)
    A = A[:, rng.random(A.shape[1]) > 0.05]                     # ragged in-degrees
    A = A[:, A[1] != 7]                                         # one isolated station
    rowptr, col = csr_by_destination(A, S)
    st = station_tiles(rowptr, col, S)
    assert st is not None
    rows, meta, nbr, invdeg = st['rows'], st['meta'], st['nbr'], st['invdeg']
    assert meta[:, 0].max() <= TILE_M and meta[:, 1].max() <= ROWS_MAX
    own = np.concatenate([rows[t, :meta[t, 0]] for t in range(len(meta))])
    assert np.array_equal(np.sort(own), np.arange(S))
    x = rng.normal(size=(S, 5)).astype(np.float32)
    want = np.zeros_like(x)
    rp, cl = rowptr.numpy(), col.numpy()
    for s in range(S):
        if rp[s + 1] > rp[s]:
            want[s] = x[cl[rp[s]:rp[s + 1]]].mean(0)
    got = np.zeros_like(x)
    for t in range(len(meta)):
        staged = np.concatenate((x[rows[t]], np.zeros((1, 5), np.float32)))          # row ROWS_MAX = the zero row
        for r in range(meta[t, 0]):
            assert (nbr[t, r] <= ROWS_MAX).all() and ((nbr[t, r] < meta[t, 1]) | (nbr[t, r] == ROWS_MAX)).all()
            got[rows[t, r]] = staged[nbr[t, r].astype(np.int64)].sum(0) * invdeg[t, r]
    assert np.abs(got - want).max() < 1e-5
    assert invdeg[[t for t in range(len(meta)) if 7 in rows[t, :meta[t, 0]]][0]].min() == 0.0


def test_bisection_groups_partition_and_compactness():
    from genie_b200 import synth
    from genie_b200.plan import bisection_groups, csr_by_destination
    from genie_b200.process_utils import knn_graph
    G = 3000
    net = synth.Network(10, G, seed=3, morton=False)
    rowptr, col = csr_by_destination(knn_graph(net.grid / 1000.0, 15), G)
    ptr, nodes = bisection_groups(rowptr, col, G, 64)
    assert ptr[0] == 0 and ptr[-1] == G and np.array_equal(np.sort(nodes), np.arange(G))
    sizes = np.diff(ptr)
    assert sizes.max() <= 64 and sizes.min() >= 1
    rp, cl = rowptr.numpy(), col.numpy()
    union = [len(set(np.concatenate([cl[rp[g]:rp[g + 1]] for g in nodes[ptr[i]:ptr[i + 1]]]).tolist()))
             for i in range(len(sizes))]
    assert np.sum(sizes) * 15 / np.sum(union) > 3.0             # neighbour rows are re-used > 3x inside a group


@pytest.mark.parametrize('name', ['subgraph_14x60', 'subgraph_30x200', 'subgraph_12x40_ragged'])
def test_subgraph_builder_matches_reference(name):
    """extract_inputs_adjacencies_subgraph (process_utils.py:744-849): all six edge lists identical to the unmodified
    reference's, including the edge order."""
    from conftest import load_golden
    from genie_b200.process_utils import extract_inputs_adjacencies_subgraph
    d, _ = load_golden(name)
    out = extract_inputs_adjacencies_subgraph(d['sta'], d['grid'], lambda x: x, None, max_deg_offset=float(d['max_deg_offset']),
                                              k_nearest_pairs=int(d['k_nearest_pairs']), k_sta_edges=int(d['k_sta']),
                                              k_spc_edges=int(d['k_spc']))
    for got, key in zip(out, ('A_sta_sta', 'A_src_src', 'A_prod_sta_sta', 'A_prod_src_src', 'A_src_in_prod', 'A_src_in_sta')):
        assert got.dtype == torch.int64 and np.array_equal(got.numpy(), d[key]), key
    assert out[5].shape[1] < d['sta'].shape[0] * d['grid'].shape[0]


def test_dense_adjacency_builder_matches_reference():
    """extract_inputs_adjacencies (process_utils.py:701-742) with a station subset: the eight returned objects identical to the
    unmodified reference's, and the product lists are recognised as the Cartesian pattern."""
    from conftest import load_golden
    from genie_b200.process_utils import extract_inputs_adjacencies
    d, _ = load_golden('dense_adjacencies_9of12x30')
    out = extract_inputs_adjacencies(None, d['sta'], d['ind_use'], d['grid'], None, d['ref_t'], d['ptr_p'], d['ptr_s'],
                                     lambda x: x, [int(d['k_sta']), int(d['k_spc']), int(d['k_time'])])
    names = ('A_sta_sta', 'A_src_src', 'A_prod_sta_sta', 'A_prod_src_src', 'A_src_in_prod', 'A_edges_time_p', 'A_edges_time_s',
             'A_edges_ref')
    for got, key in zip(out, names):
        got = got.numpy() if torch.is_tensor(got) else np.asarray(got)
        assert got.dtype == d[key].dtype and np.array_equal(got, d[key]), key


@pytest.mark.parametrize('name', ['assoc_10x100', 'assoc_18of20x160', 'assoc_14of16x120_edges'])
def test_time_embedding_vectors_match_reference(name):
    """compute_time_embedding_vectors (process_utils.py:851-877) without k-d trees: the pointer tables the unmodified reference
    built for the association fixtures, entry for entry."""
    from conftest import load_golden
    from genie_b200.process_utils import compute_time_embedding_vectors
    d, _ = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sis = np.stack((np.tile(np.arange(S), G), np.repeat(np.arange(G), S)))
    sig = float(d['kernel_sig_t'])
    ep, es, dtp = compute_time_embedding_vectors(None, d['sta'][d['ind_use']], d['grid'], A_sis, float(d['max_t']),
                                                 dt_res=sig / 5.0, t_win=sig * 2.0, trv_out=d['tlatent'])
    assert np.array_equal(dtp, d['dt_partition'])
    assert ep.dtype == np.int64 and np.array_equal(ep, d['A_edges_p']) and np.array_equal(es, d['A_edges_s'])


def test_day_processor_window_pick_count_matches_reference_rule():
    """DayProcessor.window_pick_count (prefix sums over the resident pick table) against the reference's rule for skipping a
    window (process_continuous_days.py:787 via process_utils.py:476-481, 665), on a station subset, incl. empty windows and
    windows at the ends of the pick table."""
    from conftest import load_golden
    from genie_b200.process_utils import InputExtractor
    from genie_b200.streaming import DayProcessor

    class _Plan(object):
        device = torch.device('cpu')

    d, _ = load_golden('mid_36of40x300')
    P = d['picks']
    max_t, sig = float(d['max_t']), float(d['kernel_sig_t'])
    ex = InputExtractor(_Plan(), np.zeros((2, d['sta'].shape[0], 2), dtype=np.float32), d['ind_use'], d['sta'].shape[0], max_t,
                        sig, float(d['dt']))
    ex.set_day(P)
    dp = DayProcessor(None, ex, None, None, torch.zeros(1, 3))
    rng = np.random.default_rng(0)
    t0s = np.concatenate((rng.uniform(P[:, 0].min() - 2 * max_t, P[:, 0].max() + 2 * max_t, 300), P[:40, 0] - 2.0 * sig,
                          P[-40:, 0] - max_t - 2.0 * sig, [-1e6, 1e6]))
    n_zero = 0
    for t0 in t0s:
        sel = (P[:, 0] > (t0 - 2.0 * sig)) * (P[:, 0] < (t0 + max_t + 2.0 * sig))
        sel = sel * np.isin(P[:, 1].astype('int'), d['ind_use'])
        sel = sel * (np.abs(P[:, 0] - (t0 + max_t / 2.0)) <= (10.0 + max_t / 2.0))
        assert dp.window_pick_count(float(t0)) == int(sel.sum()), t0
        n_zero += int(sel.sum() == 0)
    assert 0 < n_zero < len(t0s)


@pytest.mark.parametrize('name', ['legacy_12of14x60', 'legacy_8x30_short'])
def test_legacy_host_tables_match_reference(name):
    """Host half of a1' (process_utils.py:137-189, :270-291): per-sample pick lists identical to the unmodified reference's,
    with and without the caller's k-d tree; the merged time axes are sorted and partition into the two phases."""
    from scipy.spatial import cKDTree
    from genie_b200.process_utils import _legacy_host_tables
    d, _ = load_golden(name)
    P = d['picks']
    for tree in (cKDTree(P[:, 0][:, None]), None):
        h = _legacy_host_tables(d['sta'], d['ind_use'], P, P[:, 4], tree, d['time_samples'], float(d['max_t']),
                                [float(d['t_win']), float(d['kernel_sig_t'])])
        for i in range(len(d['time_samples'])):
            for j, k in enumerate(('lp_times', 'lp_stations', 'lp_phases', 'lp_meta')):
                got = h['lists'][j][i]
                assert got.dtype == d['%s%d' % (k, i)].dtype and np.array_equal(got, d['%s%d' % (k, i)]), (k, i)
        a_all, a_p, a_s = h['axes']
        assert np.all(np.diff(a_all) >= 0) and len(a_p) + len(a_s) == len(a_all)
        assert np.array_equal(np.sort(np.concatenate((a_p, a_s))), a_all)


def test_edge_graph_operators_are_the_gather_and_the_mean():
    """genie_b200.relocation.EdgeGraph (GraphDD consumer): the four CSR matrices it hands to the gather kernel are x[source(e)],
    its transpose, the mean over a target's in-edges (0 for targets without edges) and its transpose — checked densely."""
    from genie_b200.relocation import EdgeGraph
    rng = np.random.default_rng(0)
    n_src, n_tgt, E = 13, 9, 40
    ei = torch.from_numpy(np.stack((rng.integers(0, n_src, E), rng.integers(0, n_tgt - 2, E)), 0)).long()      # two empty targets
    eg = EdgeGraph(ei, n_src, n_tgt)

    def dense(csr, n_rows, n_cols):
        rowptr, col, val = csr
        M = torch.zeros((n_rows, n_cols), dtype=torch.float64)
        for r in range(n_rows):
            for e in range(int(rowptr[r]), int(rowptr[r + 1])):
                M[r, int(col[e])] += float(val[e])
        return M
    Gm, Gt = dense(eg.g_fwd, E, n_src), dense(eg.g_rev, n_src, E)
    Mm, Mt = dense(eg.m_fwd, n_tgt, E), dense(eg.m_rev, E, n_tgt)
    assert torch.equal(Gm.t(), Gt) and torch.allclose(Mm.t(), Mt)
    x = torch.from_numpy(rng.normal(size=(n_src, 5)))
    assert torch.equal(Gm @ x, x[eg.src])
    msg = torch.from_numpy(rng.normal(size=(E, 5)))
    want = torch.zeros((n_tgt, 5), dtype=torch.float64).index_add_(0, eg.tgt, msg)
    want = want / torch.bincount(eg.tgt, minlength=n_tgt).clamp(min=1).unsqueeze(1)
    assert torch.allclose(Mm @ msg, want, atol=1e-6) and float((Mm @ msg)[-1].abs().max()) == 0.0
    assert torch.equal(torch.sort(eg.order)[0], torch.arange(E)) and torch.equal(ei[1][eg.order], eg.tgt)
