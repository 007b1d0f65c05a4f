"""Parity of the CUDA path (through the C-ABI) against the golden fixtures and the CPU oracle.  Needs a GPU.

Tolerance: BASELINE.json's north star — fp32 outputs within 1e-4 relative (max |a-b| / max |b|); the integer
pick -> time-bin map bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

GOLD = ['c1_10x100', 'mid_36of40x300', 'small_6x40', 'ferndale_t38940']
TOL = 1e-4


def _dev():
    if not torch.cuda.is_available():
        pytest.fail('these tests need a CUDA device (run with -m gpu on the B200 box)')
    return torch.device('cuda:0')


def _model(sd, dev, scale_rel, scale_t, updated_model=False, use_absolute_pos=False):
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, scale_rel=scale_rel, device=dev, updated_model=updated_model,
                                       use_absolute_pos=use_absolute_pos)
    m.load_state_dict(sd)
    m.TemporalAttention.scale_t = scale_t
    m.eval()
    return m


def _graphs(d):
    from oracle import genie_oracle as go
    return go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']), int(d['k_spc']))


@pytest.mark.parametrize('name', GOLD)
def test_input_scatter_matches_reference(name):
    """a1: Slice/Mask against the reference's extract_input_from_data; time-bin map bit-exact against the oracle."""
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import InputExtractor
    from oracle import genie_oracle as go
    dev = _dev()
    d, _ = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, _, _, _, A_sis = _graphs(d)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    ex = InputExtractor(plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']),
                        float(d['kernel_sig_t']), float(d['dt']))
    picks = torch.from_numpy(d['picks']).to(dev)
    Slice, Mask, tb = ex(float(d['t0']), picks, want_time_bin=True)
    _, _, parts = go.input_scatter(d['picks'], float(d['t0']), d['ind_use'], d['sta'].shape[0], A_sis.numpy(),
                                   d['trv_times'], float(d['max_t']), float(d['kernel_sig_t']), float(d['dt']),
                                   return_parts=True)
    assert np.array_equal(tb.cpu().numpy(), parts['time_bin'])            # integer map: bit-exact
    assert ex.params(float(d['t0'])).n_ts == int(d['n_ts'])
    assert np.abs(Slice.cpu().numpy() - d['Slice']).max() <= 1e-6
    assert np.array_equal(Mask.cpu().numpy(), d['Mask'])
    # same result through the explicit node table (sub-graph style addressing)
    ex2 = InputExtractor(plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']),
                         float(d['kernel_sig_t']), float(d['dt']), A_sis[0].numpy(), A_sis[1].numpy())
    S2, M2 = ex2(float(d['t0']), picks)
    assert torch.equal(S2, Slice) and torch.equal(M2, Mask)


def test_input_switches_match_reference():
    """extract_input_from_data with the switches of the reference's signature (process_utils.py:460): use_sign_input
    (:610-614, the flag process_continuous_days.py:776 passes), trv_times=None with a `trv_pairwise` calculator (:594-596)
    and return_embedding (:571-572), against the unmodified reference on a station subset."""
    from genie_b200.process_utils import extract_input_from_data
    dev = _dev()
    d, _ = load_golden('input_variants_12of14x60')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A = np.stack((np.tile(np.arange(S), G), np.repeat(np.arange(G), S)), axis=0)

    def trv_pairwise(sta, src):                       # evaluated on the host in fp32, exactly as the fixture's calculator
        dd = torch.norm(sta.cpu() - src.cpu(), dim=1, keepdim=True)
        return torch.cat((dd / 6000.0, dd / 3464.0), dim=1).to(sta.device)

    for tag, kw in (('plain', dict(trv_times=d['trv_times'])), ('sign', dict(trv_times=d['trv_times'], use_sign_input=True)),
                    ('pairwise', dict(trv_times=None)), ('pairwise_sign', dict(trv_times=None, use_sign_input=True))):
        [Inpts, Masks], _ = extract_input_from_data(trv_pairwise, d['picks'], np.array([float(d['t0'])]), d['ind_use'], d['sta'],
                                                    d['grid'], A, max_t=float(d['max_t']), kernel_sig_t=float(d['kernel_sig_t']),
                                                    dt=float(d['dt']), device=dev, **kw)
        got = Inpts[0].cpu().numpy()
        assert np.abs(got - d['Slice_' + tag]).max() <= 1e-6, tag
        assert np.array_equal(np.sign(got), np.sign(d['Slice_' + tag])), tag
        assert np.array_equal(Masks[0].cpu().numpy(), d['Mask_' + tag]), tag
    # the per-station series of return_embedding against the oracle's (pinned by the fixtures of test_oracle_golden.py)
    from oracle import genie_oracle as go
    emb = extract_input_from_data(None, d['picks'], np.array([float(d['t0'])]), d['ind_use'], d['sta'], d['grid'], A,
                                  trv_times=d['trv_times'], max_t=float(d['max_t']), kernel_sig_t=float(d['kernel_sig_t']),
                                  dt=float(d['dt']), return_embedding=True, device=dev)
    _, _, parts = go.input_scatter(d['picks'], float(d['t0']), d['ind_use'], d['sta'].shape[0], A, d['trv_times'],
                                   float(d['max_t']), float(d['kernel_sig_t']), float(d['dt']), return_parts=True)
    perm = -np.ones(d['sta'].shape[0], dtype=int)
    perm[d['ind_use']] = np.arange(S)
    assert emb[4] == parts['n_ts'] and emb[5] == len(emb[2]) and abs(emb[3][0] - parts['ref0']) == 0.0
    for ph in (0, 1):
        want = parts['series'][ph][perm[emb[2]]].reshape(-1)
        assert np.abs(emb[ph].cpu().numpy() - want).max() <= 1e-6


def test_window_runner_equals_the_two_step_path():
    """genie_window_fwd (a1 fused into layer 0: Slice / Mask never reach HBM, the mask rides bit-packed in the feature rows)
    behind streaming.WindowRunner — eager, captured in a CUDA graph, and with the window's picks staged from pinned host
    memory — against extract_input + forward_fixed_source on the same windows: y and x bit for bit, and the optional
    Slice / Mask copies equal the two-step inputs."""
    from genie_b200 import ops, synth
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.process_utils import InputExtractor, extract_inputs_adjacencies_cartesian
    from genie_b200.streaming import WindowRunner
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 100, 1500
    net = synth.Network(S, G, seed=3)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    attr = torch.from_numpy(net.read_in_offsets(30000.0)).to(dev)
    m = GCN_Detection_Network_extended(None, None, scale_rel=30000.0, device=dev).eval()
    m.load_state_dict(go.init_state(seed=9), strict=False)
    m.set_adjacencies_cartesian(A_sta, A_src, attr, S, G, device=dev)
    max_t = net.max_moveout()
    for sign in (False, True):
        ex = InputExtractor(m._plan, net.travel_times(), np.arange(S), S, max_t, 3.0, 0.3, use_sign_input=sign)
        P = synth.make_picks(net, 0.0, 600.0, seed=4, false_per_sta_min=4.0)
        ex.set_day(P)
        locs = torch.from_numpy(net.sta).float().to(dev)
        grid = torch.from_numpy(net.grid).float().to(dev)
        xq = torch.from_numpy(np.random.default_rng(1).uniform(0, net.width, (300, 3))).float().to(dev)
        xq[:, 2] = -xq[:, 2] / net.width * 40000.0
        tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
        runners = [WindowRunner(m, ex, locs, grid, xq, tq, use_graph=False),
                   WindowRunner(m, ex, locs, grid, xq, tq, use_graph=True),
                   WindowRunner(m, ex, locs, grid, xq, tq, use_graph=True, source='staged', max_window_picks=4096)]
        host = torch.from_numpy(ex._day[1].cpu().numpy()).pin_memory()
        for t0 in (12.0, 15.0, 18.0, 250.5, 251.25, 400.0):
            Slice, Mask = ex(t0)
            y0, x0 = m.forward_fixed_source(Slice, Mask, None, None, None, locs, grid, xq, tq)
            if sign:
                assert bool((Slice < 0).any())
            for r in runners:
                if r.source == 'staged':
                    lo, hi = ex.window_rows(t0)
                    y, x = r.run(t0, host[lo:hi])
                else:
                    y, x = r.run(t0)
                assert torch.equal(y, y0) and torch.equal(x, x0), (t0, r.use_graph, r.source)
        # the optional copies of the inputs
        wp = capi_window_block(ex, 250.5, dev)
        packed = m._packed_weights(dev)
        lo, hi = ex.window_rows(250.5)
        out = ops.window_fwd(m._plan, packed, wp(lo, hi), hi - lo, ex.params(0.0).n_extra, ex._day[1], ex.sta_perm, ex.ind_use,
                             ex.trv_times, torch.empty(2 * S * (ex.params(0.0).n_ts + 2), device=dev), ex.params(0.0).n_ts + 2,
                             attr, grid, 30000.0, want_inputs=True, want_latent=True, want_readin=True)
        Slice, Mask = ex(250.5)
        _, lat, rin = m.front_end(Slice, Mask, grid, want_latent=True, want_readin=True)
        assert torch.equal(out[3], Slice) and torch.equal(out[4], Mask)
        assert torch.equal(out[1], lat) and torch.equal(out[2], rin)


@pytest.mark.parametrize('S,G', [(100, 1500), (300, 400)])
def test_bf16_storage_mode_tracks_the_fp32_path(S, G):
    """GENIE_STORAGE_BF16 (BASELINE.json configs[1]: bf16 inference): the gathered intermediate rows kept as bf16, fp32
    arithmetic.  A SECOND mode — it never stands in for the 1e-4 parity tests: here it is held to 2e-2 of the fp32 path and
    of the oracle, must differ from fp32 (it really stores bf16), must agree bit for bit between the fused window call and
    the two-step call, and switching back to fp32 must restore the fp32 results exactly.  300 stations: three station tiles per
    grid node, i.e. the bf16 rows of the halo and the double-buffered staging."""
    from genie_b200 import synth
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.process_utils import InputExtractor, extract_inputs_adjacencies_cartesian, product_edge_lists
    from genie_b200.streaming import WindowRunner
    from oracle import genie_oracle as go
    dev = _dev()
    net = synth.Network(S, G, seed=3)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    attr = torch.from_numpy(net.read_in_offsets(30000.0)).to(dev)
    sd = go.init_state(seed=9)
    m = GCN_Detection_Network_extended(None, None, scale_rel=30000.0, device=dev).eval()
    m.load_state_dict(sd, strict=False)
    m.set_adjacencies_cartesian(A_sta, A_src, attr, S, G, device=dev)
    ex = InputExtractor(m._plan, net.travel_times(), np.arange(S), S, net.max_moveout(), 3.0, 0.3)
    ex.set_day(synth.make_picks(net, 0.0, 600.0, seed=4, false_per_sta_min=4.0))
    locs = torch.from_numpy(net.sta).float().to(dev)
    grid = torch.from_numpy(net.grid).float().to(dev)
    xq = torch.from_numpy(np.random.default_rng(1).uniform(0, net.width, (300, 3))).float().to(dev)
    xq[:, 2] = -xq[:, 2] / net.width * 40000.0
    tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
    t0 = 250.5
    Slice, Mask = ex(t0)
    y32, x32 = m.forward_fixed_source(Slice, Mask, None, None, None, locs, grid, xq, tq)
    xs32, _, r32 = m.front_end(Slice, Mask, grid, want_readin=True)
    ws32 = m._plan.workspace_bytes
    m.set_storage('bf16')
    assert m._plan.workspace_bytes < 0.8 * ws32
    y16, x16 = m.forward_fixed_source(Slice, Mask, None, None, None, locs, grid, xq, tq)
    xs16, _, r16 = m.front_end(Slice, Mask, grid, want_readin=True)
    assert not torch.equal(r16, r32)
    for a, b in ((r16, r32), (xs16, xs32), (y16, y32), (x16, x32)):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-2
    assert rel_err(r16.cpu().numpy(), r32.cpu().numpy()) > 1e-6
    # against the oracle
    A_ps, A_pg, A_sip, _ = product_edge_lists(A_sta, A_src, S, G)
    want = go.front_end(sd, Slice.cpu(), Mask.cpu(), A_ps, A_pg, attr.cpu(), A_sip, A_src, grid.cpu(), 30000.0)
    assert rel_err(xs16.cpu().numpy(), want.numpy()) < 2e-2
    assert rel_err(xs32.cpu().numpy(), want.numpy()) < 1e-4
    # fused window call == two-step call in this mode as well (eager and graph)
    for use_graph in (False, True):
        y, x = WindowRunner(m, ex, locs, grid, xq, tq, use_graph=use_graph).run(t0)
        assert torch.equal(y, y16) and torch.equal(x, x16)
    m.set_storage('fp32')
    y, x = m.forward_fixed_source(Slice, Mask, None, None, None, locs, grid, xq, tq)
    assert torch.equal(y, y32) and torch.equal(x, x32)


def capi_window_block(ex, t0, dev):
    """Device copy of a capi.WindowParams for window t0 (test helper)."""
    import ctypes
    from genie_b200 import capi

    def make(lo, hi):
        wp = capi.WindowParams()
        wp.prm = ex.params(t0)
        wp.pick_lo, wp.pick_hi = int(lo), int(hi)
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(wp), ctypes.sizeof(wp)), dtype=np.uint8).copy()
        return torch.from_numpy(raw).to(dev)
    return make


@pytest.mark.parametrize('name', GOLD)
def test_operators_match_reference(name):
    """a2, a3, a4 one by one, each fed the reference's own input for that stage."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    dev = _dev()
    d, sd = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, _ = _graphs(d)
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']))
    packed = m._packed_weights(dev)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    pos = torch.from_numpy(d['grid']).float().to(dev)
    plans = {
        'cartesian': GraphPlan.cartesian(A_sta, A_src, S, G, device=dev),
        'explicit': GraphPlan.explicit(A_ps, A_pg, A_sip[1], A_src, S * G, G, device=dev),
    }
    for kind, plan in plans.items():
        lat = ops.data_aggregation_fwd(plan, packed, t('Slice'), t('Mask'))
        assert rel_err(lat.cpu().numpy(), d['x_latent']) < TOL, kind
        r = ops.bipartite_readin_fwd(plan, packed, t('x_latent'), t('read_in_attr'), t('Mask'))
        assert rel_err(r.cpu().numpy(), d['read_in']) < TOL, kind
        for layer, (src, dst) in enumerate((('read_in', 'sa1'), ('sa1', 'sa2'), ('sa2', 'x_spatial'))):
            o = ops.spatial_aggregation_fwd(plan, packed, layer, t(src), pos, float(d['scale_rel']))
            assert rel_err(o.cpu().numpy(), d[dst]) < TOL, (kind, layer)
        xs, lat2, r2 = ops.frontend_fwd(plan, packed, t('Slice'), t('Mask'), t('read_in_attr'), pos,
                                        float(d['scale_rel']), want_latent=True, want_readin=True)
        assert rel_err(xs.cpu().numpy(), d['x_spatial']) < TOL, kind
        assert rel_err(lat2.cpu().numpy(), d['x_latent']) < TOL, kind
        assert rel_err(r2.cpu().numpy(), d['read_in']) < TOL, kind


@pytest.mark.parametrize('name', GOLD)
def test_forward_fixed_source_matches_reference(name):
    """The nn.Module surface: set_adjacencies with the reference's explicit edge lists + forward_fixed_source."""
    from genie_b200 import capi
    from oracle.refshim.torch_geometric.data import Data
    dev = _dev()
    d, sd = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']))
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev)
    grid = torch.from_numpy(d['grid']).float().to(dev)
    A_edges = Data(x=t('read_in_attr'), edge_index=A_sip.to(dev))
    m.set_adjacencies(A_ps.to(dev), A_pg.to(dev), A_edges, None, A_sis.to(dev), A_src.to(dev), None, None, None, None,
                      locs, grid)
    assert m._plan.mode == capi.GRAPH_CARTESIAN            # the index pattern of process_utils.py:720-722 is recognised
    n0 = capi.launch_count()
    y, x = m.forward_fixed_source(t('Slice'), t('Mask'), None, None, None, locs, grid,
                                  torch.from_numpy(d['x_query']).float().to(dev),
                                  torch.from_numpy(d['t_query']).float().reshape(-1, 1).to(dev))
    assert capi.launch_count() - n0 >= 10                  # our kernels ran (no silent fallback exists)
    assert rel_err(y.cpu().numpy(), d['y']) < TOL
    assert rel_err(x.cpu().numpy(), d['x']) < TOL


def _random_case(S, G, k_s, k_g, seed, dev):
    from genie_b200 import synth
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    net = synth.Network(S, G, seed=seed)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
    rng = np.random.default_rng(seed)
    P = S * G
    Slice = (rng.random((P, 4)) * (rng.random((P, 4)) < 0.35)).astype(np.float32)
    Mask = (np.abs(Slice) > 0.01).astype(np.float32)
    attr = net.read_in_offsets(np.array([net.width, net.width, 42000.0]))
    return net, A_sta, A_src, torch.from_numpy(Slice), torch.from_numpy(Mask), torch.from_numpy(attr)


@pytest.mark.parametrize('G,Q,T', [(7, 5, 3), (300, 77, 5), (1000, 130, 12), (3000, 40, 9), (200, 60, 31)])
def test_head_kernels_match_the_torch_restatement(G, Q, T):
    """genie_heads_grid_fwd / genie_heads_query_fwd against the torch restatement of SpatialDirect, SpatialAttention and
    TemporalAttention (module.py:251-331; itself checked against the reference's y, x in the golden tests): ragged sizes,
    fewer context nodes than k = 10, other numbers of query times; the last case has fewer (query, neighbour) pairs than
    context nodes and takes the per-edge kernel instead of the per-context-node table of the x_j projections; 31 query
    times run as two launches (25 per launch)."""
    from genie_b200 import ops
    from genie_b200.module import GCN_Detection_Network_extended
    dev = _dev()
    torch.manual_seed(G)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    for mod in (m.SpatialDirect, m.SpatialAttention, m.TemporalAttention):
        for name, prm in mod.named_parameters():
            if name.endswith('weight') and prm.numel() == 1:
                prm.data.fill_(0.1 + 0.5 * float(torch.rand(1)))            # distinct PReLU slopes
    x_spatial = torch.randn((G, 30), device=dev)
    ctx = torch.rand((G, 3), device=dev) * 50000.0
    xq = torch.rand((Q, 3), device=dev) * 50000.0
    tq = torch.linspace(-4.0, 4.0, T, device=dev).reshape(-1, 1)
    with torch.no_grad():
        want_y = m.TemporalAttention(m.SpatialDirect(x_spatial), tq)
        want_x = m.TemporalAttention(m.SpatialAttention(x_spatial, xq, ctx), tq)
        hw = ops.HeadsWeights(dev)
        hp, fold, Tn = hw.update(m, tq)
        edges = m.SpatialAttention._edges(xq, ctx, 10)
        nbr = m.SpatialAttention._nbr_table(edges, Q)
        assert nbr.shape == (Q, min(10, G))
        y, x = ops.heads_fwd(hw, hp, fold, Tn, x_spatial, ctx, xq, nbr, float(m.SpatialAttention.scale_rel))
    assert y.shape == want_y.shape and x.shape == want_x.shape
    assert rel_err(y.cpu().numpy(), want_y.cpu().numpy()) < 2e-5
    assert rel_err(x.cpu().numpy(), want_x.cpu().numpy()) < 2e-5


@pytest.mark.parametrize('name', ['legacy_12of14x60', 'legacy_8x30_short'])
def test_legacy_input_features_match_reference(name):
    """a1': extract_inputs_from_data_fixed_grids_with_phase_type (process_utils.py:102-308) through the reference's own
    call signature, against the unmodified reference: the fp64 feature rounded to fp32 and the 0.01 mask, bit for bit
    up to the last bit of the device's fp64 exp."""
    from scipy.spatial import cKDTree
    from genie_b200.process_utils import extract_inputs_from_data_fixed_grids_with_phase_type as legacy
    dev = _dev()
    d, _ = load_golden(name)
    P = d['picks']
    tree = cKDTree(P[:, 0][:, None])
    for use_tree in (tree, None):
        [Inpts, Masks], lists = legacy(None, d['sta'], d['ind_use'], P, P[:, 4], use_tree, d['time_samples'], d['grid'],
                                       d['trv_times'], None, None, None, float(d['max_t']), None, [8, 15, 10],
                                       [float(d['t_win']), float(d['kernel_sig_t'])], None, None, device=dev)
        for i in range(len(d['time_samples'])):
            want = d['Inpts%d' % i].astype(np.float32)
            got = Inpts[i].cpu().numpy()
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 2.0 * np.spacing(np.abs(want).max()) and np.mean(got == want) > 0.99
            m_got, m_want = Masks[i].cpu().numpy(), d['Masks%d' % i].astype(np.float32)
            edge = np.abs(d['Inpts%d' % i] - 0.01) < 1e-12          # a value within one ulp of the threshold may flip
            assert np.array_equal(m_got[~edge], m_want[~edge])
            for j, k in enumerate(('lp_times', 'lp_stations', 'lp_phases', 'lp_meta')):
                assert np.array_equal(lists[j][i], d['%s%d' % (k, i)])


def test_legacy_input_features_match_oracle_seeded():
    """The same at 100 x 2000 with three samples per call, against the pinned oracle (empty S axis: all picks are P)."""
    from genie_b200 import synth
    from genie_b200.process_utils import extract_inputs_from_data_fixed_grids_with_phase_type as legacy
    from oracle import genie_oracle as go
    dev = _dev()
    net = synth.Network(100, 2000, seed=2)
    P = synth.make_picks(net, 0.0, 900.0, seed=3)
    P = P[np.argsort(P[:, 0])]
    ind_use = np.arange(100)
    trv = net.travel_times()
    ts = np.array([200.0, 260.5, 410.25])
    for picks in (P, P[P[:, 4] == 0]):
        want, want_m = go.legacy_input_features(picks, picks[:, 4], ind_use, ts, trv, net.max_moveout(), 10.0, 3.0)
        [Inpts, Masks], _ = legacy(None, net.sta, ind_use, picks, picks[:, 4], None, ts, net.grid, trv, None, None, None,
                                   net.max_moveout(), None, [8, 15, 10], [10.0, 3.0], None, None, device=dev)
        for i in range(len(ts)):
            got, w = Inpts[i].cpu().numpy(), want[i].astype(np.float32)
            assert np.abs(got - w).max() <= 2.0 * np.spacing(1.0) and np.mean(got == w) > 0.99
            edge = np.abs(want[i] - 0.01) < 1e-12
            assert np.array_equal(Masks[i].cpu().numpy()[~edge], want_m[i].astype(np.float32)[~edge])


@pytest.mark.parametrize('name', ['c1_10x100_edges', 'mid_36of40x300_edges'])
def test_updated_model_definition_matches_reference(name):
    """a2': `use_updated_model_definition: True` (DataAggregationEdges, module.py:102-174, 1024-1186) through the
    nn.Module surface, against fixtures made by the unmodified reference (oracle/gen_golden.py synthetic_edges); the
    state_dict loads with strict=True (l1_t*_2 [30,68], l2_t*_2 [15,98])."""
    from genie_b200 import capi
    from oracle.refshim.torch_geometric.data import Data
    dev = _dev()
    d, sd = load_golden(name)
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']), updated_model=True)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev)
    grid = torch.from_numpy(d['grid']).float().to(dev)
    A_edges = Data(x=t('read_in_attr'), edge_index=A_sip.to(dev))
    m.set_adjacencies(A_ps.to(dev), A_pg.to(dev), A_edges, None, A_sis.to(dev), A_src.to(dev), None, None, None, None,
                      locs, grid)
    xs, lat, r = m.front_end(t('Slice'), t('Mask'), grid, want_latent=True, want_readin=True)
    assert rel_err(lat.cpu().numpy(), d['x_latent']) < TOL
    assert rel_err(r.cpu().numpy(), d['read_in']) < TOL
    assert rel_err(xs.cpu().numpy(), d['x_spatial']) < TOL
    y, x = m.forward_fixed_source(t('Slice'), t('Mask'), None, None, None, locs, grid,
                                  torch.from_numpy(d['x_query']).float().to(dev),
                                  torch.from_numpy(d['t_query']).float().reshape(-1, 1).to(dev))
    assert rel_err(y.cpu().numpy(), d['y']) < TOL
    assert rel_err(x.cpu().numpy(), d['x']) < TOL


@pytest.mark.parametrize('S,G,tiling', [(150, 260, True), (150, 260, False), (20, 90, True)])
def test_updated_model_definition_matches_oracle_seeded(S, G, tiling):
    """The edge-feature terms in the split tensor-core station pass (tiling tables), in the generic kernels (no tables /
    fewer than 32 stations) and on an EXPLICIT plan (per-product-node terms), against the pinned oracle."""
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import product_edge_lists
    from genie_b200 import capi
    from oracle import genie_oracle as go
    dev = _dev()
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 29, dev)
    sd = go.init_state(seed=9, edges=True)
    A_ps, A_pg, A_sip, A_sis = product_edge_lists(A_sta, A_src, S, G)
    sta, grid = torch.from_numpy(net.sta).float(), torch.from_numpy(net.grid).float()
    pos_rel = go.edge_features(sta, grid, A_sis, A_ps, A_pg, 30000.0)
    want, parts = go.front_end(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid, 30000.0, return_parts=True,
                               pos_rel=pos_rel)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev, updated_model=True)
    m.load_state_dict(sd, strict=False)
    m.eval()
    plans = [GraphPlan.cartesian(A_sta, A_src, S, G, device=dev, tiling=tiling)]
    if not tiling:
        plans.append(GraphPlan.explicit(A_ps, A_pg, A_sip[1], A_src, S * G, G, device=dev))
    for plan in plans:
        assert (plan.tiles is not None) == tiling          # tiling tables exist for every station count >= 2
        m._plan, m._read_in_attr = plan, attr.to(dev)
        m._set_edge_means(sta.to(dev), grid.to(dev), A_sis.to(dev))
        with torch.no_grad():
            xs, lat, r = m.front_end(Slice.to(dev), Mask.to(dev), grid.to(dev), want_latent=True, want_readin=True)
        assert rel_err(lat.cpu().numpy(), parts['x_latent'].numpy()) < TOL
        assert rel_err(r.cpu().numpy(), parts['read_in'].numpy()) < TOL
        assert rel_err(xs.cpu().numpy(), want.numpy()) < TOL


@pytest.mark.parametrize('S,G,k_s,k_g', [(100, 500, 15, 15), (37, 211, 8, 15), (130, 64, 10, 5)])
def test_front_end_matches_oracle_seeded(S, G, k_s, k_g):
    """Seeded synthetic networks against the CPU oracle computed on the spot (ragged tiles: P not a multiple of 128)."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import product_edge_lists
    from oracle import genie_oracle as go
    dev = _dev()
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, k_s, k_g, 11, dev)
    sd = go.init_state(seed=2)
    A_ps, A_pg, A_sip, _ = product_edge_lists(A_sta, A_src, S, G)
    grid = torch.from_numpy(net.grid).float()
    want, parts = go.front_end(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid, 30000.0, return_parts=True)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    packed = m._packed_weights(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    xs, lat, r = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid.to(dev), 30000.0,
                                  want_latent=True, want_readin=True)
    assert rel_err(lat.cpu().numpy(), parts['x_latent'].numpy()) < TOL
    assert rel_err(r.cpu().numpy(), parts['read_in'].numpy()) < TOL
    assert rel_err(xs.cpu().numpy(), want.numpy()) < TOL


@pytest.mark.parametrize('a11,a12', [(0.07, 0.9), (1.7, 0.02), (5e-4, 0.3), (0.3, 0.0), (-0.2, 0.25)])
def test_prelu_slopes_pick_the_kernel_family(a11, a12):
    """The tensor-core station pass stages PReLU11(tr0) and PReLU12(tr0) and recovers tr0 from them, which needs both slopes
    in (1e-3, 1e3); otherwise the generic FFMA kernels run (decided on the device from the packed weights, layout.h TCS_OK).
    Either way the result is the reference's (module.py:88-96)."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import product_edge_lists
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 150, 260
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 23, dev)
    sd = go.init_state(seed=7)
    sd['DataAggregation.activate11.weight'] = torch.full((1,), a11)
    sd['DataAggregation.activate12.weight'] = torch.full((1,), a12)
    A_ps, A_pg, A_sip, _ = product_edge_lists(A_sta, A_src, S, G)
    grid = torch.from_numpy(net.grid).float()
    want, parts = go.front_end(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid, 30000.0, return_parts=True)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    packed = m._packed_weights(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    assert plan.tiles is not None
    xs, lat, r = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid.to(dev), 30000.0,
                                  want_latent=True, want_readin=True)
    assert rel_err(lat.cpu().numpy(), parts['x_latent'].numpy()) < TOL
    assert rel_err(r.cpu().numpy(), parts['read_in'].numpy()) < TOL
    assert rel_err(xs.cpu().numpy(), want.numpy()) < TOL


def test_one_pass_kernels_match_split_kernels():
    """Plans without tiling tables run the one-pass kernels (tcgen05 layer 1 with L2 gathers, FFMA layer 2 with atomics);
    plans with them run the split source-pass / station-pass kernels.  Same mathematics, different tiling."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 300, 700                                          # 3 station tiles with halos, 11 grid groups
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 17, dev)
    sd = go.init_state(seed=5)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    packed = m._packed_weights(dev)
    grid = torch.from_numpy(net.grid).float().to(dev)
    outs = []
    for tiling in (True, False):
        plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev, tiling=tiling)
        assert (plan.tiles is not None) == tiling
        if tiling:
            assert plan.tiles['n_tiles'] == 3
        outs.append(ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid, 30000.0,
                                     want_latent=True, want_readin=True))
        lat_only = ops.data_aggregation_fwd(plan, packed, Slice.to(dev), Mask.to(dev))
        assert rel_err(lat_only.cpu().numpy(), outs[-1][1].cpu().numpy()) < 1e-5
    for a, b in zip(outs[0], outs[1]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-5
    # the station pass has no atomics: bit-reproducible run to run
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    again = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid, 30000.0, want_latent=True,
                             want_readin=True)
    for a, b in zip(outs[0], again):
        assert torch.equal(a, b)


def test_irregular_explicit_graph_matches_oracle():
    """Sub-graph style product graph: variable in-degree, isolated nodes (mean of nothing = 0), unsorted edge order."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from oracle import genie_oracle as go
    dev = _dev()
    rng = np.random.default_rng(5)
    G, P = 57, 3001
    prod_grid = np.sort(rng.integers(0, G, P))
    prod_grid[:3] = 0
    E1, E2 = 7 * P, 11 * P
    A1 = torch.from_numpy(np.stack((rng.integers(0, P, E1), rng.integers(0, P - 40, E1)), 0)).long()   # last 40 isolated
    A2 = torch.from_numpy(np.stack((rng.integers(0, P, E2), rng.integers(40, P, E2)), 0)).long()       # first 40 isolated
    A_src = torch.from_numpy(np.stack((rng.integers(0, G, 9 * G), rng.integers(0, G - 1, 9 * G)), 0)).long()
    Slice = torch.from_numpy((rng.random((P, 4)) * (rng.random((P, 4)) < 0.4)).astype(np.float32))
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.from_numpy(rng.normal(size=(P, 3)).astype(np.float32))
    pos = torch.from_numpy(rng.uniform(0, 1e5, (G, 3)).astype(np.float32))
    read_idx = torch.stack((torch.arange(P), torch.from_numpy(prod_grid).long()), 0)
    sd = go.init_state(seed=3)
    # the oracle's read-in sizes its output by max target + 1 (module.py:227): make the last grid node present
    assert prod_grid.max() == G - 1
    want, parts = go.front_end(sd, Slice, Mask, A1, A2, attr, read_idx, A_src, pos, 30000.0, return_parts=True)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    packed = m._packed_weights(dev)
    plan = GraphPlan.from_edge_lists(A1.to(dev), A2.to(dev), read_idx.to(dev), A_src.to(dev), 1, G, device=dev)
    assert plan.mode == 1
    xs, lat, r = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), pos.to(dev), 30000.0,
                                  want_latent=True, want_readin=True)
    assert rel_err(lat.cpu().numpy(), parts['x_latent'].numpy()) < TOL
    assert rel_err(r.cpu().numpy(), parts['read_in'].numpy()) < TOL
    assert rel_err(xs.cpu().numpy(), want.numpy()) < TOL


def test_station_permutation_invariance_large():
    """Size-independent property at a size the oracle cannot reach quickly: relabelling the stations (graphs, inputs and
    read-in features permuted consistently) permutes x_latent and leaves the grid-level output unchanged."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 1000, 1500
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 21, dev)
    sd = go.init_state(seed=4)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    packed = m._packed_weights(dev)
    grid = torch.from_numpy(net.grid).float().to(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    xs, lat, _ = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid, 30000.0,
                                  want_latent=True)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(S))          # new station id -> old station id
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(S)
    A_sta_p = inv[A_sta]                                                       # relabel both endpoints
    pm = lambda x: x.view(G, S, -1)[:, perm, :].reshape(G * S, -1).contiguous()
    plan_p = GraphPlan.cartesian(A_sta_p, A_src, S, G, device=dev)
    xs_p, lat_p, _ = ops.frontend_fwd(plan_p, packed, pm(Slice).to(dev), pm(Mask).to(dev), pm(attr).to(dev), grid,
                                      30000.0, want_latent=True)
    assert rel_err(lat_p.cpu().numpy(), pm(lat.cpu()).numpy()) < 1e-5
    assert rel_err(xs_p.cpu().numpy(), xs.cpu().numpy()) < 1e-5
    # run-to-run: everything but the cross-tile read-in atomics is order-fixed
    xs2, lat2, _ = ops.frontend_fwd(plan, packed, Slice.to(dev), Mask.to(dev), attr.to(dev), grid, 30000.0,
                                    want_latent=True)
    assert torch.equal(lat2, lat)
    assert rel_err(xs2.cpu().numpy(), xs.cpu().numpy()) < 1e-5


def test_full_size_c4_split_equals_one_pass():
    """BASELINE.json's headline size (1000 stations x 50000 grid nodes, P = 5e7) cannot be checked against the CPU oracle;
    instead the two independent kernel families are compared on it — the split source-pass / two-pipeline tcgen05 station
    pass (tiling tables) and the one-pass kernels (no tables: L2 gathers, FFMA read-in with atomics) — plus run-to-run bit
    equality of the atomic-free split path.  Inputs are generated on the device."""
    from genie_b200 import ops, synth
    from genie_b200.plan import GraphPlan
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    from oracle import genie_oracle as go
    dev = _dev()
    if torch.cuda.get_device_properties(dev).total_memory < 100e9:
        pytest.skip('needs ~60 GB of device memory')
    S, G = 1000, 50000
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    gen = torch.Generator(device=dev).manual_seed(5)
    P = S * G
    Slice = torch.rand((P, 4), device=dev, generator=gen) * (torch.rand((P, 4), device=dev, generator=gen) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.rand((P, 3), device=dev, generator=gen) - 0.5
    grid = torch.from_numpy(net.grid).float().to(dev)
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(go.init_state(seed=6), strict=False)
    packed = m._packed_weights(dev)
    outs = []
    for tiling in (True, False):
        plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev, tiling=tiling)
        xs, _, r = ops.frontend_fwd(plan, packed, Slice, Mask, attr, grid, 30000.0, want_readin=True)
        if tiling:
            xs2, _, r2 = ops.frontend_fwd(plan, packed, Slice, Mask, attr, grid, 30000.0, want_readin=True)
            assert torch.equal(xs, xs2) and torch.equal(r, r2)
        outs.append((xs.cpu().numpy(), r.cpu().numpy()))
        del plan
        torch.cuda.empty_cache()
    assert np.isfinite(outs[0][0]).all() and np.abs(outs[0][0]).max() > 0
    assert rel_err(outs[0][1], outs[1][1]) < 2e-5
    assert rel_err(outs[0][0], outs[1][0]) < 2e-5


@pytest.mark.parametrize('workload,window', [('c2_100x5000_dense', 3), ('c4_1000x50000_dense', 7)])
def test_window_against_oracle_on_sampled_closure(workload, window):
    """The headline size anchored on the ORACLE (not on a second kernel family): one of bench.py's windows at 1000 stations x
    50000 grid nodes with its real a1 inputs; for 20 sampled grid nodes (all stations) the integer time bins (exact), Slice /
    Mask, x_latent rows and Bipartite_ReadIn rows equal the CPU oracle run on their 2-hop source-graph closure
    (module.py:85-98, 224-229; process_utils.py:599-629) within 1e-4 under the element-wise (row-scaled) metric, and y / x of
    forward_fixed_source equal the oracle's SpatialAggregation + heads on the full grid.  bench.py runs the same check on a
    timed window after every run (`parity_check` in its JSON line)."""
    import bench
    dev = _dev()
    if workload.startswith('c4') and torch.cuda.get_device_properties(dev).total_memory < 100e9:
        pytest.skip('needs ~60 GB of device memory')
    wl = bench.Workload(workload, dev, day_s=1500.0)
    wl.runners(use_graph=workload.startswith('c2'))         # C2 through the CUDA-graph replay, C4 eager (as bench.py runs them)
    rep = bench.closure_parity(wl, window)
    assert rep['nodes'] >= 16 and rep['picks_in_window'] > 0
    assert rep['time_bin_equal'] and rep['mask_equal'] and rep['slice_max_abs'] <= 1e-6, rep
    assert rep['x_latent_rowwise_rel'] < 1e-4 and rep['read_in_rowwise_rel'] < 1e-4, rep
    assert rep['y_rel'] < 1e-4 and rep['x_rel'] < 1e-4, rep
    assert rep['fused_equals_two_step'], rep
    assert rep['ok']
    del wl
    torch.cuda.empty_cache()


def test_config2_against_oracle():
    """BASELINE.json configs[1]: 100 stations x 5000 grid nodes, k = 15 / 15 (P = 500 000), full front end vs oracle."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import product_edge_lists
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 100, 5000
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 31, dev)
    sd = go.init_state(seed=2)
    A_ps, A_pg, A_sip, _ = product_edge_lists(A_sta, A_src, S, G)
    grid = torch.from_numpy(net.grid).float()
    want, parts = go.front_end(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid, 30000.0, return_parts=True)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    plan = GraphPlan.from_edge_lists(A_ps.to(dev), A_pg.to(dev), A_sip.to(dev), A_src.to(dev), S, G, device=dev)
    assert plan.mode == 0
    xs, lat, r = ops.frontend_fwd(plan, m._packed_weights(dev), Slice.to(dev), Mask.to(dev), attr.to(dev),
                                  grid.to(dev), 30000.0, want_latent=True, want_readin=True)
    assert rel_err(lat.cpu().numpy(), parts['x_latent'].numpy()) < TOL
    assert rel_err(r.cpu().numpy(), parts['read_in'].numpy()) < TOL
    assert rel_err(xs.cpu().numpy(), want.numpy()) < TOL


def test_sharded_front_end_single_process():
    """Grid sharding with the product backend (genie_b200.sharded.CudaBackend): two ranks emulated one after the other on
    one GPU (the halo exchange done by hand from GridPartition's lists) must reproduce the unsharded front end."""
    from genie_b200 import ops
    from genie_b200.plan import GraphPlan
    from genie_b200.sharded import CudaBackend, GridPartition
    from oracle import genie_oracle as go
    dev = _dev()
    S, G, world = 160, 900, 2
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 41, dev)
    sd = go.init_state(seed=7)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    grid = torch.from_numpy(net.grid).float().to(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    want_xs, _, want_r = ops.frontend_fwd(plan, m._packed_weights(dev), Slice.to(dev), Mask.to(dev), attr.to(dev), grid,
                                          30000.0, want_readin=True)
    part = GridPartition(A_src, G, world)
    bes, nodes = [], []
    for r in range(world):
        nd = torch.from_numpy(part.local_nodes(r))
        loc = lambda x: x.view(G, S, -1).index_select(0, nd).reshape(len(nd) * S, -1).contiguous().to(dev)
        be = CudaBackend(m, A_sta, part.local_graph(r), S, len(nd), len(part.owned[r]), loc(attr), A_src, G, dev)
        assert be.plan.tiles is not None and be.plan.n_grid_owned == len(part.owned[r])
        be.layer1(loc(Slice), loc(Mask))
        bes.append(be)
        nodes.append(nd)
    rows = [be.message_rows() for be in bes]
    g2l = [{int(g): i for i, g in enumerate(nd.tolist())} for nd in nodes]
    for r in range(world):                                   # what all_to_all_single does in ShardedFrontEnd.exchange
        n_own = len(part.owned[r])
        rows[r][n_own:] = float('nan')
        for i, g in enumerate(part.halo[r].tolist()):
            q = int(part.owner[g])
            rows[r][n_own + i] = rows[q][g2l[q][g]]
    read_in = torch.empty((G, 15), device=dev)
    for r in range(world):
        read_in[torch.from_numpy(part.owned[r]).to(dev)] = bes[r].layer2_readin()
    assert rel_err(read_in.cpu().numpy(), want_r.cpu().numpy()) < 1e-5
    xs = bes[0].spatial(read_in, grid, 30000.0)
    assert rel_err(xs.cpu().numpy(), want_xs.cpu().numpy()) < 1e-5


class _RawDeviceBytes(object):
    """Zero-copy uint8 view of raw device memory for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 2}


@pytest.mark.parametrize('storage', ['fp32', 'bf16'])
def test_sharded_halo_rows_over_peer_stores(storage):
    """genie_plan_set_halo_export: the layer-1 station pass of every rank stores the v_b rows its peers hold as halo straight
    into the peers' landing buffers (genie_peer_alloc), the layer-2 source pass reads them from there.  Three ranks emulated in
    one process on one GPU (the peers' buffers are then plain device pointers); the landing buffers must equal the owners' rows
    bit for bit and the read-in rows the unsharded front end's."""
    import ctypes
    from genie_b200 import capi, ops
    from genie_b200.plan import GraphPlan
    from genie_b200.sharded import CudaBackend, GridPartition, PeerHalo
    from oracle import genie_oracle as go
    dev = _dev()
    S, G, world = 160, 900, 3
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 15, 15, 43, dev)
    sd = go.init_state(seed=7)
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(sd, strict=False)
    grid = torch.from_numpy(net.grid).float().to(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    plan.set_storage(storage)
    want_r = ops.frontend_fwd(plan, m._packed_weights(dev), Slice.to(dev), Mask.to(dev), attr.to(dev), grid, 30000.0,
                              want_readin=True)[2].clone()
    part = GridPartition(A_src, G, world)
    lib = capi.load()
    row_bytes = S * 16 * (2 if storage == 'bf16' else 4)
    bes, bufs, keep = [], [], []
    for r in range(world):
        nd = torch.from_numpy(part.local_nodes(r))
        loc = lambda x: x.view(G, S, -1).index_select(0, nd).reshape(len(nd) * S, -1).contiguous().to(dev)
        be = CudaBackend(m, A_sta, part.local_graph(r), S, len(nd), len(part.owned[r]), loc(attr), A_src, G, dev,
                         grid_groups=part.local_groups(r))
        be.plan.set_storage(storage)
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(capi.PEER_HANDLE_BYTES)
        capi.check(lib.genie_peer_alloc(max(len(part.halo[r]), 1) * row_bytes, ctypes.byref(ptr), handle))
        bes.append((be, loc(Slice), loc(Mask)))
        bufs.append(ptr.value)
    try:
        peer_base = torch.tensor(bufs, dtype=torch.int64, device=dev)
        for r in range(world):
            ep, eq, er = (torch.from_numpy(a).to(dev) for a in PeerHalo.export_tables(part, r))
            assert int(ep[-1]) == eq.numel() and eq.numel() > 0
            bes[r][0].plan.set_halo_export(ep, eq, er, peer_base, bufs[r])
        for be, sl, mk in bes:                               # every rank's layer 1 (fills the peers' landing buffers) ...
            be.layer1(sl, mk)
        torch.cuda.synchronize()
        # ... the landing buffer of rank r = the owners' rows of its halo nodes
        rows = [be.message_rows() for be, _, _ in bes]
        for r in range(world):
            n_halo = len(part.halo[r])
            got = torch.as_tensor(_RawDeviceBytes(bufs[r], n_halo * row_bytes), device=dev).view(rows[r].dtype).view(n_halo, -1)
            for i, g in enumerate(part.halo[r].tolist()):
                q = int(part.owner[g])
                j = int(np.nonzero(part.owned[q] == g)[0][0])
                assert torch.equal(got[i], rows[q][j]), (r, i, g)
            rows[r][len(part.owned[r]):] = float('nan') if storage == 'fp32' else 0xff    # the local halo rows are NOT used
        read_in = torch.empty((G, 15), device=dev)
        for r in range(world):
            read_in[torch.from_numpy(part.owned[r]).to(dev)] = bes[r][0].layer2_readin()
        assert rel_err(read_in.cpu().numpy(), want_r.cpu().numpy()) < (1e-5 if storage == 'fp32' else 2e-3)
        assert torch.isfinite(read_in).all()
    finally:
        torch.cuda.synchronize()
        for r in range(world):
            bes[r][0].plan.set_halo_export(None, None, None, None, None)
            lib.genie_peer_free(ctypes.c_void_p(bufs[r]))


# ---- association branch: forward_fixed / forward ------------------------------------------------------

ASSOC = ['assoc_10x100', 'assoc_18of20x160']
ASSOC_VARIANTS = ['assoc_14of16x120_edges', 'assoc_14of16x120_abspos']       # use_updated_model_definition / use_absolute_pos


def _assoc_setup(d, sd, dev, name=''):
    from oracle.refshim.torch_geometric.data import Data
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']), updated_model=name.endswith('_edges'),
               use_absolute_pos=name.endswith('_abspos'))
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev)
    grid = torch.from_numpy(d['grid']).float().to(dev)
    A_edges = Data(x=t('read_in_attr'), edge_index=A_sip.to(dev))
    A_Lg = Data(x=t('read_in_attr'), edge_index=A_sip.flip(0).contiguous().to(dev))
    graphs = (A_ps.to(dev), A_pg.to(dev), A_edges, A_Lg, A_sis.to(dev), A_src.to(dev), t('A_edges_p'), t('A_edges_s'),
              t('dt_partition').float(), t('tlatent').float())
    window = (t('tpick').float(), t('ipick').long(), t('phase_label').long().reshape(-1, 1), locs, grid,
              t('x_query').float(), t('x_query_src').float(), t('t_query').float().reshape(-1, 1), t('tq_sample').float(),
              t('trv_out_q').float())
    return m, graphs, window, locs, grid


@pytest.mark.parametrize('name', ASSOC + ASSOC_VARIANTS)
def test_forward_fixed_matches_reference(name):
    """forward_fixed (module.py:963-997) through the nn.Module surface against the unmodified reference, plus the
    intermediate tensors of the association kernels through the operator-level C-ABI wrappers."""
    from genie_b200 import capi, ops
    dev = _dev()
    d, sd = load_golden(name)
    m, graphs, window, locs, grid = _assoc_setup(d, sd, dev, name)
    m.set_adjacencies(*graphs, locs, grid)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    n0 = capi.launch_count()
    y, x, arv_p, arv_s = m.forward_fixed(t('Slice'), t('Mask'), *window)
    assert capi.launch_count() - n0 >= 15
    assert rel_err(y.cpu().numpy(), d['y']) < TOL and rel_err(x.cpu().numpy(), d['x']) < TOL
    assert arv_p.shape == d['arv_p'].shape and arv_s.shape == d['arv_s'].shape
    assert rel_err(arv_p.cpu().numpy(), d['arv_p']) < TOL
    assert rel_err(arv_s.cpu().numpy(), d['arv_s']) < TOL
    # operator level
    packed = m._assoc_w.update(m, m._update_assoc_terms(locs, grid))
    s_rows, s0, mask_out = ops.assoc_product_fwd(m._plan, packed, t('x_spatial'), t('y').reshape(len(d['grid']), -1),
                                                 t('read_in_attr'), t('x_latent'), t('Mask'), want_parts=True)
    S = len(d['ind_use'])
    assert np.array_equal(mask_out.cpu().numpy()[np.arange(S * len(d['grid'])) // S][:, None], d['mask_out_1'])
    assert rel_err(s0.cpu().numpy(), d['assoc_s0']) < TOL
    s = torch.cat((s_rows[:, 0:15], s_rows[:, 16:31]), dim=1)
    assert rel_err(s.cpu().numpy(), d['assoc_s']) < TOL
    assert not s_rows[:, 15].any() and not s_rows[:, 31].any()
    arrival = ops.assoc_collapse_fwd(packed, s_rows, t('A_edges_p'), t('A_edges_s'), t('dt_partition').float(),
                                     t('tlatent').float(), t('tpick').float(), t('ipick').long(), t('phase_label').float(), S,
                                     float(d['eps']))
    assert arrival.shape == (len(d['tpick']) + 1, 30) and not arrival[-1].any()
    assert rel_err(arrival[:-1, :15].cpu().numpy(), d['arv_p_embed']) < TOL
    assert rel_err(arrival[:-1, 15:].cpu().numpy(), d['arv_s_embed']) < TOL


def test_forward_equals_forward_fixed_with_and_without_gradients():
    """`forward` (module.py:908-939) takes the adjacencies per call; same numbers, and no silent no-grad result in training."""
    dev = _dev()
    d, sd = load_golden(ASSOC[0])
    m, graphs, window, locs, grid = _assoc_setup(d, sd, dev)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    with torch.no_grad():
        out = m.forward(t('Slice'), t('Mask'), *graphs, *window)
    for a, key in zip(out, ('y', 'x', 'arv_p', 'arv_s')):
        assert rel_err(a.cpu().numpy(), d[key]) < TOL, key
    # with gradients enabled the differentiable path runs instead (test_training_steps_match_oracle)
    m.train()
    out = m.forward(t('Slice'), t('Mask'), *graphs, *window)
    assert out[0].requires_grad and out[2].requires_grad
    for a, key in zip(out, ('y', 'x', 'arv_p', 'arv_s')):
        assert rel_err(a.detach().cpu().numpy(), d[key]) < TOL, key


@pytest.mark.parametrize('explicit', [False, True])
def test_association_matches_oracle_seeded(explicit):
    """Seeded 60 x 400 network (split tcgen05 front end feeding the association kernels) and the same network as an EXPLICIT
    product graph (generic kernels, prod_grid look-ups), against the oracle; ragged picks: stations without picks, a pick
    whose time bin is the last of the table, picks outside every 2 eps window."""
    from genie_b200 import ops, synth
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    from oracle import genie_oracle as go
    dev = _dev()
    S, G, n_src, Q = 60, 400, 4, 50
    net, A_sta, A_src, Slice, Mask, attr = _random_case(S, G, 8, 15, 21, dev)
    rng = np.random.default_rng(77)
    sd = load_golden(ASSOC[0])[1]                                        # the trained Ferndale weights the fixtures carry
    A_sta, A_src, A_ps, A_pg, A_sip = go.build_adjacencies_dense(net.sta, net.grid, 8, 15)[:5]
    P = S * G
    trv = torch.from_numpy(net.travel_times()).float()                  # [G, S, 2]
    tlatent = trv.reshape(-1, 2)
    max_t = float(tlatent.max())
    eps, k_inf = 15.0, 10
    dt_partition = torch.arange(-6.0, max_t + 6.0 + 0.6, 0.6)
    l_dt = dt_partition.numel()
    A_edges_p = torch.from_numpy(rng.integers(0, P, S * l_dt * k_inf)).long()
    A_edges_s = torch.from_numpy(rng.integers(0, P, S * l_dt * k_inf)).long()
    n_arv = 70
    tpick = torch.from_numpy(rng.uniform(-5.9, max_t + 5.9, n_arv)).float()
    tpick[0] = float(dt_partition[-1]) + 0.1                             # last bin of the table
    tpick[1] = -5.999
    ipick = torch.from_numpy(rng.integers(0, S // 2, n_arv)).long()      # half of the stations have no picks
    phase = torch.from_numpy(rng.integers(0, 2, n_arv)).long().reshape(-1, 1)
    grid_cart = torch.from_numpy(net.grid).float()
    x_query = torch.from_numpy(np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q),
                                         rng.uniform(-40000.0, 0.0, Q)), axis=1)).float()
    isrc = rng.choice(G, n_src, replace=False)
    x_query_src = grid_cart[isrc]
    t_query = torch.arange(-3.0, 3.75, 0.75).reshape(-1, 1)
    tq_sample = torch.tensor([0.0, 10.0, 29.0, 200.0])
    trv_out_q = trv[isrc]
    want = go.forward_fixed(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid_cart, A_edges_p, A_edges_s, dt_partition,
                            tlatent, tpick, ipick, phase, x_query, x_query_src, t_query, tq_sample, trv_out_q, 30000.0, 9.0,
                            eps, return_parts=True)
    parts = want[4]
    ym = want[0][:, :, 0].max(1)[0]
    assert 0.05 < float(parts['mask_out'].mean()) < 0.95 and float((ym - 0.01).abs().min()) > 1e-3 * float(want[0].abs().max())
    m = _model(sd, dev, 30000.0, 9.0)
    from oracle.refshim.torch_geometric.data import Data
    mv = lambda a: a.to(dev)
    if explicit:
        m._plan = GraphPlan.from_edge_lists(mv(A_ps[:, torch.randperm(A_ps.shape[1])]), mv(A_pg), mv(A_sip), mv(A_src), 1, G,
                                            device=dev)
        # from_edge_lists recognises the Cartesian pattern only for the reference's edge order; the shuffled list is EXPLICIT
        assert m._plan.mode == 1
        m._read_in_attr = mv(attr)
        m.A_Lg_in_src, m.A_edges_p, m.A_edges_s = Data(x=mv(attr), edge_index=mv(A_sip.flip(0).contiguous())), mv(A_edges_p), mv(A_edges_s)
        m.dt_partition, m.tlatent = mv(dt_partition), mv(tlatent)
    else:
        m.set_adjacencies(mv(A_ps), mv(A_pg), Data(x=mv(attr), edge_index=mv(A_sip)),
                          Data(x=mv(attr), edge_index=mv(A_sip.flip(0).contiguous())), None, mv(A_src), mv(A_edges_p),
                          mv(A_edges_s), mv(dt_partition), mv(tlatent), mv(torch.from_numpy(net.sta).float()), mv(grid_cart))
        assert m._plan.mode == 0
    y, x, arv_p, arv_s = m.forward_fixed(mv(Slice), mv(Mask), mv(tpick), mv(ipick), mv(phase),
                                         mv(torch.from_numpy(net.sta).float()), mv(grid_cart), mv(x_query), mv(x_query_src),
                                         mv(t_query), mv(tq_sample), mv(trv_out_q))
    assert rel_err(y.cpu().numpy(), want[0].numpy()) < TOL and rel_err(x.cpu().numpy(), want[1].numpy()) < TOL
    assert rel_err(arv_p.cpu().numpy(), want[2].numpy()) < TOL
    assert rel_err(arv_s.cpu().numpy(), want[3].numpy()) < TOL


def test_association_station_passes_with_halo_equal_generic_kernels():
    """The association phase on the station-pass kernels (ASSOC instances of da_layer1_s_kernel / da_layer2_s_kernel) at a
    station count with several tiles per grid node — halo rows, ragged last tile — against the generic association kernels
    (a plan without tiling tables: L2 gathers, FFMA) on the same inputs: 300 stations x 500 grid nodes, random activations,
    about half of the source mask set, the trained Ferndale weights of the fixtures."""
    from genie_b200 import ops, synth
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    dev = _dev()
    S, G, T = 300, 500, 9
    net = synth.Network(S, G, seed=21)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    d0, sd = load_golden(ASSOC[0])
    m = _model(sd, dev, float(d0['scale_rel']), float(d0['scale_t']))
    P = S * G
    g = torch.Generator(device=dev).manual_seed(5)
    x_spatial = torch.randn((G, 30), device=dev, generator=g)
    y = torch.rand((G, T), device=dev, generator=g) * 0.0115                     # mask_out = (max_t y > 0.01): about half set
    attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
    x_latent = torch.randn((P, 30), device=dev, generator=g)
    Mask = (torch.rand((P, 4), device=dev, generator=g) < 0.3).float()
    packed = ops.AssocWeights(dev).update(m)
    outs = []
    for tiling in (True, False):
        plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev, tiling=tiling)
        assert (plan.tiles is not None) == tiling and (not tiling or plan.tiles['n_tiles'] >= 3)
        s_rows, s0, mask_out = ops.assoc_product_fwd(plan, packed, x_spatial, y, attr, x_latent, Mask, want_parts=True)
        outs.append((s_rows.clone(), s0.clone(), mask_out.clone()))
    (s_t, s0_t, mo_t), (s_g, s0_g, mo_g) = outs
    assert 0.2 < float(mo_t.mean()) < 0.8 and torch.equal(mo_t, mo_g)
    assert rel_err(s0_t.cpu().numpy(), s0_g.cpu().numpy()) < 1e-6
    assert not s_t[:, 15].any() and not s_t[:, 31].any()
    assert rel_err(s_t.cpu().numpy(), s_g.cpu().numpy()) < 2e-5


# ---- use_absolute_pos: True (module.py:913-914) --------------------------------------------------------------------------

def _abspos_model(sd, dev, d):
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, scale_rel=float(d['scale_rel']), use_absolute_pos=True, device=dev)
    m.load_state_dict(sd)
    m.TemporalAttention.scale_t = float(d['scale_t'])
    return m.eval()


@pytest.mark.parametrize('name', ['c1_10x100_abspos', 'mid_36of40x300_abspos'])
@pytest.mark.parametrize('explicit', [False, True])
def test_absolute_pos_matches_reference(name, explicit):
    """The six position channels of init_trns as per-station / per-grid-node additive terms (genie_plan_set_init_terms) on a
    CARTESIAN plan, and as per-product-node terms on an EXPLICIT plan (shuffled edge list), against the unmodified reference."""
    from genie_b200 import capi
    from oracle.refshim.torch_geometric.data import Data
    dev = _dev()
    d, sd = load_golden(name)
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    m = _abspos_model(sd, dev, d)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev)
    grid = torch.from_numpy(d['grid']).float().to(dev)
    if explicit:
        A_ps = A_ps[:, torch.randperm(A_ps.shape[1], generator=torch.Generator().manual_seed(1))]
    m.set_adjacencies(A_ps.to(dev), A_pg.to(dev), Data(x=t('read_in_attr'), edge_index=A_sip.to(dev)), None, A_sis.to(dev),
                      A_src.to(dev), None, None, None, None, locs, grid)
    assert m._plan.mode == (capi.GRAPH_EXPLICIT if explicit else capi.GRAPH_CARTESIAN)
    xs, lat, _ = m.front_end(t('Slice'), t('Mask'), grid, want_latent=True, locs_use_cart=locs)
    assert rel_err(lat.cpu().numpy(), d['x_latent']) < TOL
    assert rel_err(xs.cpu().numpy(), d['x_spatial']) < TOL
    y, x = m.forward_fixed_source(t('Slice'), t('Mask'), None, None, None, locs, grid,
                                  torch.from_numpy(d['x_query']).float().to(dev),
                                  torch.from_numpy(d['t_query']).float().reshape(-1, 1).to(dev))
    assert rel_err(y.cpu().numpy(), d['y']) < TOL and rel_err(x.cpu().numpy(), d['x']) < TOL
    # a changed scale_rel / weight must refresh the tables (no stale cache)
    m.scale_rel = 2.0 * float(d['scale_rel'])
    lat2 = m.front_end(t('Slice'), t('Mask'), grid, want_latent=True, locs_use_cart=locs)[1]
    assert rel_err(lat2.cpu().numpy(), d['x_latent']) > 1e-3


# ---- device kNN --------------------------------------------------------------------------------------

@pytest.mark.parametrize('n_x,n_y,k', [(3000, 3000, 16), (500, 4100, 10), (9, 40, 9), (20000, 257, 8), (70, 70, 32), (50000, 2, 10),
                                        (300, 2049, 12)])
def test_device_knn_matches_kdtree(n_x, n_y, k):
    """genie_knn_fwd against an fp64 k-d tree on the same fp32 points: identical indices, nearest first.  Up to 2048 queries
    take the CTA-per-query kernel, more the thread-per-query kernel."""
    from scipy.spatial import cKDTree
    from genie_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(n_x + k)
    x = (rng.uniform(0, 300.0, (n_x, 3)) * np.array([1.0, 1.0, 0.15])).astype(np.float32)
    y = x if n_x == n_y else (rng.uniform(0, 300.0, (n_y, 3)) * np.array([1.0, 1.0, 0.15])).astype(np.float32)
    got = ops.knn(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), k).cpu().numpy()
    want = cKDTree(x.astype(np.float64)).query(y.astype(np.float64), k=k)[1].reshape(n_y, k)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_device_knn_graphs_and_query_edges_match_host():
    """The three call sites: station / source graphs (process_utils.py:718-719) and SpatialAttention's query edges
    (module.py:282), against the host k-d tree builder / the oracle's knn; empty query set; k clipped to n_x."""
    from genie_b200 import ops, synth
    from genie_b200.module import knn_query_edges
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    from oracle import genie_oracle as go
    dev = _dev()
    net = synth.Network(120, 2500, seed=11)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 8, 15)
    D_sta, D_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 8, 15, device=dev)
    assert D_sta.is_cuda and torch.equal(D_sta.cpu(), A_sta) and torch.equal(D_src.cpu(), A_src)
    tiny = synth.Network(6, 40, seed=5)                                  # k_sta clipped to S - 2 (process_utils.py:712)
    T_sta, _ = extract_inputs_adjacencies_cartesian(tiny.sta, tiny.grid, 8, 15, device=dev)
    assert torch.equal(T_sta.cpu(), extract_inputs_adjacencies_cartesian(tiny.sta, tiny.grid, 8, 15)[0])
    rng = np.random.default_rng(2)
    xq = np.stack((rng.uniform(0, net.width, 333), rng.uniform(0, net.width, 333), rng.uniform(-40000, 0, 333)), 1)
    grid_t, xq_t = torch.from_numpy(net.grid).float(), torch.from_numpy(xq).float()
    e = knn_query_edges(grid_t.to(dev), xq_t.to(dev), 10)
    assert torch.equal(e.cpu(), go.knn(grid_t / 1000.0, xq_t / 1000.0, 10).flip(0))
    assert ops.knn(grid_t.to(dev), torch.zeros((0, 3), device=dev), 10).shape == (0, 10)
    assert ops.knn(grid_t[:4].to(dev), xq_t[:5].to(dev), 10).shape == (5, 4)


# ---- streaming loop with device-side stacking -------------------------------------------------------

@pytest.mark.parametrize('step_size', ['half', 'full'])
def test_day_processor_matches_oracle_loop(step_size):
    """DayProcessor (picks, travel times and Out_2 resident on the device) against the restated loop of
    process_continuous_days.py:757-813 on the 10 x 100 fixture network with the trained weights: overlapping windows, windows
    at the edges of the solution grid, windows without picks (skipped)."""
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import InputExtractor
    from genie_b200.streaming import DayProcessor, nearest_index
    from oracle import genie_oracle as go
    dev = _dev()
    d, sd = load_golden('assoc_10x100')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    step = 3.0 if step_size == 'half' else 6.0
    tsteps = np.concatenate((np.arange(0.0, 45.0, step), [150.0, 153.0, 5000.0, 5003.0]))
    tsteps_abs = np.arange(-3.0, 160.0 + 0.75, 0.75)
    xq = torch.from_numpy(d['x_query']).float()
    grid = torch.from_numpy(d['grid']).float()
    attr = torch.from_numpy(d['read_in_attr'])
    want, n_done = go.continuous_day_stack(
        sd, d['picks'], tsteps, tsteps_abs, d['ind_use'], d['sta'].shape[0], A_sis.numpy(), d['trv_times'], float(d['max_t']),
        float(d['kernel_sig_t']), float(d['dt']), A_ps, A_pg, attr, A_sip, A_src, grid, xq, float(d['scale_rel']),
        float(d['scale_t']), step_size=step_size)
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']))
    m.set_adjacencies_cartesian(A_sta, A_src, attr.to(dev), S, G, device=dev)
    ex = InputExtractor(m._plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']), float(d['kernel_sig_t']),
                        float(d['dt']))
    ex.set_day(d['picks'])
    dp = DayProcessor(m, ex, torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev), grid.to(dev), xq.to(dev),
                      step_size=step_size)
    out = dp.run(tsteps, tsteps_abs)
    assert dp.windows_done == n_done and dp.windows_skipped == len(tsteps) - n_done and dp.windows_skipped >= 2
    assert rel_err(out.cpu().numpy(), want) < TOL
    sp = DayProcessor.sparse(out, 0.01).cpu().numpy()
    iz1, iz2 = np.where(out.cpu().numpy() > 0.01)
    assert np.array_equal(sp[:, 0], iz1) and np.array_equal(sp[:, 1], iz2)
    # the 1-D nearest look-up against the k-d tree it replaces
    from scipy.spatial import cKDTree
    t = np.random.default_rng(0).uniform(-10.0, 170.0, 500)
    assert np.array_equal(nearest_index(tsteps_abs, t), cKDTree(tsteps_abs.reshape(-1, 1)).query(t.reshape(-1, 1))[1])


@pytest.mark.parametrize('step_size', ['half', 'full'])
def test_day_processor_matches_reference_loop(step_size):
    """DayProcessor against Out_2 of the REFERENCE's own loop body (process_continuous_days.py:757-813, executed verbatim by
    oracle/gen_golden.py `streaming` over the unmodified reference): overlapping windows, skipped windows, both step sizes."""
    from genie_b200.process_utils import InputExtractor
    from genie_b200.streaming import DayProcessor
    dev = _dev()
    d, sd = load_golden('streaming_10x100')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']))
    m.set_adjacencies_cartesian(torch.from_numpy(d['A_sta_sta']), torch.from_numpy(d['A_src_src']),
                                torch.from_numpy(d['read_in_attr']).to(dev), S, G, device=dev)
    ex = InputExtractor(m._plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']), float(d['kernel_sig_t']),
                        float(d['dt']))
    ex.set_day(d['picks'])
    dp = DayProcessor(m, ex, torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev), torch.from_numpy(d['grid']).float().to(dev),
                      torch.from_numpy(d['x_query']).float().to(dev), t_win=float(d['t_win']), dt_win=float(d['dt_win']),
                      step_size=step_size)
    out = dp.run(d['tsteps_' + step_size], d['tsteps_abs_' + step_size]).cpu().numpy()
    want = d['Out_2_' + step_size]
    assert dp.windows_skipped > 0 and dp.windows_done > 0
    assert rel_err(out, want) < TOL
    assert np.array_equal(np.abs(out).sum(axis=0) > 0, np.abs(want).sum(axis=0) > 0)


def test_day_processor_fused_runner_equals_two_step_loop():
    """The streaming loop on the fused window path (WindowRunner: genie_window_fwd + heads as one CUDA graph per window) and on
    the reference-shaped two-step path produce the same Out_2, bit for bit (100 stations: a plan with tiling tables)."""
    from genie_b200 import synth
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.process_utils import InputExtractor, extract_inputs_adjacencies_cartesian
    from genie_b200.streaming import DayProcessor
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 100, 800
    net = synth.Network(S, G, seed=5)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    m = GCN_Detection_Network_extended(None, None, scale_rel=30000.0, device=dev).eval()
    m.load_state_dict(go.init_state(seed=4), strict=False)
    m.set_adjacencies_cartesian(A_sta, A_src, torch.from_numpy(net.read_in_offsets(30000.0)).to(dev), S, G, device=dev)
    ex = InputExtractor(m._plan, net.travel_times(), np.arange(S), S, net.max_moveout(), 3.0, 0.3)
    ex.set_day(synth.make_picks(net, 0.0, 300.0, seed=6, false_per_sta_min=3.0))
    xq = torch.from_numpy(np.random.default_rng(2).uniform(0, net.width, (200, 3))).float().to(dev)
    xq[:, 2] = -xq[:, 2] / net.width * 40000.0
    args = (m, ex, torch.from_numpy(net.sta).float().to(dev), torch.from_numpy(net.grid).float().to(dev), xq)
    tsteps, tsteps_abs = np.arange(0.0, 150.0, 3.0), np.arange(-3.0, 303.0 + 0.75, 0.75)
    a = DayProcessor(*args, use_runner=True).run(tsteps, tsteps_abs)
    b = DayProcessor(*args, use_runner=False).run(tsteps, tsteps_abs)
    assert torch.equal(a, b) and float(a.abs().max()) > 0


# ---- training path (BASELINE.json configs[2]): forward with gradients, loss and Adam steps against the oracle ------------------

@pytest.mark.parametrize('mode,C', [(0, 30), (1, 30), (1, 15), (2, 34), (0, 1)])
def test_kron_spmm_forward_and_transpose(mode, C):
    """genie_kron_spmm_fwd on the forward (by target) and the transposed (by source) CSR of one edge type against a dense
    product; together they are the mean aggregation of `propagate` and its gradient."""
    from genie_b200 import ops
    from genie_b200.plan import csr_by_destination
    from genie_b200.training import KronGraph
    from oracle import genie_oracle as go
    dev = _dev()
    S, G = 23, 41
    rng = np.random.default_rng(mode * 10 + C)
    if mode == 2:
        n = 777
        E = 5 * n
        A = torch.from_numpy(np.stack((rng.integers(0, n, E), rng.integers(0, n - 30, E)), 0)).long()      # 30 isolated targets
        rowptr, col = csr_by_destination(A, n)
        kg = KronGraph(2, 0, 0, n, rowptr, col, dev)
        P = n
    else:
        pts = rng.uniform(0, 100.0, ((S if mode == 0 else G), 3)).astype(np.float32)
        A = go.knn_graph_no_self(pts, 6)
        n = S if mode == 0 else G
        rowptr, col = csr_by_destination(A, n)
        kg = KronGraph(mode, S, G, S * G, rowptr, col, dev)
        P = S * G
    x = torch.from_numpy(rng.normal(size=(P, C)).astype(np.float32))
    dense = torch.zeros((n, n), dtype=torch.float64)
    deg = torch.bincount(A[1], minlength=n).clamp(min=1).double()
    dense.index_put_((A[1], A[0]), 1.0 / deg[A[1]], accumulate=True)

    def apply(M):
        if mode == 0:
            return torch.einsum('st,gtc->gsc', M, x.double().view(G, S, C)).reshape(P, C)
        if mode == 1:
            return torch.einsum('gh,hsc->gsc', M, x.double().view(G, S, C)).reshape(P, C)
        return M @ x.double()
    got_f = ops.kron_spmm(kg, kg.fwd, x.to(dev)).cpu().double()
    got_r = ops.kron_spmm(kg, kg.rev, x.to(dev)).cpu().double()
    assert rel_err(got_f.numpy(), apply(dense).numpy()) < 1e-6
    assert rel_err(got_r.numpy(), apply(dense.t()).numpy()) < 1e-6
    # the forward mean equals the oracle's propagate_mean over the explicit product edge list
    if mode == 0:
        A_prod = (A.repeat(1, G) + S * torch.arange(G).repeat_interleave(A.shape[1]).view(1, -1))
        want = go.propagate_mean(x.index_select(0, A_prod[0]), A_prod[1], P)
        assert rel_err(got_f.numpy(), want.numpy()) < 1e-6


@pytest.mark.parametrize('name', ['assoc_10x100', 'assoc_14of16x120_edges', 'assoc_14of16x120_abspos'])
def test_training_steps_match_oracle(name):
    """`mz(*input_tensors)` with gradients (train_GENIE_model.py:1786) + Adam (lr 1e-3) for three steps on the device against
    the oracle's autograd on the CPU: outputs, the weighted MSE loss of :1384, :1789 on synthetic labels, every parameter gradient
    of the first step, and the loss trajectory."""
    from genie_b200 import capi
    from test_oracle_golden import assoc_inputs, assoc_variant
    from oracle import genie_oracle as go
    dev = _dev()
    d, sd = load_golden(name)
    m, graphs, window, locs, grid = _assoc_setup(d, sd, dev, name)
    m.train()
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    kw = assoc_inputs(d)
    pos_rel, abs_pos = assoc_variant(d, name, A_ps, A_pg, A_sis)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    rng = np.random.default_rng(5)
    lbl = [torch.from_numpy(rng.uniform(0, 1, d[k].shape[:2]).astype(np.float32)) for k in ('y', 'x', 'arv_p', 'arv_s')]
    wts = (0.1, 0.4, 0.25, 0.25)
    loss_fn = torch.nn.MSELoss()                     # train_GENIE_model.py:1384

    def loss_of(out, to=lambda a: a):
        return sum(w * loss_fn(o[:, :, 0], to(l)) for w, o, l in zip(wts, out, lbl))
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt_o = torch.optim.Adam(list(sdo.values()), lr=1e-3)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    n0 = capi.launch_count()
    for step in range(3):
        opt_o.zero_grad()
        want = go.forward_fixed(sdo, torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask']), A_ps, A_pg,
                                torch.from_numpy(d['read_in_attr']), A_sip, A_src, torch.from_numpy(d['grid']).float(),
                                scale_rel=float(d['scale_rel']), scale_t=float(d['scale_t']), eps=float(d['eps']),
                                pos_rel=pos_rel, abs_pos=abs_pos, **kw)
        loss_o = loss_of(want)
        loss_o.backward()
        opt.zero_grad()
        out = m(t('Slice'), t('Mask'), *graphs, *window)
        loss = loss_of(out, lambda a: a.to(dev))
        loss.backward()
        assert abs(float(loss.detach()) - float(loss_o.detach())) < 1e-4 * abs(float(loss_o.detach())), step
        if step == 0:
            for a, b, key in zip(out, want, ('y', 'x', 'arv_p', 'arv_s')):
                assert rel_err(a.detach().cpu().numpy(), b.detach().numpy()) < TOL, key
            checked = 0
            for k, p in m.named_parameters():
                g_o = sdo[k].grad
                if g_o is None or p.grad is None:
                    assert (g_o is None or not g_o.any()) and (p.grad is None or not p.grad.any()), k
                    continue
                assert rel_err(p.grad.cpu().numpy(), g_o.numpy()) < 1e-3, k
                checked += 1
            assert checked > 100
        opt_o.step()
        opt.step()
    assert capi.launch_count() - n0 >= 3 * 16           # 8 aggregations forward + 8 backward per step ran on our kernel


@pytest.mark.parametrize('widths,n_out,act,bias', [((4, 4), 30, True, True), ((30, 30, 4), 30, True, True), ((60,), 30, True, True),
                                                   ((60, 30, 4), 15, True, True), ((15, 30, 1, 4), 30, True, True),
                                                   ((30, 3), 30, True, True), ((30,), 15, False, True), ((60, 30, 8), 15, True, False),
                                                   ((104,), 32, True, True)])
def test_node_mlp_kernels_match_torch(widths, n_out, act, bias):
    """genie_node_mlp_fwd / genie_node_mlp_bwd — `activate(Linear(cat(parts)))` of the training path, forward and backward in
    one kernel each — against the torch ops they replace (cat + linear + prelu and autograd): outputs, the gradient of every
    part, of the weight, the bias and the PReLU slope.  Ragged row count, strided parts."""
    from genie_b200.training import NodeMLP
    dev = _dev()
    n = 5000 + 37
    g = torch.Generator(device='cpu').manual_seed(sum(widths) * 7 + n_out)
    parts_cpu = [torch.randn((n, w + 3), generator=g)[:, 1:1 + w] for w in widths]                # column-sliced: row stride w + 3
    W = torch.randn((n_out, sum(widths)), generator=g) * 0.3
    b = torch.randn(n_out, generator=g) if bias else None
    a = torch.tensor([0.2 if n_out != 15 else -0.3]) if act else None      # a PReLU slope may be negative (trained weights are)
    gy = torch.randn((n, n_out), generator=g)

    def run(device, fused):
        ps = [p.to(device).clone().requires_grad_(True) for p in parts_cpu]
        Wd = W.to(device).clone().requires_grad_(True)
        bd = b.to(device).clone().requires_grad_(True) if b is not None else None
        ad = a.to(device).clone().requires_grad_(True) if a is not None else None
        if fused:
            y = NodeMLP.apply(Wd, bd, ad, *[p[:, :] for p in ps])
        else:
            y = torch.nn.functional.linear(torch.cat(ps, dim=1).double(), Wd.double(), bd.double() if bd is not None else None)
            if ad is not None:
                y = torch.where(y >= 0, y, ad.double() * y)
        y.backward(gy.to(device).to(y.dtype))
        return [y] + [p.grad for p in ps] + [Wd.grad] + ([bd.grad] if bd is not None else []) + ([ad.grad] if ad is not None else [])
    got = run(dev, True)
    want = run('cpu', False)                    # fp64 on the host
    for i, (x, w_) in enumerate(zip(got, want)):
        assert rel_err(x.detach().cpu().numpy(), w_.detach().numpy()) < 2e-5, i


def test_training_batch32_trajectory_matches_oracle(monkeypatch):
    """BASELINE.json configs[2]: batch 32 windows, Adam, loss match.  The step of train_GENIE_model.py:1593, 1786-1789, 1843-1861:
    optimizer.zero_grad(); for each of the 32 samples loss = weighted MSE / n_batch, loss.backward() (gradients accumulate);
    optimizer.step() — five steps on the device with the fused per-node layer kernels (genie_node_mlp_*) and the product-graph
    gather kernel, against the oracle's autograd on the CPU: the summed loss of every step."""
    import genie_b200.training as training
    from genie_b200 import capi
    from test_oracle_golden import assoc_inputs, assoc_variant
    from oracle import genie_oracle as go
    monkeypatch.setattr(training, 'MLP_MIN_ROWS', 0)             # the fixture network has 2880 product nodes: use the kernels anyway
    monkeypatch.setattr(training, 'LIN_SPLIT_MIN_ROWS', 0)
    dev = _dev()
    name = 'assoc_18of20x160'
    d, sd = load_golden(name)
    m, graphs, window, locs, grid = _assoc_setup(d, sd, dev, name)
    m.train()
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    kw = assoc_inputs(d)
    pos_rel, abs_pos = assoc_variant(d, name, A_ps, A_pg, A_sis)
    n_batch, n_steps = 32, 5
    rng = np.random.default_rng(11)
    Sl0, Mk0 = d['Slice'], d['Mask']
    samples = []
    for i in range(n_batch):
        keep = rng.random(Sl0.shape) < 0.8                                       # every sample: its own thinned inputs and labels
        Sl = (Sl0 * keep * rng.uniform(0.5, 1.0, Sl0.shape)).astype(np.float32)
        Mk = (np.abs(Sl) > 0.01).astype(np.float32)
        lbl = [rng.uniform(0, 1, d[k].shape[:2]).astype(np.float32) for k in ('y', 'x', 'arv_p', 'arv_s')]
        samples.append((Sl, Mk, lbl))
    wts = (0.1, 0.4, 0.25, 0.25)
    loss_fn = torch.nn.MSELoss()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt_o = torch.optim.Adam(list(sdo.values()), lr=1e-3)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    n0 = capi.launch_count()
    for step in range(n_steps):
        opt_o.zero_grad()
        opt.zero_grad()
        tot_o = tot = 0.0
        for Sl, Mk, lbl in samples:
            want = go.forward_fixed(sdo, torch.from_numpy(Sl), torch.from_numpy(Mk), A_ps, A_pg, torch.from_numpy(d['read_in_attr']),
                                    A_sip, A_src, torch.from_numpy(d['grid']).float(), scale_rel=float(d['scale_rel']),
                                    scale_t=float(d['scale_t']), eps=float(d['eps']), pos_rel=pos_rel, abs_pos=abs_pos, **kw)
            lo = sum(w * loss_fn(o[:, :, 0], torch.from_numpy(l)) for w, o, l in zip(wts, want, lbl)) / n_batch
            lo.backward()
            tot_o += float(lo.detach())
            out = m(torch.from_numpy(Sl).to(dev), torch.from_numpy(Mk).to(dev), *graphs, *window)
            ls = sum(w * loss_fn(o[:, :, 0], torch.from_numpy(l).to(dev)) for w, o, l in zip(wts, out, lbl)) / n_batch
            ls.backward()
            tot += float(ls.detach())
        assert abs(tot - tot_o) < 1e-4 * abs(tot_o), (step, tot, tot_o)
        if step == 0:
            # every accumulated parameter gradient: 1e-3 of its own scale (parameters whose gradient is below 1e-4 of the
            # largest one are held to that floor — their own scale is rounding noise of the 32-sample sum)
            top = max(float(v.grad.abs().max()) for v in sdo.values() if v.grad is not None)
            bad = []
            for k, p in m.named_parameters():
                g_o = sdo[k].grad
                if g_o is not None and p.grad is not None and g_o.any():
                    err = float((p.grad.cpu() - g_o).abs().max())
                    scale = max(float(g_o.abs().max()), 1e-4 * top)
                    if err > 1e-3 * scale:
                        bad.append((k, err, float(g_o.abs().max())))
            assert not bad, (top, bad[:8])
        opt_o.step()
        opt.step()
    assert capi.launch_count() - n0 >= n_steps * n_batch * 40       # 16 gather + 2 x 16 fused layer launches per sample at least


# ---- sub-graph mode (process_utils.py:744-849; SURVEY.md §8d "C4 subgraph variant") -------------------------------------------

@pytest.mark.parametrize('name', ['subgraph_14x60', 'subgraph_30x200', 'subgraph_12x40_ragged'])
def test_subgraph_mode_matches_reference(name):
    """The sub-graph builder on the device (kNN through genie_knn_fwd), then inputs (a1 over the pair list) and
    forward_fixed_source on the EXPLICIT plan, against the unmodified reference."""
    from genie_b200 import capi
    from genie_b200.process_utils import extract_inputs_adjacencies_subgraph, InputExtractor
    from oracle.refshim.torch_geometric.data import Data
    dev = _dev()
    d, sd = load_golden(name)
    out = extract_inputs_adjacencies_subgraph(d['sta'], d['grid'], lambda x: x, None, max_deg_offset=float(d['max_deg_offset']),
                                              k_nearest_pairs=int(d['k_nearest_pairs']), k_sta_edges=int(d['k_sta']),
                                              k_spc_edges=int(d['k_spc']), device=dev)
    for got, key in zip(out, ('A_sta_sta', 'A_src_src', 'A_prod_sta_sta', 'A_prod_src_src', 'A_src_in_prod', 'A_src_in_sta')):
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), d[key]), key
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = out
    m = _model(sd, dev, float(d['scale_rel']), float(d['scale_t']))
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    locs, grid = t('sta').float(), t('grid').float()
    m.set_adjacencies(A_ps, A_pg, Data(x=t('read_in_attr'), edge_index=A_sip), None, A_sis, A_src, None, None, None, None,
                      locs, grid)
    assert m._plan.mode == capi.GRAPH_EXPLICIT and m._plan.n_prod == d['Slice'].shape[0]
    ex = InputExtractor(m._plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']), float(d['kernel_sig_t']),
                        float(d['dt']), node_sta=A_sis[0], node_grid=A_sis[1])
    ex.set_day(d['picks'])
    Slice, Mask = ex(float(d['t0']))
    assert np.array_equal(Slice.cpu().numpy(), d['Slice']) and np.array_equal(Mask.cpu().numpy(), d['Mask'])
    xs, lat, _ = m.front_end(Slice, Mask, grid, want_latent=True)
    assert rel_err(lat.cpu().numpy(), d['x_latent']) < TOL and rel_err(xs.cpu().numpy(), d['x_spatial']) < TOL
    y, x = m.forward_fixed_source(Slice, Mask, None, None, None, locs, grid, t('x_query').float(),
                                  t('t_query').float().reshape(-1, 1))
    assert rel_err(y.cpu().numpy(), d['y']) < TOL and rel_err(x.cpu().numpy(), d['x']) < TOL


def test_setup_builders_on_the_device_match_reference():
    """extract_inputs_adjacencies and compute_time_embedding_vectors with device= (kNN through genie_knn_fwd, index arithmetic
    in torch on the GPU) against the fixtures of the unmodified reference."""
    from genie_b200.process_utils import extract_inputs_adjacencies, compute_time_embedding_vectors
    dev = _dev()
    d, _ = load_golden('dense_adjacencies_9of12x30')
    out = extract_inputs_adjacencies(None, d['sta'], d['ind_use'], d['grid'], None, d['ref_t'], d['ptr_p'], d['ptr_s'],
                                     lambda x: x, [int(d['k_sta']), int(d['k_spc']), int(d['k_time'])], device=dev)
    for got, key in zip(out, ('A_sta_sta', 'A_src_src', 'A_prod_sta_sta', 'A_prod_src_src', 'A_src_in_prod', 'A_edges_time_p',
                              'A_edges_time_s', 'A_edges_ref')):
        if torch.is_tensor(got):
            assert got.is_cuda
            got = got.cpu().numpy()
        assert np.array_equal(np.asarray(got), d[key]), key
    d, _ = load_golden('assoc_18of20x160')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sis = np.stack((np.tile(np.arange(S), G), np.repeat(np.arange(G), S)))
    sig = float(d['kernel_sig_t'])
    ep, es, dtp = compute_time_embedding_vectors(None, d['sta'][d['ind_use']], d['grid'], A_sis, float(d['max_t']),
                                                 dt_res=sig / 5.0, t_win=sig * 2.0, trv_out=d['tlatent'], device=dev)
    assert np.array_equal(ep, d['A_edges_p']) and np.array_equal(es, d['A_edges_s']) and np.array_equal(dtp, d['dt_partition'])


def test_forward_fixed_edge_cases():
    """Association branch with no picks at all, and with a source mask that is zero everywhere (y below the threshold):
    shapes as the reference's, values against the oracle."""
    from test_oracle_golden import assoc_inputs
    from oracle import genie_oracle as go
    dev = _dev()
    d, sd = load_golden('assoc_10x100')
    sd = {k: v.clone() for k, v in sd.items()}
    m, graphs, window, locs, grid = _assoc_setup(d, sd, dev)
    m.set_adjacencies(*graphs, locs, grid)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    # no picks: arv_p / arv_s are [n_src, 0, 1]
    w = list(window)
    w[0], w[1], w[2] = torch.zeros(0, device=dev), torch.zeros(0, dtype=torch.long, device=dev), torch.zeros((0, 1), dtype=torch.long, device=dev)
    y, x, arv_p, arv_s = m.forward_fixed(t('Slice'), t('Mask'), *w)
    n_src = len(d['tq_sample'])
    assert arv_p.shape == (n_src, 0, 1) and arv_s.shape == (n_src, 0, 1)
    assert rel_err(y.cpu().numpy(), d['y']) < TOL
    # source mask zero everywhere: shift the last bias so that max y < 0.01 on every grid node
    sd['TemporalAttention.proj_2.bias'] = sd['TemporalAttention.proj_2.bias'] - 5.0
    m.load_state_dict(sd)
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    want = go.forward_fixed(sd, torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask']), A_ps, A_pg,
                            torch.from_numpy(d['read_in_attr']), A_sip, A_src, torch.from_numpy(d['grid']).float(),
                            scale_rel=float(d['scale_rel']), scale_t=float(d['scale_t']), eps=float(d['eps']),
                            return_parts=True, **assoc_inputs(d))
    assert not want[4]['mask_out'].any()
    out = m.forward_fixed(t('Slice'), t('Mask'), *window)
    for a, b, key in zip(out, want[:4], ('y', 'x', 'arv_p', 'arv_s')):
        assert rel_err(a.cpu().numpy(), b.numpy()) < TOL, key


# ---- GraphDD: the second consumer of the kernel family (Relocation/train_double_difference_model.py:333-536) --------------------

@pytest.mark.parametrize('name', ['graphdd_12x9', 'graphdd_10x14_memory'])
def test_graphdd_location_network_matches_reference(name, monkeypatch):
    """genie_b200.relocation.GNN_Location (same classes and state_dict keys as the reference's script) with the reference's own
    weights: the four outputs against the unmodified reference's, and every parameter gradient of a sum-of-outputs loss against
    the oracle's autograd — with the per-edge layers on genie_node_mlp_* and the gathers / means on genie_kron_spmm_fwd."""
    import genie_b200.training as training
    from genie_b200 import capi
    from genie_b200.relocation import GNN_Location
    from oracle import graphdd_oracle as gd
    monkeypatch.setattr(training, 'MLP_MIN_ROWS', 0)            # the fixture graphs have a few hundred edges: use the kernels anyway
    dev = _dev()
    d, sd = load_golden(name)
    use_memory = bool(int(d['use_memory']))
    m = GNN_Location(None, None, inpt_sources=True, use_memory=use_memory, device=dev)
    missing = m.load_state_dict({k: v.to(dev) for k, v in sd.items()}, strict=True)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    mem = t('memory') if use_memory else False
    args = (t('A_in_pick'), t('A_in_src'), t('A_src_in_product'), t('A_sta_in_product'), t('A_src_in_sta'), t('locs').float(),
            t('srcs').float())
    n0 = capi.launch_count()
    out = m(t('x'), t('mask'), *args, memory=mem)
    assert capi.launch_count() - n0 > 100
    for i, o in enumerate(out):
        assert rel_err(o.detach().cpu().numpy(), d['out%d' % i]) < TOL, i
    m.set_adjacencies(*args)
    out_fixed = m.forward_fixed(t('x'), t('mask'), memory=mem)
    assert all(torch.equal(a, b) for a, b in zip(out, out_fixed))
    # gradients
    w = [torch.from_numpy(np.random.default_rng(i).normal(size=d['out%d' % i].shape).astype(np.float32)) for i in range(4)]
    sum(float(1.0) * (o * wi.to(dev)).sum() / (5000.0 if i == 0 else 1.0) for i, (o, wi) in enumerate(zip(out, w))).backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    tc = lambda k: torch.from_numpy(d[k])
    want = gd.gnn_location(sdo, tc('x'), tc('mask'), tc('A_in_pick'), tc('A_in_src'), tc('A_src_in_product'), tc('A_sta_in_product'),
                           tc('A_src_in_sta'), tc('locs').float(), tc('srcs').float(), memory=tc('memory') if use_memory else None)
    sum((o * wi).sum() / (5000.0 if i == 0 else 1.0) for i, (o, wi) in enumerate(zip(want, w))).backward()
    top = max(float(v.grad.abs().max()) for v in sdo.values() if v.grad is not None)
    checked = 0
    for k, p in m.named_parameters():
        g_o = sdo[k].grad
        if g_o is None or not g_o.any():
            continue
        err = float((p.grad.cpu() - g_o).abs().max())
        assert err <= 1e-3 * max(float(g_o.abs().max()), 1e-4 * top), (k, err, float(g_o.abs().max()))
        checked += 1
    assert checked > 120
