"""The oracle restatement against the reference's own outputs (tests/golden, made by oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import genie_oracle as go

SYNTH = ['c1_10x100', 'mid_36of40x300', 'small_6x40', 'ferndale_t38940']


def _graphs(d):
    S = len(d['ind_use'])
    G = d['grid'].shape[0]
    out = go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']), int(d['k_spc']))
    return (S, G) + tuple(out)


@pytest.mark.parametrize('name', SYNTH)
def test_adjacencies_match_reference(name):
    d, _ = load_golden(name)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    assert np.array_equal(A_sta.numpy(), d['A_sta_sta'])
    assert np.array_equal(A_src.numpy(), d['A_src_src'])
    assert A_ps.shape[1] == A_sta.shape[1] * G and A_pg.shape[1] == A_src.shape[1] * S


@pytest.mark.parametrize('name', SYNTH)
def test_input_scatter_bit_exact(name):
    d, _ = load_golden(name)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    Slice, Mask, parts = go.input_scatter(d['picks'], float(d['t0']), d['ind_use'], d['sta'].shape[0], A_sis.numpy(),
                                          d['trv_times'], float(d['max_t']), float(d['kernel_sig_t']), float(d['dt']),
                                          return_parts=True)
    assert parts['n_ts'] == int(d['n_ts']) and parts['ref0'] == float(d['ref0'])
    assert np.array_equal(Slice, d['Slice'])          # same numpy expressions -> identical bits
    assert np.array_equal(Mask, d['Mask'])
    # per-station series against the reference's return_embedding output (only stations with picks are kept there)
    perm = -np.ones(d['sta'].shape[0], dtype=int)
    perm[d['ind_use']] = np.arange(S)
    loc = perm[d['ind_unique']]
    assert np.array_equal(parts['series'][0][loc], d['embed_p'])
    assert np.array_equal(parts['series'][1][loc], d['embed_s'])
    rest = np.setdiff1d(np.arange(S), loc)
    assert not parts['series'][:, rest].any()


@pytest.mark.parametrize('name', SYNTH)
def test_front_end_and_heads_match_reference(name):
    d, sd = load_golden(name)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    Slice, Mask = torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask'])
    y, x, parts = go.forward_fixed_source(
        sd, Slice, Mask, A_ps, A_pg, torch.from_numpy(d['read_in_attr']), A_sip, A_src,
        torch.from_numpy(d['grid']).float(), torch.from_numpy(d['x_query']).float(),
        torch.from_numpy(d['t_query']).float().reshape(-1, 1), float(d['scale_rel']), float(d['scale_t']),
        return_parts=True)
    for key in ('x_latent', 'read_in', 'sa1', 'sa2', 'x_spatial', 'y_latent', 'x_query_embed'):
        assert rel_err(parts[key].numpy(), d[key]) < 2e-6, key
    assert rel_err(y.numpy(), d['y']) < 2e-6
    assert rel_err(x.numpy(), d['x']) < 2e-6


@pytest.mark.parametrize('name', ['c1_10x100_edges', 'mid_36of40x300_edges'])
def test_edge_feature_model_matches_reference(name):
    """a2': the reference's `use_updated_model_definition: True` classes (DataAggregationEdges, module.py:102-174) run
    unmodified by oracle/gen_golden.py synthetic_edges."""
    d, sd = load_golden(name)
    assert sd['DataAggregation.l1_t1_2.weight'].shape == (30, 68) and sd['DataAggregation.l2_t2_2.weight'].shape == (15, 98)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    Slice, Mask = torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask'])
    sta = torch.from_numpy(d['sta'][d['ind_use']]).float()
    grid = torch.from_numpy(d['grid']).float()
    pos_rel = go.edge_features(sta, grid, A_sis, A_ps, A_pg, float(d['scale_rel']))
    y, x, parts = go.forward_fixed_source(
        sd, Slice, Mask, A_ps, A_pg, torch.from_numpy(d['read_in_attr']), A_sip, A_src, grid,
        torch.from_numpy(d['x_query']).float(), torch.from_numpy(d['t_query']).float().reshape(-1, 1),
        float(d['scale_rel']), float(d['scale_t']), return_parts=True, pos_rel=pos_rel)
    for key in ('x_latent', 'read_in', 'sa1', 'sa2', 'x_spatial', 'y_latent', 'x_query_embed'):
        assert rel_err(parts[key].numpy(), d[key]) < 2e-6, key
    assert rel_err(y.numpy(), d['y']) < 2e-6
    assert rel_err(x.numpy(), d['x']) < 2e-6


@pytest.mark.parametrize('name', ['legacy_12of14x60', 'legacy_8x30_short'])
def test_legacy_input_features_match_reference(name):
    """a1': extract_inputs_from_data_fixed_grids_with_phase_type run unmodified (oracle/gen_golden.py legacy_input)."""
    d, _ = load_golden(name)
    P = d['picks']
    Inpts, Masks, parts = go.legacy_input_features(P, P[:, 4], d['ind_use'], d['time_samples'], d['trv_times'],
                                                    float(d['max_t']), float(d['t_win']), float(d['kernel_sig_t']),
                                                    return_parts=True)
    lists = go.legacy_pick_lists(P, P[:, 4], d['ind_use'], d['time_samples'], parts['lp'], d['sta'].shape[0])
    for i in range(len(d['time_samples'])):
        assert np.array_equal(Inpts[i], d['Inpts%d' % i])          # same numpy expressions -> identical bits
        assert np.array_equal(Masks[i], d['Masks%d' % i])
        assert np.array_equal(lists[0][i], d['lp_times%d' % i]) and np.array_equal(lists[1][i], d['lp_stations%d' % i])
        assert np.array_equal(lists[2][i], d['lp_phases%d' % i]) and np.array_equal(lists[3][i], d['lp_meta%d' % i])


def test_mean_of_empty_neighbourhood_is_zero():
    msg = torch.ones(3, 2)
    out = go.propagate_mean(msg, torch.tensor([0, 0, 2]), 4)
    assert out.tolist() == [[1.0, 1.0], [0.0, 0.0], [1.0, 1.0], [0.0, 0.0]]


def test_time_axis_bin_count_follows_fp64():
    # 3*3.5/0.35 = 30.000000000000004 in fp64 -> ceil 31 -> 63 bins
    assert np.ceil(3 * 3.5 / np.round(3.5 / 10.0, 2)) == 31
    assert np.ceil(3 * 3.0 / np.round(3.0 / 10.0, 2)) == 30


def test_ferndale_known_answers():
    """Digests of the reference run on Examples/Ferndale.zip agree with the survey's independent run."""
    d, _ = load_golden('ferndale_t38940')
    assert d['Slice'].shape == (11550, 4) and int((d['Slice'] != 0).sum()) == 23047 and int(d['Mask'].sum()) == 22701
    assert abs(float(d['Slice'].astype(np.float64).sum()) - 12107.058853) < 1e-3
    assert abs(float(d['x_latent'].astype(np.float64).sum()) - 33790.148093) < 5e-2
    assert abs(float(np.abs(d['x_latent'].astype(np.float64)).sum()) - 45289.203036) < 5e-2
    assert abs(float(d['x_latent'].max()) - 3.682671) < 1e-5
    assert abs(float(d['read_in'].astype(np.float64).sum()) - 901.758727) < 1e-2
    assert abs(float(d['x_spatial'].max()) - 7.177079) < 1e-5
    assert abs(float(d['y'].max()) - 0.946654) < 1e-5


ASSOC = ['assoc_10x100', 'assoc_18of20x160', 'assoc_14of16x120_edges', 'assoc_14of16x120_abspos']


def assoc_variant(d, name, A_ps, A_pg, A_sis):
    """(pos_rel, abs_pos) arguments of oracle.forward_fixed for the fixture's model variant."""
    sta = torch.from_numpy(d['sta'][d['ind_use']]).float()
    grid = torch.from_numpy(d['grid']).float()
    pos_rel = go.edge_features(sta, grid, A_sis, A_ps, A_pg, float(d['scale_rel'])) if name.endswith('_edges') else None
    abs_pos = (sta, A_sis) if name.endswith('_abspos') else None
    return pos_rel, abs_pos


def assoc_inputs(d):
    """Tensors of an association fixture in the order oracle.forward_fixed takes them (after the graphs)."""
    t = torch.from_numpy
    return dict(A_edges_p=t(d['A_edges_p']), A_edges_s=t(d['A_edges_s']), dt_partition=t(d['dt_partition']).float(),
                tlatent=t(d['tlatent']).float(), tpick=t(d['tpick']).float(), ipick=t(d['ipick']).long(),
                phase_label=t(d['phase_label']).long().reshape(-1, 1), x_query_cart=t(d['x_query']).float(),
                x_query_src_cart=t(d['x_query_src']).float(), t_query=t(d['t_query']).float().reshape(-1, 1),
                tq_sample=t(d['tq_sample']).float(), trv_out_q=t(d['trv_out_q']).float())


@pytest.mark.parametrize('name', ASSOC)
def test_association_branch_matches_reference(name):
    """forward_fixed (module.py:963-997) incl. BipartiteGraphReadOutOperator, DataAggregationAssociationPhase,
    LocalSliceLgCollapse{P,S} and Arrivals against the unmodified reference."""
    d, sd = load_golden(name)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    Slice, Mask = torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask'])
    kw = assoc_inputs(d)
    kw['pos_rel'], kw['abs_pos'] = assoc_variant(d, name, A_ps, A_pg, A_sis)
    y, x, arv_p, arv_s, parts = go.forward_fixed(
        sd, Slice, Mask, A_ps, A_pg, torch.from_numpy(d['read_in_attr']), A_sip, A_src,
        torch.from_numpy(d['grid']).float(), scale_rel=float(d['scale_rel']), scale_t=float(d['scale_t']),
        eps=float(d['eps']), return_parts=True, **kw)
    assert np.array_equal(parts['mask_out'].numpy()[A_sip[1].numpy()], d['mask_out_1'])
    assert 0.05 < d['mask_out_1'].mean() < 0.95          # the fixture exercises both mask values
    for key in ('x_latent', 'x_spatial', 'y_latent', 'x_src', 'assoc_s0', 'assoc_s', 'arv_p_embed', 'arv_s_embed'):
        assert rel_err(parts[key].numpy(), d[key]) < 5e-6, key
    assert rel_err(y.numpy(), d['y']) < 5e-6 and rel_err(x.numpy(), d['x']) < 5e-6
    assert rel_err(arv_p.numpy(), d['arv_p']) < 5e-6 and rel_err(arv_s.numpy(), d['arv_s']) < 5e-6
    assert arv_p.shape == (len(d['tq_sample']), len(d['tpick']), 1)


@pytest.mark.parametrize('name', ['c1_10x100_abspos', 'mid_36of40x300_abspos'])
def test_absolute_pos_matches_reference(name):
    """`use_absolute_pos: True` (module.py:913-914): init_trns takes 14 inputs."""
    d, sd = load_golden(name)
    assert tuple(sd['DataAggregation.init_trns.weight'].shape) == (30, 14)
    S, G, A_sta, A_src, A_ps, A_pg, A_sip, A_sis = _graphs(d)
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float()
    y, x, parts = go.forward_fixed_source(
        sd, torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask']), A_ps, A_pg, torch.from_numpy(d['read_in_attr']), A_sip,
        A_src, torch.from_numpy(d['grid']).float(), torch.from_numpy(d['x_query']).float(),
        torch.from_numpy(d['t_query']).float().reshape(-1, 1), float(d['scale_rel']), float(d['scale_t']),
        return_parts=True, abs_pos=(locs, A_sis))
    for key in ('x_latent', 'read_in', 'x_spatial'):
        assert rel_err(parts[key].numpy(), d[key]) < 2e-6, key
    assert rel_err(y.numpy(), d['y']) < 2e-6 and rel_err(x.numpy(), d['x']) < 2e-6


@pytest.mark.parametrize('name', ['subgraph_14x60', 'subgraph_30x200', 'subgraph_12x40_ragged'])
def test_subgraph_window_matches_reference(name):
    """Sub-graph mode: inputs (a1 with the pair list as A_src_in_sta) and the front end + heads on the explicit product graph."""
    d, sd = load_golden(name)
    t = torch.from_numpy
    Slice, Mask = go.input_scatter(d['picks'], float(d['t0']), d['ind_use'], d['sta'].shape[0], d['A_src_in_sta'],
                                   d['trv_times'], float(d['max_t']), float(d['kernel_sig_t']), float(d['dt']))
    assert np.array_equal(Slice, d['Slice']) and np.array_equal(Mask, d['Mask'])
    y, x, parts = go.forward_fixed_source(
        sd, t(d['Slice']), t(d['Mask']), t(d['A_prod_sta_sta']), t(d['A_prod_src_src']), t(d['read_in_attr']),
        t(d['A_src_in_prod']), t(d['A_src_src']), t(d['grid']).float(), t(d['x_query']).float(),
        t(d['t_query']).float().reshape(-1, 1), float(d['scale_rel']), float(d['scale_t']), return_parts=True)
    for key in ('x_latent', 'read_in', 'x_spatial'):
        assert rel_err(parts[key].numpy(), d[key]) < 2e-6, key
    assert rel_err(y.numpy(), d['y']) < 2e-6 and rel_err(x.numpy(), d['x']) < 2e-6


def test_closure_check_reproduces_the_full_oracle():
    """oracle/closure_check.py (the full-size parity anchor of the GPU tests and of bench.py): on a network small enough for
    the whole oracle, the rows of sampled grid nodes computed on their 2-hop closure equal the full computation."""
    from genie_b200 import synth
    from oracle import closure_check as cc
    S, G = 12, 600
    net = synth.Network(S, G, seed=4)
    A = go.build_adjacencies_dense(net.sta, net.grid, 8, 15)
    trv = net.travel_times()
    attr = net.read_in_offsets(30000.0)
    max_t = net.max_moveout()
    P = synth.make_picks(net, 0.0, 400.0, seed=5, false_per_sta_min=6.0)
    sd = go.init_state(seed=2)
    t0 = 100.0
    Sl, Mk, parts = go.input_scatter(P, t0, np.arange(S), S, A[5].numpy(), trv, max_t, 3.0, 0.3, return_parts=True)
    with torch.no_grad():
        x_lat = go.data_aggregation(sd, 'DataAggregation.', torch.from_numpy(Sl), torch.from_numpy(Mk), A[2], A[3])
        r = go.bipartite_read_in(sd, 'Bipartite_ReadIn.', x_lat, torch.from_numpy(attr), A[4], torch.from_numpy(Mk))
    targets = cc.sample_targets(G, 5, 4)
    assert len(targets) >= 16
    want = cc.oracle_on_closure(sd, A[0], A[1], S, G, targets, P, t0, lambda n: trv[n],
                                lambda n: attr.reshape(G, S, 3)[n].reshape(-1, 3), max_t, 3.0, 0.3)
    assert want['n_closure'] < G                       # a real sub-network, not the whole grid
    rep = cc.compare(want, parts['time_bin'][want['nodes']], Sl[want['nodes']], Mk[want['nodes']],
                     x_lat.numpy()[want['nodes']], r.numpy()[targets])
    assert rep['time_bin_equal'] and rep['mask_equal'] and rep['slice_max_abs'] == 0.0
    assert rep['max_rel'] < 1e-6, rep
    # the cheap tail on the full grid from the full read-in table reproduces forward_fixed_source
    xq = torch.from_numpy(np.random.default_rng(0).uniform(0, net.width, (50, 3))).float()
    xq[:, 2] = -xq[:, 2] / net.width * 40000.0
    tq = torch.arange(-3.0, 3.01, 0.75).reshape(-1, 1)
    grid = torch.from_numpy(net.grid).float()
    y0, x0 = go.forward_fixed_source(sd, torch.from_numpy(Sl), torch.from_numpy(Mk), A[2], A[3], torch.from_numpy(attr), A[4],
                                     A[1], grid, xq, tq, 30000.0, 9.0)
    y1, x1 = cc.oracle_tail(sd, r, A[1], grid, xq, tq, 30000.0, 9.0)
    assert np.array_equal(y0.numpy(), y1) and np.array_equal(x0.numpy(), x1)
    # the row-wise metric is the stricter one
    a = np.array([[1.0, 1e-3], [1e-3, 1e-3]]); b = a.copy(); b[1, 1] += 1e-6
    assert cc.rowwise_rel(b, a) > 100 * cc.global_rel(b, a)


def test_input_variants_match_reference():
    """a1 with use_sign_input (process_utils.py:610-614) and with per-pair travel times from `trv_pairwise` (:594-596):
    the oracle against the unmodified reference (oracle/gen_golden.py input_variants)."""
    d, _ = load_golden('input_variants_12of14x60')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A = np.stack((np.tile(np.arange(S), G), np.repeat(np.arange(G), S)), axis=0)
    args = (d['picks'], float(d['t0']), d['ind_use'], d['sta'].shape[0], A, d['trv_times'], float(d['max_t']),
            float(d['kernel_sig_t']), float(d['dt']))
    for tag, kw in (('plain', {}), ('sign', dict(use_sign_input=True)), ('pairwise', dict(trv_node=d['trv_pairwise'])),
                    ('pairwise_sign', dict(trv_node=d['trv_pairwise'], use_sign_input=True))):
        Sl, Mk = go.input_scatter(*args, **kw)
        assert np.array_equal(Sl, d['Slice_' + tag]), tag
        assert np.array_equal(Mk, d['Mask_' + tag]), tag
    assert (d['Slice_sign'] < 0).sum() > 100 and np.array_equal(np.abs(d['Slice_sign']) > 0.01, d['Mask_sign'] > 0)


@pytest.mark.parametrize('step_size', ['half', 'full'])
def test_streaming_loop_matches_reference(step_size):
    """The caller-side loop of process_continuous_days.py:757-813 (executed verbatim by oracle/gen_golden.py `streaming`):
    oracle.continuous_day_stack reproduces the reference's Out_2, including the windows it skips for lack of picks."""
    d, sd = load_golden('streaming_10x100')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src = torch.from_numpy(d['A_sta_sta']), torch.from_numpy(d['A_src_src'])
    A_ps = (A_sta.repeat(1, G) + S * torch.arange(G).repeat_interleave(A_sta.shape[1]).view(1, -1)).contiguous()
    A_pg = (S * A_src.repeat(1, S) + torch.arange(S).repeat_interleave(A_src.shape[1]).view(1, -1)).contiguous()
    A_sip = torch.stack((torch.arange(S * G), torch.arange(G).repeat_interleave(S)), dim=0)
    A_sis = np.stack((np.tile(np.arange(S), G), np.repeat(np.arange(G), S)), axis=0)
    out, n_done = go.continuous_day_stack(
        sd, d['picks'], d['tsteps_' + step_size], d['tsteps_abs_' + step_size], d['ind_use'], d['sta'].shape[0], A_sis,
        d['trv_times'], float(d['max_t']), float(d['kernel_sig_t']), float(d['dt']), A_ps, A_pg, torch.from_numpy(d['read_in_attr']),
        A_sip, A_src, torch.from_numpy(d['grid']).float(), torch.from_numpy(d['x_query']).float(), float(d['scale_rel']),
        float(d['scale_t']), t_win=float(d['t_win']), dt_win=float(d['dt_win']), step_size=step_size)
    want = d['Out_2_' + step_size]
    assert 0 < n_done < len(d['tsteps_' + step_size])          # some windows were skipped (no picks), some processed
    assert np.abs(out - want).max() <= 1e-6 * max(np.abs(want).max(), 1e-30)
    cols = np.abs(want).sum(axis=0) > 0
    assert np.array_equal(np.abs(out).sum(axis=0) > 0, cols)


@pytest.mark.parametrize('name', ['graphdd_12x9', 'graphdd_10x14_memory'])
def test_graphdd_oracle_matches_reference(name):
    """GraphDD's GNN_Location (Relocation/train_double_difference_model.py:333-536; class definitions executed as they are by
    oracle/gen_golden.py `graphdd`): the oracle's restatement reproduces all four outputs."""
    from oracle import graphdd_oracle as gd
    d, sd = load_golden(name)
    t = lambda k: torch.from_numpy(d[k])
    mem = t('memory') if int(d['use_memory']) else None
    out = gd.gnn_location(sd, t('x'), t('mask'), t('A_in_pick'), t('A_in_src'), t('A_src_in_product'), t('A_sta_in_product'),
                          t('A_src_in_sta'), t('locs').float(), t('srcs').float(), memory=mem)
    for i, o in enumerate(out):
        assert o.shape == d['out%d' % i].shape
        assert rel_err(o.numpy(), d['out%d' % i]) < 1e-5, i
