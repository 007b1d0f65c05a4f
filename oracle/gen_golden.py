"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

    python oracle/gen_golden.py synthetic      # seeded synthetic networks, reference default-initialised weights
    python oracle/gen_golden.py synthetic_edges  # same, `use_updated_model_definition: True` (DataAggregationEdges)
    python oracle/gen_golden.py synthetic_abspos # same, `use_absolute_pos: True` (station / source positions join Slice)
    python oracle/gen_golden.py legacy_input     # a1': extract_inputs_from_data_fixed_grids_with_phase_type
    python oracle/gen_golden.py association      # forward_fixed incl. the association branch
    python oracle/gen_golden.py dense_adjacencies # the dense graph builder incl. the time-pointer re-indexing (process_utils.py:701-742)
    python oracle/gen_golden.py graphdd          # GraphDD's GNN_Location (Relocation/train_double_difference_model.py:333-536)
    python oracle/gen_golden.py streaming        # the script-body loop of process_continuous_days.py:757-813, executed verbatim
    python oracle/gen_golden.py input_variants   # a1 with use_sign_input / trv_times=None (process_utils.py:594-614)
    python oracle/gen_golden.py subgraph         # sub-graph mode builder (process_utils.py:744-849) + one window on it
    python oracle/gen_golden.py ferndale       # Examples/Ferndale.zip: real stations/grids/picks + trained checkpoint

The reference classes (`/root/reference/Code/module.py`, `process_utils.py`) are imported as they are, with
`oracle/refshim/` supplying the third-party packages that cannot be installed offline (see its README).  The reference
reads `config.yaml` / `train_config.yaml` from the current directory at import time (module.py:27-46), so each mode
runs from a scratch directory holding the matching YAML files.  `/root/reference` does not exist on the GPU box — only
the committed fixtures travel.
"""
import os
import sys
import shutil
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference'
GOLD = os.path.join(REPO, 'tests', 'golden')


def _import_reference(code_dir, workdir):
    os.chdir(workdir)
    sys.path.insert(0, os.path.join(HERE, 'refshim'))
    sys.path.insert(0, code_dir)
    sys.path.insert(0, REPO)
    import torch
    import module
    import process_utils
    from torch_geometric.data import Data
    torch.set_grad_enabled(False)
    return torch, module, process_utils, Data


def _hook_outputs(mz):
    store = {}

    def mk(name):
        def hook(_m, _inp, out):
            store.setdefault(name, []).append(out.detach().clone())
        return hook

    for name in ('DataAggregation', 'Bipartite_ReadIn', 'SpatialAggregation1', 'SpatialAggregation2',
                 'SpatialAggregation3', 'SpatialDirect', 'SpatialAttention', 'TemporalAttention'):
        getattr(mz, name).register_forward_hook(mk(name))
    return store


def _pack(sd):
    return {'sd/' + k: v.detach().cpu().numpy() for k, v in sd.items()}


def _run_reference_window(torch, module, pu, Data, mz, locs, ind_use, grid, trv_times, P, t0, max_t, sig, dt,
                          k_sta, k_spc, attr_scale, x_query, t_query, identity):
    """set-up (process_continuous_days.py:627-634) + one window (:776-797) through the reference's own functions."""
    S = len(ind_use)
    G = grid.shape[0]
    n_locs = locs.shape[0]
    # graphs: extract_inputs_adjacencies (process_utils.py:701); the time-edge pointer arguments are association-only,
    # so they are given minimal dummies (k_time_edges = 1, one reference time).
    graph_params = [k_sta, k_spc, 1]
    dummy_ptr = np.zeros(n_locs, dtype='int')
    out = pu.extract_inputs_adjacencies(None, locs, ind_use, grid, None, np.zeros(1), dummy_ptr, dummy_ptr,
                                        identity, graph_params, device='cpu')
    A_sta_sta, A_src_src, A_prod_sta, A_prod_src, A_src_in_prod = out[0:5]
    A_src_in_sta = torch.Tensor(np.concatenate((np.tile(np.arange(S), G).reshape(1, -1),
                                                np.arange(G).repeat(S, axis=0).reshape(1, -1)), axis=0)).long()
    spatial_vals = torch.Tensor(((np.repeat(np.expand_dims(grid, axis=1), S, axis=1)
                                  - np.repeat(np.expand_dims(locs[ind_use], axis=0), G, axis=0)).reshape(-1, 3))
                                / attr_scale)
    A_src_in_edges = Data(x=spatial_vals, edge_index=A_src_in_prod)
    locs_cart = torch.Tensor(identity(locs[ind_use]))
    grid_cart = torch.Tensor(identity(grid))
    mz.set_adjacencies(A_prod_sta, A_prod_src, A_src_in_edges, None, A_src_in_sta, A_src_src, None, None, None, None,
                       locs_cart, grid_cart)
    # input
    [Inpts, Masks], [lp_t, lp_s, lp_p, _] = pu.extract_input_from_data(
        None, P, np.array([t0]), ind_use, locs, grid, A_src_in_sta.numpy(), trv_times=trv_times, max_t=max_t,
        kernel_sig_t=sig, dt=dt, device='cpu')
    emb = pu.extract_input_from_data(None, P, np.array([t0]), ind_use, locs, grid, A_src_in_sta.numpy(),
                                     trv_times=trv_times, max_t=max_t, kernel_sig_t=sig, dt=dt,
                                     return_embedding=True, device='cpu')
    embed_p, embed_s, ind_unique, abs_time_ref, n_ts, n_su = emb
    Slice, Mask = Inpts[0], Masks[0]
    store = _hook_outputs(mz)
    y, x = mz.forward_fixed_source(Slice, Mask, torch.Tensor(lp_t[0]), torch.Tensor(lp_s[0]).long(),
                                   torch.Tensor(lp_p[0].reshape(-1, 1)).float(), locs_cart, grid_cart,
                                   torch.Tensor(x_query), torch.Tensor(t_query.reshape(-1, 1)))
    res = dict(
        A_sta_sta=A_sta_sta.numpy(), A_src_src=A_src_src.numpy(), read_in_attr=spatial_vals.numpy(),
        Slice=Slice.numpy(), Mask=Mask.numpy(),
        embed_p=embed_p.numpy().reshape(n_su, n_ts), embed_s=embed_s.numpy().reshape(n_su, n_ts),
        ind_unique=np.asarray(ind_unique), ref0=np.float64(abs_time_ref[0]), n_ts=np.int64(n_ts),
        x_latent=store['DataAggregation'][0].numpy(), read_in=store['Bipartite_ReadIn'][0].numpy(),
        sa1=store['SpatialAggregation1'][0].numpy(), sa2=store['SpatialAggregation2'][0].numpy(),
        x_spatial=store['SpatialAggregation3'][0].numpy(), y_latent=store['SpatialDirect'][0].numpy(),
        x_query_embed=store['SpatialAttention'][0].numpy(), y=y.numpy(), x=x.numpy())
    return res


def synthetic(edges=False, abs_pos=False):
    """edges=True: the reference's `use_updated_model_definition: True` classes (DataAggregationEdges, module.py:102-174,
    1024-1111); the YAML copy in the scratch directory is switched, the reference sources are untouched."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    if edges:
        import re
        cfg = open(os.path.join(work, 'config.yaml')).read()
        cfg, n = re.subn(r'(?m)^use_updated_model_definition:\s*\w+', 'use_updated_model_definition: True', cfg)
        assert n == 1
        open(os.path.join(work, 'config.yaml'), 'w').write(cfg)
    if abs_pos:
        import re
        cfg = open(os.path.join(work, 'config.yaml')).read()
        cfg, n = re.subn(r'(?m)^use_absolute_pos:\s*\w+', 'use_absolute_pos: True', cfg)
        assert n == 1
        open(os.path.join(work, 'config.yaml'), 'w').write(cfg)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    assert bool(module.use_updated_model_definition) == edges and bool(module.use_absolute_pos) == abs_pos
    from genie_b200 import synth

    def identity(x):
        return x

    cases = [  # name, S_all, n_use, G, k_sta, k_spc, Q, seed
        ('c1_10x100', 10, 10, 100, 8, 15, 64, 0),           # BASELINE.json configs[0]
        ('mid_36of40x300', 40, 36, 300, 8, 15, 128, 3),     # station subset (ind_use != arange), k_s < S-2
        ('small_6x40', 6, 6, 40, 8, 15, 16, 5),             # k_sta clipped to S-2 (process_utils.py:712)
    ]
    if abs_pos:
        cases = [('c1_10x100_abspos', 10, 10, 100, 8, 15, 64, 0), ('mid_36of40x300_abspos', 40, 36, 300, 8, 15, 128, 3)]
    if edges:
        cases = [('c1_10x100_edges', 10, 10, 100, 8, 15, 64, 0), ('mid_36of40x300_edges', 40, 36, 300, 8, 15, 128, 3)]
    for name, S_all, n_use, G, k_sta, k_spc, Q, seed in cases:
        net = synth.Network(S_all, G, seed=seed, width_km=60.0 if S_all <= 10 else 120.0)
        rng = np.random.default_rng(100 + seed)
        ind_use = np.sort(rng.choice(S_all, size=n_use, replace=False))
        max_t = net.max_moveout()
        sig, dt = 3.0, float(np.round(3.0 / 10.0, 2))
        P = synth.make_picks(net, 0.0, 600.0, seed=seed + 1, events_per_3h=400.0, false_per_sta_min=2.0)
        t0 = 200.0 + 3.0 * seed
        trv_times = net.travel_times()
        torch.manual_seed(seed)
        mz = module.GCN_Detection_Network_extended(identity, identity, device='cpu')
        mz.eval()
        x_query = np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q),
                            rng.uniform(-40000.0, 0.0, Q)), axis=1)
        t_query = np.arange(-3.0, 3.0 + 0.75, 0.75)
        attr_scale = np.array([net.width, net.width, 42000.0]).reshape(1, -1)
        res = _run_reference_window(torch, module, pu, Data, mz, net.sta, ind_use, net.grid, trv_times, P, t0, max_t,
                                    sig, dt, k_sta, k_spc, attr_scale, x_query, t_query, identity)
        res.update(_pack(mz.state_dict()))
        res.update(sta=net.sta, grid=net.grid, ind_use=ind_use, trv_times=trv_times, picks=P, t0=np.float64(t0),
                   max_t=np.float64(max_t), kernel_sig_t=np.float64(sig), dt=np.float64(dt),
                   k_sta=np.int64(k_sta), k_spc=np.int64(k_spc), scale_rel=np.float64(module.scale_rel),
                   scale_t=np.float64(module.scale_t), x_query=x_query, t_query=t_query, attr_scale=attr_scale)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **res)
        print(name, 'P=%d picks=%d Slice.sum=%.6f nnz=%d x_latent.abs=%.6f y.max=%.6f x.max=%.6f' % (
            res['Slice'].shape[0], len(P), res['Slice'].sum(), (res['Slice'] != 0).sum(),
            np.abs(res['x_latent']).sum(), res['y'].max(), res['x'].max()))


def legacy_input():
    """a1': extract_inputs_from_data_fixed_grids_with_phase_type (process_utils.py:102-308), the nearest-pick input
    features used when `use_updated_input: False` and in training.  Two time samples per call exercise the batch offsets."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from scipy.spatial import cKDTree
    from genie_b200 import synth
    for name, S_all, n_use, G, seed, t_win, sig in (('legacy_12of14x60', 14, 12, 60, 4, 10.0, 3.0),
                                                    ('legacy_8x30_short', 8, 8, 30, 6, 5.0, 2.0)):
        net = synth.Network(S_all, G, seed=seed, width_km=60.0 if 'short' not in name else 12.0)
        rng = np.random.default_rng(200 + seed)
        ind_use = np.sort(rng.choice(S_all, size=n_use, replace=False))
        max_t = net.max_moveout()
        P = synth.make_picks(net, 0.0, 400.0, seed=seed + 1, events_per_3h=600.0, false_per_sta_min=3.0)
        P = P[np.argsort(P[:, 0])]
        time_samples = np.array([120.0 + seed, 131.5 + seed])
        trv_times = net.travel_times()                       # [G, S_all, 2]
        tree = cKDTree(P[:, 0][:, None])
        rng_lat = [0.0, 1.0]
        [Inpts, Masks], [lp_t, lp_s, lp_p, lp_m] = pu.extract_inputs_from_data_fixed_grids_with_phase_type(
            None, net.sta, ind_use, P, P[:, 4], tree, time_samples, net.grid, trv_times, rng_lat, rng_lat, rng_lat, max_t,
            None, [8, 15, 10], [t_win, sig], None, None)
        res = dict(sta=net.sta, grid=net.grid, ind_use=ind_use, picks=P, time_samples=time_samples, trv_times=trv_times,
                   max_t=np.float64(max_t), t_win=np.float64(t_win), kernel_sig_t=np.float64(sig))
        for i in range(len(time_samples)):
            res['Inpts%d' % i], res['Masks%d' % i] = np.asarray(Inpts[i]), np.asarray(Masks[i])
            res['lp_times%d' % i], res['lp_stations%d' % i] = np.asarray(lp_t[i]), np.asarray(lp_s[i])
            res['lp_phases%d' % i], res['lp_meta%d' % i] = np.asarray(lp_p[i]), np.asarray(lp_m[i])
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **res)
        print(name, 'max_t=%.2f picks=%d' % (max_t, len(P)), [float(np.asarray(x).sum()) for x in Inpts],
              [int((np.asarray(x) > 0).sum()) for x in Inpts], Inpts[0].dtype, Inpts[0].shape)


def association(edges=False, abs_pos=False):
    """§8f rank 2: `forward_fixed` (module.py:963-997) of the UNMODIFIED reference — front end + heads + association branch
    (BipartiteGraphReadOutOperator, DataAggregationAssociationPhase, LocalSliceLgCollapse{P,S}, Arrivals).  Set-up as in
    process_continuous_days.py:627-634; the time-pointer tables through the reference's own
    compute_time_embedding_vectors (process_utils.py:851-877)."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    if edges or abs_pos:
        import re
        cfg = open(os.path.join(work, 'config.yaml')).read()
        for on, key in ((edges, 'use_updated_model_definition'), (abs_pos, 'use_absolute_pos')):
            if on:
                cfg, n = re.subn(r'(?m)^%s:\s*\w+' % key, '%s: True' % key, cfg)
                assert n == 1
        open(os.path.join(work, 'config.yaml'), 'w').write(cfg)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    assert bool(module.use_updated_model_definition) == edges and bool(module.use_absolute_pos) == abs_pos
    from genie_b200 import synth

    def identity(x):
        return x

    cases = (('assoc_10x100', 10, 10, 100, 8, 15, 32, 3, 0), ('assoc_18of20x160', 20, 18, 160, 8, 15, 48, 2, 9))
    if edges or abs_pos:      # one case per variant: station subset, 14 x 120
        cases = (('assoc_14of16x120' + ('_edges' if edges else '') + ('_abspos' if abs_pos else ''), 16, 14, 120, 8, 15, 32, 3,
                  4 + int(edges) + 2 * int(abs_pos)),)
    for name, S_all, n_use, G, k_sta, k_spc, Q, n_src, seed in cases:
        net = synth.Network(S_all, G, seed=seed, width_km=60.0 if S_all <= 10 else 90.0)
        rng = np.random.default_rng(300 + seed)
        ind_use = np.sort(rng.choice(S_all, size=n_use, replace=False))
        S = n_use
        max_t = net.max_moveout()
        sig, dt = 3.0, float(np.round(3.0 / 10.0, 2))
        P = synth.make_picks(net, 0.0, 600.0, seed=seed + 1, events_per_3h=400.0, false_per_sta_min=2.0)
        trv_times = net.travel_times()
        locs, grid = net.sta, net.grid
        torch.manual_seed(seed)
        mz = module.GCN_Detection_Network_extended(identity, identity, device='cpu')
        mz.eval()
        graph_params = [k_sta, k_spc, 1]
        dummy_ptr = np.zeros(S_all, dtype='int')
        out = pu.extract_inputs_adjacencies(None, locs, ind_use, grid, None, np.zeros(1), dummy_ptr, dummy_ptr,
                                            identity, graph_params, device='cpu')
        A_sta_sta, A_src_src, A_prod_sta, A_prod_src, A_src_in_prod = out[0:5]
        A_src_in_sta = torch.Tensor(np.concatenate((np.tile(np.arange(S), G).reshape(1, -1),
                                                    np.arange(G).repeat(S, axis=0).reshape(1, -1)), axis=0)).long()
        attr_scale = np.array([net.width, net.width, 42000.0]).reshape(1, -1)
        spatial_vals = torch.Tensor(((np.repeat(np.expand_dims(grid, axis=1), S, axis=1)
                                      - np.repeat(np.expand_dims(locs[ind_use], axis=0), G, axis=0)).reshape(-1, 3))
                                    / attr_scale)
        A_src_in_edges = Data(x=spatial_vals, edge_index=A_src_in_prod)
        A_Lg_in_src = Data(x=spatial_vals, edge_index=torch.Tensor(
            np.ascontiguousarray(np.flip(A_src_in_prod.numpy(), axis=0))).long())     # process_continuous_days.py:632
        tt_use = torch.Tensor(trv_times[:, ind_use, :])                              # [G, S, 2]

        def trv_pairwise(sta_rows, src_rows):
            # synthetic travel times of the (station, source) pairs; rows are in product-node order here
            assert sta_rows.shape[0] == G * S
            return tt_use.reshape(-1, 2)

        A_edges_p, A_edges_s, dt_partition = pu.compute_time_embedding_vectors(
            trv_pairwise, locs[ind_use], grid, A_src_in_sta, max_t, dt_res=sig / 5.0, t_win=sig * 2.0, device='cpu')
        tlatent = tt_use.reshape(-1, 2)
        locs_cart = torch.Tensor(locs[ind_use])
        grid_cart = torch.Tensor(grid)
        mz.set_adjacencies(A_prod_sta, A_prod_src, A_src_in_edges, A_Lg_in_src, A_src_in_sta, A_src_src,
                           torch.Tensor(A_edges_p).long(), torch.Tensor(A_edges_s).long(), torch.Tensor(dt_partition),
                           tlatent, locs_cart, grid_cart)
        x_query = np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q),
                            rng.uniform(-40000.0, 0.0, Q)), axis=1)
        t_query = np.arange(-3.0, 3.0 + 0.75, 0.75)
        # weights: the trained Ferndale checkpoint (Examples/Ferndale.zip), so that y genuinely varies over the grid and the
        # source mask (y.max > 0.01, module.py:983) is a mixture of zeros and ones decided by O(0.1) margins, not by rounding
        with zipfile.ZipFile(os.path.join(REF, 'Examples', 'Ferndale.zip')) as z:
            ck_name = [n for n in z.namelist() if n.endswith('trained_gnn_model_step_20000_ver_1.h5')][0]
            z.extract(ck_name, work)
        ck = torch.load(os.path.join(work, ck_name), map_location='cpu')
        own = mz.state_dict()
        if edges or abs_pos:   # the variants widen some Linear layers: those keep their default initialisation
            ck = {k: v for k, v in ck.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        missing = mz.load_state_dict(ck, strict=False)
        assert not missing.unexpected_keys
        assert edges or abs_pos or all(k.startswith('SpatialAttention.f_queries') for k in missing.missing_keys)
        # window: the first origin time (3 s steps) whose source mask is a clear mixture — at least 8 % of either value and
        # every grid node's y.max at least 2e-3 * max|y| away from the 0.01 threshold
        for t0 in 200.0 + 3.0 * seed + 3.0 * np.arange(120):
            [Inpts, Masks], [lp_t, lp_s, lp_p, _] = pu.extract_input_from_data(
                None, P, np.array([t0]), ind_use, locs, grid, A_src_in_sta.numpy(), trv_times=trv_times, max_t=max_t,
                kernel_sig_t=sig, dt=dt, device='cpu')
            Slice, Mask = Inpts[0], Masks[0]
            y0, _ = mz.forward_fixed_source(Slice, Mask, torch.Tensor(lp_t[0]), torch.Tensor(lp_s[0]).long(),
                                            torch.Tensor(lp_p[0].reshape(-1, 1)).float(), locs_cart, grid_cart,
                                            torch.Tensor(x_query), torch.Tensor(t_query.reshape(-1, 1)))
            ym = y0[:, :, 0].max(1)[0].numpy()
            frac = float((ym > 0.01).mean())
            if 0.08 < frac < 0.92 and np.abs(ym - 0.01).min() > 2e-3 * np.abs(y0.numpy()).max() and len(lp_t[0]) >= 12:
                break
        else:
            raise SystemExit('no window with a mixed source mask')
        # association queries: n_src sources at grid nodes (so their travel times are rows of trv_times), origin times
        # relative to t0 inside and outside the 2*eps keep window of the null arrival (module.py:723-727)
        isrc = rng.choice(G, size=n_src, replace=False)
        x_query_src = grid[isrc]
        tq_sample = np.linspace(0.0, 2.5 * module.eps, n_src)
        trv_out_q = trv_times[isrc][:, ind_use, :]
        store = _hook_outputs(mz)
        extra = {}
        for nm in ('BipartiteGraphReadOutOperator', 'DataAggregationAssociationPhase', 'LocalSliceLgCollapseP',
                   'LocalSliceLgCollapseS', 'Arrivals'):
            getattr(mz, nm).register_forward_hook(
                (lambda key: (lambda _m, _i, o: extra.setdefault(key, o)))(nm))
        y, x, arv_p, arv_s = mz.forward_fixed(
            Slice, Mask, torch.Tensor(lp_t[0]), torch.Tensor(lp_s[0]).long(), torch.Tensor(lp_p[0].reshape(-1, 1)).long(),
            locs_cart, grid_cart, torch.Tensor(x_query), torch.Tensor(x_query_src), torch.Tensor(t_query.reshape(-1, 1)),
            torch.Tensor(tq_sample), torch.Tensor(trv_out_q))
        res = dict(
            A_sta_sta=A_sta_sta.numpy(), A_src_src=A_src_src.numpy(), read_in_attr=spatial_vals.numpy(),
            Slice=Slice.numpy(), Mask=Mask.numpy(), A_edges_p=np.asarray(A_edges_p).astype('int64'),
            A_edges_s=np.asarray(A_edges_s).astype('int64'), dt_partition=np.asarray(dt_partition, dtype='float64'),
            tlatent=tlatent.numpy(), tpick=np.asarray(lp_t[0]), ipick=np.asarray(lp_s[0]).astype('int64'),
            phase_label=np.asarray(lp_p[0]).astype('int64'), x_query=x_query, x_query_src=x_query_src,
            t_query=t_query, tq_sample=tq_sample, trv_out_q=trv_out_q,
            x_latent=store['DataAggregation'][0].numpy(), x_spatial=store['SpatialAggregation3'][0].numpy(),
            y_latent=store['SpatialDirect'][0].numpy(), x_src=store['SpatialAttention'][1].numpy(),
            assoc_s0=extra['BipartiteGraphReadOutOperator'][0].numpy(),
            mask_out_1=extra['BipartiteGraphReadOutOperator'][1].numpy(),
            assoc_s=extra['DataAggregationAssociationPhase'].numpy(),
            arv_p_embed=extra['LocalSliceLgCollapseP'].numpy(), arv_s_embed=extra['LocalSliceLgCollapseS'].numpy(),
            y=y.numpy(), x=x.numpy(), arv_p=arv_p.numpy(), arv_s=arv_s.numpy())
        res.update(_pack(mz.state_dict()))
        res.update(sta=locs, grid=grid, ind_use=ind_use, trv_times=trv_times, picks=P, t0=np.float64(t0),
                   max_t=np.float64(max_t), kernel_sig_t=np.float64(sig), dt=np.float64(dt), k_sta=np.int64(k_sta),
                   k_spc=np.int64(k_spc), scale_rel=np.float64(module.scale_rel), scale_t=np.float64(module.scale_t),
                   eps=np.float64(module.eps), attr_scale=attr_scale)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **res)
        print(name, 'P=%d picks=%d mask_out.mean=%.3f margin=%.2e s.abs=%.6f arv_p.abs=%.6f arv_s.abs=%.6f y.max=%.6f' % (
            Slice.shape[0], len(lp_t[0]), float(res['mask_out_1'].mean()),
            float(np.abs(res['y'][:, :, 0].max(1) - 0.01).min()), np.abs(res['assoc_s']).sum(),
            np.abs(res['arv_p']).sum(), np.abs(res['arv_s']).sum(), res['y'].max()))


def subgraph():
    """extract_inputs_adjacencies_subgraph (process_utils.py:744-849) of the unmodified reference on two synthetic networks
    (Cartesian metres, ftrns1 = identity) + one window of inputs and the front end on the resulting explicit product graph."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from genie_b200 import synth

    def identity(x):
        return x

    # the third case is ragged: two nearest stations per source and a 3 km radius leave stations with a single source or
    # none at all (the reference's empty-slice branch, process_utils.py:832-834) and sources whose station list has no
    # station-graph edge inside it
    for name, S, G, k_sta, k_spc, k_pairs, max_deg, seed in (('subgraph_14x60', 14, 60, 8, 15, 5, 0.2, 2),
                                                           ('subgraph_30x200', 30, 200, 10, 15, 8, 0.25, 6),
                                                           ('subgraph_12x40_ragged', 12, 40, 8, 15, 2, 0.03, 8)):
        net = synth.Network(S, G, seed=seed, width_km=60.0 if S < 20 else 100.0)
        out = pu.extract_inputs_adjacencies_subgraph(net.sta, net.grid, identity, identity, max_deg_offset=max_deg,
                                                     k_nearest_pairs=k_pairs, k_sta_edges=k_sta, k_spc_edges=k_spc, device='cpu')
        A_sta_sta, A_src_src, A_prod_sta, A_prod_src, A_src_in_prod, A_src_in_sta = out
        P = A_src_in_sta.shape[1]
        assert P < S * G
        # one window through the reference on this explicit graph (process_continuous_days.py:640-648, 776-797)
        torch.manual_seed(seed)
        mz = module.GCN_Detection_Network_extended(identity, identity, device='cpu')
        mz.eval()
        attr_scale = np.array([net.width, net.width, 42000.0]).reshape(1, -1)
        spatial_vals = torch.Tensor((net.grid[A_src_in_prod[1].numpy()] - net.sta[A_src_in_sta[0][A_src_in_prod[0]].numpy()])
                                    / attr_scale)                                                   # :642
        mz.set_adjacencies(A_prod_sta, A_prod_src, Data(x=spatial_vals, edge_index=A_src_in_prod), None, A_src_in_sta,
                           A_src_src, None, None, None, None, torch.Tensor(net.sta), torch.Tensor(net.grid))
        Pk = synth.make_picks(net, 0.0, 600.0, seed=seed + 1, events_per_3h=400.0, false_per_sta_min=2.0)
        t0, sig, dt = 210.0, 3.0, 0.3
        max_t = net.max_moveout()
        trv_times = net.travel_times()
        [Inpts, Masks], [lp_t, lp_s, lp_p, _] = pu.extract_input_from_data(
            None, Pk, np.array([t0]), np.arange(S), net.sta, net.grid, A_src_in_sta.numpy(), trv_times=trv_times, max_t=max_t,
            kernel_sig_t=sig, dt=dt, device='cpu')
        rng = np.random.default_rng(seed)
        Q = 40
        x_query = np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q), rng.uniform(-40000.0, 0.0, Q)), axis=1)
        t_query = np.arange(-3.0, 3.0 + 0.75, 0.75)
        store = _hook_outputs(mz)
        y, x = mz.forward_fixed_source(Inpts[0], Masks[0], torch.Tensor(lp_t[0]), torch.Tensor(lp_s[0]).long(),
                                       torch.Tensor(lp_p[0].reshape(-1, 1)).float(), torch.Tensor(net.sta), torch.Tensor(net.grid),
                                       torch.Tensor(x_query), torch.Tensor(t_query.reshape(-1, 1)))
        res = dict(sta=net.sta, grid=net.grid, ind_use=np.arange(S), k_sta=np.int64(k_sta), k_spc=np.int64(k_spc),
                   k_nearest_pairs=np.int64(k_pairs), max_deg_offset=np.float64(max_deg), A_sta_sta=A_sta_sta.numpy(),
                   A_src_src=A_src_src.numpy(), A_prod_sta_sta=A_prod_sta.numpy(), A_prod_src_src=A_prod_src.numpy(),
                   A_src_in_prod=A_src_in_prod.numpy(), A_src_in_sta=A_src_in_sta.numpy(), read_in_attr=spatial_vals.numpy(),
                   picks=Pk, t0=np.float64(t0), max_t=np.float64(max_t), kernel_sig_t=np.float64(sig), dt=np.float64(dt),
                   trv_times=trv_times, Slice=Inpts[0].numpy(), Mask=Masks[0].numpy(), x_query=x_query, t_query=t_query,
                   x_latent=store['DataAggregation'][0].numpy(), read_in=store['Bipartite_ReadIn'][0].numpy(),
                   x_spatial=store['SpatialAggregation3'][0].numpy(), y=y.numpy(), x=x.numpy(),
                   scale_rel=np.float64(module.scale_rel), scale_t=np.float64(module.scale_t))
        res.update(_pack(mz.state_dict()))
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **res)
        print(name, 'P=%d of %d, sta edges %d, src edges %d, Slice nnz %d, x_latent.abs %.4f' % (
            P, S * G, A_prod_sta.shape[1], A_prod_src.shape[1], int((Inpts[0] != 0).sum()), float(np.abs(res['x_latent']).sum())))


def dense_adjacencies():
    """extract_inputs_adjacencies (process_utils.py:701-742) of the unmodified reference incl. the time-pointer re-indexing of a
    station subset (:723-734): all eight returned objects."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from genie_b200 import synth

    def identity(x):
        return x

    S_all, n_use, G, k_sta, k_spc, k_time, len_dt, seed = 12, 9, 30, 8, 15, 3, 7, 12
    net = synth.Network(S_all, G, seed=seed, width_km=60.0)
    rng = np.random.default_rng(seed)
    ind_use = np.sort(rng.choice(S_all, size=n_use, replace=False))
    # pointer tables over ALL stations: row (station, time bin, k) names a product node g*S_all + station (utils.py:976)
    sta_of = np.repeat(np.arange(S_all), len_dt * k_time)
    ptr_p = rng.integers(0, G, S_all * len_dt * k_time) * S_all + sta_of
    ptr_s = rng.integers(0, G, S_all * len_dt * k_time) * S_all + sta_of
    ref_t = np.linspace(-6.0, 12.0, len_dt)
    out = pu.extract_inputs_adjacencies(None, net.sta, ind_use, net.grid, None, ref_t, ptr_p, ptr_s, identity,
                                        [k_sta, k_spc, k_time], device='cpu')
    names = ('A_sta_sta', 'A_src_src', 'A_prod_sta_sta', 'A_prod_src_src', 'A_src_in_prod', 'A_edges_time_p', 'A_edges_time_s',
             'A_edges_ref')
    res = {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in zip(names, out)}
    res.update(sta=net.sta, grid=net.grid, ind_use=ind_use, ptr_p=ptr_p, ptr_s=ptr_s, ref_t=ref_t, k_sta=np.int64(k_sta),
               k_spc=np.int64(k_spc), k_time=np.int64(k_time))
    np.savez_compressed(os.path.join(GOLD, 'dense_adjacencies_9of12x30.npz'), **res)
    print('dense_adjacencies', {k: (res[k].shape, res[k].dtype) for k in names})


def input_variants():
    """a1 with the switches of extract_input_from_data (process_utils.py:460) the day-processing script uses:
    `use_sign_input=True` (process_continuous_days.py:776 passes the flag; :610-614) and `trv_times=None` (travel times from
    the `trv_pairwise` callable, :594-596), on a station subset — through the unmodified reference."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from genie_b200 import synth
    S_all, n_use, G, seed = 14, 12, 60, 11
    net = synth.Network(S_all, G, seed=seed, width_km=120.0)
    rng = np.random.default_rng(100 + seed)
    ind_use = np.sort(rng.choice(S_all, size=n_use, replace=False))
    max_t = net.max_moveout()
    sig, dt = 3.0, float(np.round(3.0 / 10.0, 2))
    P = synth.make_picks(net, 0.0, 600.0, seed=seed + 1, events_per_3h=400.0, false_per_sta_min=3.0)
    t0 = 233.0
    trv_times = net.travel_times()
    A_src_in_sta = np.concatenate((np.tile(np.arange(n_use), G).reshape(1, -1),
                                   np.arange(G).repeat(n_use, axis=0).reshape(1, -1)), axis=0)

    def trv_pairwise(sta, src):                        # homogeneous half-space in fp32 torch, as a travel-time net would return
        d = torch.norm(sta - src, dim=1, keepdim=True)
        return torch.cat((d / 6000.0, d / 3464.0), dim=1)

    res = dict(sta=net.sta, grid=net.grid, ind_use=ind_use, trv_times=trv_times, picks=P, t0=np.float64(t0),
               max_t=np.float64(max_t), kernel_sig_t=np.float64(sig), dt=np.float64(dt))
    for tag, kw in (('sign', dict(trv_times=trv_times, use_sign_input=True)),
                    ('plain', dict(trv_times=trv_times)),
                    ('pairwise', dict(trv_times=None)),
                    ('pairwise_sign', dict(trv_times=None, use_sign_input=True))):
        [Inpts, Masks], _ = pu.extract_input_from_data(trv_pairwise, P, np.array([t0]), ind_use, net.sta, net.grid,
                                                       A_src_in_sta, max_t=max_t, kernel_sig_t=sig, dt=dt, device='cpu', **kw)
        res['Slice_' + tag], res['Mask_' + tag] = Inpts[0].numpy(), Masks[0].numpy()
    sta_t, grid_t = torch.Tensor(net.sta[ind_use]), torch.Tensor(net.grid)
    res['trv_pairwise'] = trv_pairwise(sta_t[A_src_in_sta[0]], grid_t[A_src_in_sta[1]]).numpy()       # [P,2] fp32
    np.savez_compressed(os.path.join(GOLD, 'input_variants_12of14x60.npz'), **res)
    print('input_variants', {k: (float(res[k].sum()), int((res[k] != 0).sum())) for k in res if k.startswith('Slice_')},
          'negatives with sign input:', int((res['Slice_sign'] < 0).sum()))


def streaming():
    """The caller-side streaming loop (process_continuous_days.py:757-813: per origin-time sample extract_input_from_data ->
    forward_fixed_source -> `Out_2[:, ip_need] += ...`).  The loop lives in the reference's SCRIPT BODY and cannot be imported;
    its source lines are read from the reference file at generation time and executed verbatim (one tab of indentation
    removed) in a namespace holding the same variable names the script defines, with the unmodified reference module /
    process_utils behind them.  Result: Out_2 for a short synthetic pick stream, step_size 'half' and 'full'."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from scipy.spatial import cKDTree
    from genie_b200 import synth
    lines = open(os.path.join(REF, 'Code', 'process_continuous_days.py')).read().split('\n')
    beg = [i for i, l in enumerate(lines) if l.strip().startswith('Out_2 = np.zeros((X_query_cart.shape[0], len(tsteps_abs)))')]
    end = [i for i, l in enumerate(lines) if l.strip().startswith('Out_2_sparse = np.concatenate(')]
    assert len(beg) == 1 and len(end) == 1 and 750 < beg[0] < end[0] < 830
    body = '\n'.join(l[1:] if l.startswith('\t') else l for l in lines[beg[0]:end[0] + 1])

    def identity(x):
        return x

    S_all, n_use, G, k_sta, k_spc, Q, seed = 10, 10, 100, 8, 15, 48, 0
    net = synth.Network(S_all, G, seed=seed, width_km=60.0)
    rng = np.random.default_rng(100 + seed)
    ind_use = np.arange(S_all)
    max_t = net.max_moveout()
    sig, dt_embed = 3.0, float(np.round(3.0 / 10.0, 2))
    P = synth.make_picks(net, 0.0, 240.0, seed=seed + 1, events_per_3h=900.0, false_per_sta_min=1.5)
    P = P[~((P[:, 0] > 90.0) & (P[:, 0] < 90.0 + max_t + 30.0))]                   # a quiet stretch: windows without picks (:787)
    trv_times = net.travel_times()
    torch.manual_seed(seed)
    mz = module.GCN_Detection_Network_extended(identity, identity, device='cpu')
    mz.eval()
    x_query = np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q), rng.uniform(-40000.0, 0.0, Q)), axis=1)
    attr_scale = np.array([net.width, net.width, 42000.0]).reshape(1, -1)
    # set-up as process_continuous_days.py:627-634
    out = pu.extract_inputs_adjacencies(None, net.sta, ind_use, net.grid, None, np.zeros(1), np.zeros(S_all, dtype='int'),
                                        np.zeros(S_all, dtype='int'), identity, [k_sta, k_spc, 1], device='cpu')
    A_sta_sta, A_src_src, A_prod_sta, A_prod_src, A_src_in_prod = out[0:5]
    A_src_in_sta = torch.Tensor(np.concatenate((np.tile(np.arange(n_use), G).reshape(1, -1),
                                                np.arange(G).repeat(n_use, axis=0).reshape(1, -1)), axis=0)).long()
    spatial_vals = torch.Tensor(((np.repeat(np.expand_dims(net.grid, axis=1), n_use, axis=1)
                                  - np.repeat(np.expand_dims(net.sta[ind_use], axis=0), G, axis=0)).reshape(-1, 3)) / attr_scale)
    mz.set_adjacencies(A_prod_sta, A_prod_src, Data(x=spatial_vals, edge_index=A_src_in_prod), None, A_src_in_sta, A_src_src,
                       None, None, None, None, torch.Tensor(net.sta[ind_use]), torch.Tensor(net.grid))
    res = dict(sta=net.sta, grid=net.grid, ind_use=ind_use, trv_times=trv_times, picks=P, max_t=np.float64(max_t),
               kernel_sig_t=np.float64(sig), dt=np.float64(dt_embed), k_sta=np.int64(k_sta), k_spc=np.int64(k_spc),
               scale_rel=np.float64(module.scale_rel), scale_t=np.float64(module.scale_t), x_query=x_query,
               attr_scale=attr_scale, read_in_attr=spatial_vals.numpy(), A_sta_sta=A_sta_sta.numpy(), A_src_src=A_src_src.numpy())
    res.update(_pack(mz.state_dict()))
    day_len = 240.0
    for step_size in ('half', 'full'):
        # the script's window geometry (:360-379, :411-412, :534, :571) for n_resolution = 9, t_win = 6 s
        n_resolution, t_win = 9, 6.0
        dt_win = np.diff(np.linspace(-t_win / 2.0, t_win / 2.0, n_resolution))[0]
        step = n_resolution * dt_win if step_size == 'full' else int(np.floor(n_resolution / 2)) * dt_win
        n_overlap = 1.0 if step_size == 'full' else 2.0
        tsteps = np.arange(np.maximum(0.0, P[:, 0].min() - max_t), np.minimum(day_len, P[:, 0].max()), step)
        tsteps_abs = np.arange(-t_win / 2.0, day_len + t_win / 2.0 + dt_win, dt_win)
        ns = dict(np=np, torch=torch, X_query_cart=torch.Tensor(x_query), tsteps_abs=tsteps_abs,
                  times_need=[tsteps[j:j + 1] for j in range(len(tsteps))], tree_tsteps=cKDTree(tsteps_abs.reshape(-1, 1)),
                  x_grid_ind_list=[0], use_updated_input=True, extract_input_from_data=pu.extract_input_from_data,
                  trv_pairwise=None, P=P, ind_use=ind_use, locs=net.sta, x_grids=[net.grid], A_src_in_sta_l=[A_src_in_sta.numpy()],
                  x_grids_trv=[trv_times], max_t=max_t, pred_params=[t_win, sig, t_win / 2.0, 25e3], dt_embed_discretize=dt_embed,
                  use_sign_input=False, device='cpu', use_phase_types=True, mz_list=[mz], ftrns1=identity, locs_use=net.sta[ind_use],
                  x_grids_cart_torch=[torch.Tensor(net.grid)],
                  tq=torch.arange(-t_win / 2.0, t_win / 2.0 + dt_win, dt_win).reshape(-1, 1).float(), t_win=t_win, dt_win=dt_win,
                  step_size=step_size, n_overlap=n_overlap, n_scale_x_grid=1)
        exec(compile(body, 'process_continuous_days.py:%d-%d' % (beg[0] + 1, end[0] + 1), 'exec'), ns)
        res['Out_2_' + step_size] = ns['Out_2']
        res['tsteps_' + step_size], res['tsteps_abs_' + step_size] = tsteps, tsteps_abs
        res['Out_2_sparse_' + step_size] = ns['Out_2_sparse']
        print('streaming', step_size, 'windows', len(tsteps), 'Out_2 max %.6f sum %.6f nnz(>0.01) %d' % (
            ns['Out_2'].max(), ns['Out_2'].sum(), len(ns['Out_2_sparse'])))
    res['t_win'], res['dt_win'], res['loop_lines'] = np.float64(6.0), np.float64(0.75), np.array([beg[0] + 1, end[0] + 1])
    np.savez_compressed(os.path.join(GOLD, 'streaming_10x100.npz'), **res)


def graphdd():
    """GraphDD's location network (Relocation/train_double_difference_model.py:333-536), the second consumer of the
    DataAggregation kernel family.  The file is a script with top-level code (it cannot be imported): the class definitions
    are read from it at generation time and executed as they are, with oracle/refshim behind `MessagePassing`."""
    work = tempfile.mkdtemp(prefix='genie_golden_')
    for f in ('config.yaml', 'train_config.yaml'):
        shutil.copy(os.path.join(REF, 'Code', f), work)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), work)
    from torch import nn
    from torch_geometric.nn import MessagePassing
    lines = open(os.path.join(REF, 'Relocation', 'train_double_difference_model.py')).read().split('\n')
    beg = [i for i, l in enumerate(lines) if l.startswith('class DataAggregation(MessagePassing)')]
    end = [i for i, l in enumerate(lines) if l.startswith('n_batch = ') and i > beg[0]]
    assert len(beg) == 1 and len(end) >= 1 and 300 < beg[0] < end[0] < 600
    ns = dict(torch=torch, nn=nn, np=np, MessagePassing=MessagePassing)
    exec(compile('\n'.join(lines[beg[0]:end[0]]), 'train_double_difference_model.py:%d-%d' % (beg[0] + 1, end[0]), 'exec'), ns)
    from genie_b200.process_utils import knn_graph, product_edge_lists
    rng = np.random.default_rng(21)
    for name, n_sta, n_src, use_memory in (('graphdd_12x9', 12, 9, False), ('graphdd_10x14_memory', 10, 14, True)):
        torch.manual_seed(n_sta)
        locs = np.stack((rng.uniform(0, 80e3, n_sta), rng.uniform(0, 80e3, n_sta), rng.uniform(0, 2e3, n_sta)), 1)
        srcs = np.stack((rng.uniform(0, 80e3, n_src), rng.uniform(0, 80e3, n_src), rng.uniform(-30e3, 0, n_src)), 1)
        A_sta, A_src = knn_graph((locs / 1000.0).astype(np.float32), 5), knn_graph((srcs / 1000.0).astype(np.float32), 4)
        A_in_pick, A_in_src, A_src_in_prod, A_src_in_sta = product_edge_lists(A_sta, A_src, n_sta, n_src)
        P = n_sta * n_src
        keep = torch.from_numpy(rng.random(P) < 0.8)                      # picks exist for a subset of (station, source) pairs
        new_id = torch.cumsum(keep.long(), 0) - 1

        def sub(A):
            ok = keep[A[0]] & keep[A[1]]
            return torch.stack((new_id[A[0][ok]], new_id[A[1][ok]]), 0).contiguous()
        A_in_pick, A_in_src = sub(A_in_pick), sub(A_in_src)
        A_src_in_sta = A_src_in_sta[:, keep].contiguous()
        n_prod = int(keep.sum())
        A_src_in_product = torch.stack((torch.arange(n_prod), A_src_in_sta[1]), 0)
        A_sta_in_product = torch.stack((torch.arange(n_prod), A_src_in_sta[0]), 0)
        m = ns['GNN_Location'](None, None, inpt_sources=True, use_sta_corr=True, use_memory=use_memory, use_mask=False,
                               use_aggregation=False, use_attention=False, device='cpu')
        m.eval()
        n_inpt, n_mask = 15 + 3, 15 + 3
        x = torch.from_numpy(rng.normal(size=(n_prod, n_inpt)).astype(np.float32))
        mask = torch.from_numpy((rng.random((n_prod, n_mask)) < 0.6).astype(np.float32))
        memory = torch.from_numpy(rng.normal(size=(n_src, 4)).astype(np.float32)) if use_memory else False
        lc, sc = torch.Tensor(locs), torch.Tensor(srcs)
        out = m(x, mask, A_in_pick, A_in_src, A_src_in_product, A_sta_in_product, A_src_in_sta, lc, sc, memory=memory)
        res = dict(locs=locs, srcs=srcs, x=x.numpy(), mask=mask.numpy(), A_in_pick=A_in_pick.numpy(), A_in_src=A_in_src.numpy(),
                   A_src_in_product=A_src_in_product.numpy(), A_sta_in_product=A_sta_in_product.numpy(),
                   A_src_in_sta=A_src_in_sta.numpy(), use_memory=np.int64(use_memory),
                   out0=out[0].numpy(), out1=out[1].numpy(), out2=out[2].numpy(), out3=out[3].numpy())
        if use_memory:
            res['memory'] = memory.numpy()
        res.update(_pack(m.state_dict()))
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **res)
        print(name, 'P=%d edges %d/%d keys %d' % (n_prod, A_in_pick.shape[1], A_in_src.shape[1], len(m.state_dict())),
              [float(np.abs(o.numpy()).sum()) for o in out])


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'synthetic'
    if mode == 'synthetic':
        synthetic()
    elif mode == 'synthetic_edges':
        synthetic(edges=True)
    elif mode == 'synthetic_abspos':
        synthetic(abs_pos=True)
    elif mode == 'legacy_input':
        legacy_input()
    elif mode == 'association':
        association()
    elif mode == 'association_edges':
        association(edges=True)
    elif mode == 'association_abspos':
        association(abs_pos=True)
    elif mode == 'subgraph':
        subgraph()
    elif mode == 'dense_adjacencies':
        dense_adjacencies()
    elif mode == 'input_variants':
        input_variants()
    elif mode == 'streaming':
        streaming()
    elif mode == 'graphdd':
        graphdd()
    elif mode == 'ferndale':
        from gen_golden_ferndale import ferndale
        ferndale()
    else:
        raise SystemExit('usage: gen_golden.py synthetic|ferndale')
