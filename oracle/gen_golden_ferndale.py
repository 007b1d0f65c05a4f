"""Ferndale known-answer fixture (TEST INFRASTRUCTURE; build container only) — see oracle/gen_golden.py.

Runs the reference's inference set-up and one window of its hot loop (process_continuous_days.py:190-340, 627-634,
776-797) on `Examples/Ferndale.zip`: 92 stations (77 active on 2022-12-20), source grid 0 (150 nodes), the trained GNN
checkpoint and the physics-informed travel-time network, at origin time t0 = 38 940 s (the busiest minute of that day,
SURVEY.md §8c).  The current reference code is used with the example's YAML files; the checkpoint is loaded with
strict=False because only `SpatialAttention.f_queries.*` (added after the example was trained) is missing.
"""
import os
import shutil
import tempfile
import zipfile

import numpy as np

from gen_golden import REF, GOLD, _import_reference, _run_reference_window, _pack


def ferndale(t0=38940.0, n_query=400):
    work = tempfile.mkdtemp(prefix='genie_ferndale_')
    with zipfile.ZipFile(os.path.join(REF, 'Examples', 'Ferndale.zip')) as z:
        z.extractall(work)
    root = os.path.join(work, 'Ferndale') + '/'
    for f in ('module.py', 'utils.py', 'process_utils.py'):      # stale copies shipped inside the example: not used
        if os.path.exists(root + f):
            os.remove(root + f)
    torch, module, pu, Data = _import_reference(os.path.join(REF, 'Code'), root)
    import utils as ru
    import yaml
    config = yaml.safe_load(open(root + 'config.yaml'))
    name = config['name_of_project']

    z = np.load(root + '%s_region.npz' % name)
    lat_range, lon_range, depth_range, deg_pad = z['lat_range'], z['lon_range'], z['depth_range'], z['deg_pad']
    z = np.load(root + 'Grids/%s_seismic_network_templates_ver_1.npz' % name)
    x_grids = z['x_grids']
    z = np.load(root + '%s_stations.npz' % name)
    locs, mn, rbest = z['locs'], z['mn'], z['rbest']
    z = np.load(root + 'GNN_TrainedModels/%s_trained_gnn_model_step_20000_ver_1_losses.npz' % name)
    graph_params, pred_params = z['graph_params'], z['pred_params']
    k_sta, k_spc = int(graph_params[0]), int(graph_params[1])
    sig = float(pred_params[1])

    lat_e = [lat_range[0] - deg_pad, lat_range[1] + deg_pad]
    lon_e = [lon_range[0] - deg_pad, lon_range[1] + deg_pad]
    scale_x_extend = np.array([lat_e[1] - lat_e[0], lon_e[1] - lon_e[0], depth_range[1] - depth_range[0]]).reshape(1, -1)
    rbest_t, mn_t = torch.Tensor(rbest), torch.Tensor(mn)

    def ftrns1(x):
        return (rbest @ (ru.lla2ecef(x) - mn).T).T

    def ftrns1_diff(x):
        return (rbest_t @ (ru.lla2ecef_diff(x, device='cpu') - mn_t).T).T

    def ftrns2_diff(x):
        return ru.ecef2lla_diff((rbest_t.T @ x.T).T + mn_t, device='cpu')

    trv = ru.load_travel_time_neural_network(root, ftrns1_diff, ftrns2_diff, 1,
                                             use_physics_informed=config['use_physics_informed'], device='cpu')
    P, ind_use = ru.load_picks(root, [2022, 12, 20], spr_picks=1, n_ver=1)
    x_grids_trv = ru.compute_travel_times(trv, locs, x_grids, device='cpu')
    max_t = float(np.ceil(max([x.max() for x in x_grids_trv])))
    grid = x_grids[0]
    trv_times = x_grids_trv[0]
    dt = float(np.round(sig / 10.0, 2))

    mz = module.GCN_Detection_Network_extended(ftrns1_diff, ftrns2_diff, device='cpu')
    ck = torch.load(root + 'GNN_TrainedModels/%s_trained_gnn_model_step_20000_ver_1.h5' % name, map_location='cpu')
    torch.manual_seed(0)
    missing = mz.load_state_dict(ck, strict=False)
    print('missing', missing.missing_keys, 'unexpected', missing.unexpected_keys)
    mz.eval()

    rng = np.random.default_rng(7)
    X_query = np.stack((rng.uniform(lat_range[0], lat_range[1], n_query), rng.uniform(lon_range[0], lon_range[1], n_query),
                        rng.uniform(depth_range[0], depth_range[1], n_query)), axis=1)
    t_win = float(pred_params[0])
    dt_win = 1.0 if t_win == 10.0 else t_win / 8.0
    t_query = np.arange(-t_win / 2.0, t_win / 2.0 + dt_win, dt_win)

    # the model sees Cartesian coordinates: hand the reference already-projected positions and identity transforms
    # for the graph builder (extract_inputs_adjacencies applies ftrns1 itself), exactly as the script does.
    # _run_reference_window works on geographic `locs`/`grid` through `identity`=ftrns1
    res = _run_reference_window(torch, module, pu, Data, mz, locs, ind_use, grid, trv_times, P, t0, max_t, sig, dt,
                                k_sta, k_spc, scale_x_extend, ftrns1(X_query), t_query, ftrns1)
    keep = (P[:, 0] > t0 - 3.0 * sig) & (P[:, 0] < t0 + max_t + 3.0 * sig)
    res.update(_pack(mz.state_dict()))
    res.update(sta=ftrns1(locs), grid=ftrns1(grid), ind_use=ind_use, trv_times=trv_times, picks=P[keep],
               t0=np.float64(t0), max_t=np.float64(max_t), kernel_sig_t=np.float64(sig), dt=np.float64(dt),
               k_sta=np.int64(k_sta), k_spc=np.int64(k_spc), scale_rel=np.float64(mz.scale_rel),
               scale_t=np.float64(module.scale_t), x_query=ftrns1(X_query), t_query=t_query,
               attr_scale=scale_x_extend)
    np.savez_compressed(os.path.join(GOLD, 'ferndale_t38940.npz'), **res)
    print('ferndale P=%d picks=%d sig=%.2f dt=%.2f max_t=%.1f k=%d/%d scale_rel=%.0f scale_t=%.2f' % (
        res['Slice'].shape[0], keep.sum(), sig, dt, max_t, k_sta, k_spc, mz.scale_rel, module.scale_t))
    for k in ('Slice', 'Mask', 'x_latent', 'read_in', 'sa1', 'sa2', 'x_spatial'):
        print('  %-10s sum %.6f abs-sum %.6f max %.6f' % (k, res[k].sum(), np.abs(res[k]).sum(), res[k].max()))
    print('  Slice nnz %d nonzero rows %d  y.max %.6f x.max %.6f' % (
        (res['Slice'] != 0).sum(), (np.abs(res['Slice']).sum(1) > 0).sum(), res['y'].max(), res['x'].max()))
    shutil.rmtree(work, ignore_errors=True)
