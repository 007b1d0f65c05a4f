"""bench.py — time-windows/sec of GENIE's product-graph front end on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

One step = one time window = input scatter (a1: picks -> Slice/Mask) + front end (a2-a4: DataAggregation ->
Bipartite_ReadIn -> SpatialAggregation x3) + read-out heads, i.e. one pass of the inference loop of the reference's
process_continuous_days.py:788-805 for one origin-time sample.  Workload: the synthetic 1000-station x 50000-grid-node
network of SURVEY.md §8d (dense product graph, P = 5e7 product nodes, k = 15/15), 24 h of synthetic picks.

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  value     windows/s with the day's picks, travel times and graphs resident in HBM (device-timed, max over ranks)
  e2e       the same metric through the public API with HOST buffers: per window the picks are copied from pinned host
            memory, and y [G,T,1] / x [Q,T,1] are copied back (the reference's loop does exactly this, :797-805)
  roofline  dominant kernel: algorithmic bytes per launch / its mean device time (cudaEvents recorded by the library on
            the launching stream inside the timed region), against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the oracle (CPU port of the reference algorithm) timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = 'time-windows/sec at 1k stations x 50k grid nodes'
UNIT = 'windows/s'
WORKLOADS = {
    # name: (stations, grid nodes, k_sta, k_grid)
    'c4_1000x50000_dense': (1000, 50000, 15, 15),
    'c2_100x5000_dense': (100, 5000, 15, 15),
    'c1_10x100_dense': (10, 100, 8, 15),
    # grid-sharded over the ranks (genie_b200/sharded.py): every window is computed cooperatively, needs --gpus >= 2
    'c5_2000x200000_sharded': (2000, 200000, 15, 15),
    'c5s_1000x20000_sharded': (1000, 20000, 15, 15),
}
KERNEL_SIG_T, DT, STEP_S, N_QUERY, SCALE_REL = 3.0, 0.3, 3.0, 10000, 30000.0
DAY_S = 86400.0
# algorithmic (compulsory) bytes per product node, fp32 intermediates — SURVEY.md §8d / DESIGN.md §4
BYTES_PER_NODE_WINDOW = 764.0
BYTES_PER_NODE_KERNEL = {'da_init_kernel': 136.0, 'da_layer1_kernel': 400.0, 'da_layer2_readin_kernel': 284.0,
                         'da_layer1_tc_kernel': 400.0, 'src_mean32_kernel': 256.0, 'da_layer1_s_kernel': 528.0,
                         'src_mean16_kernel': 128.0, 'da_layer2_s_kernel': 268.0}
# da_layer1_s_kernel: p 128 + msrc 128 + Mask 16 in, zc 128 + va 64 + vb 64 out; on the fused window path (genie_window_fwd) the mask
# rides in the p rows: 512.  da_layer2_s_kernel: zc 128 + va 64 + mean_src(vb) 64 + edge attr 12 (max(mask) rides in the zc rows).


def _peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def _traffic(kernel):
    """Per-launch DRAM bytes of `kernel` from the committed ncu --set full capture (profiles/roofline_traffic.json)."""
    p = os.path.join(REPO, 'profiles', 'roofline_traffic.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(kernel)
    return None


def _query_points(net, n, seed=3):
    rng = np.random.default_rng(seed)
    return np.stack((rng.uniform(0, net.width, n), rng.uniform(0, net.width, n), rng.uniform(-40000.0, 0.0, n)), axis=1)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md's clocks line)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[3:7]):
                if v == 'Active':
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---- the reference algorithm on the host cores (oracle port) ------------------------------------------------------------

def cpu_reference(workload, steps, warmup, sample_nodes=5.0e6):
    """Times oracle.input_scatter + oracle.forward_fixed_source (torch-CPU fp32, all host threads) on a bounded sample:
    all stations, the first Gs grid nodes (Morton order => a compact sub-volume) with their own kNN graph, Gs chosen so that
    the sample has about `sample_nodes` product nodes (P = 5e6 for the dense C4 / C5 sizes) — fewer when the
    host has little free memory (the oracle's message tensors take ~6 kB per product node).  Cost is linear in the number of
    product nodes, so windows/s at full size = sample rate x Gs/G, labelled extrapolated.  Median over the
    timed windows."""
    import torch
    from genie_b200 import synth
    from oracle import genie_oracle as go
    S, G, k_s, k_g = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        import psutil
        sample_nodes = min(sample_nodes, 0.5 * psutil.virtual_memory().available / 6000.0)
    except ImportError:
        pass
    Gs = int(max(20, min(G, sample_nodes // S)))
    net = synth.Network(S, G, seed=0)
    grid = net.grid[:Gs]
    A = go.build_adjacencies_dense(net.sta, grid, k_s, k_g)
    trv = net.travel_times(0, Gs)
    attr = torch.from_numpy(net.read_in_offsets(SCALE_REL, 0, Gs))
    max_t = net.max_moveout()
    n_win = steps + warmup
    picks = synth.make_picks(net, 0.0, n_win * STEP_S + max_t + 4 * KERNEL_SIG_T, seed=1)
    sd = go.init_state(seed=2)
    xq = torch.from_numpy(_query_points(net, min(N_QUERY, max(100, N_QUERY * Gs // G)))).float()
    tq = torch.arange(-3.0, 3.01, 0.75).reshape(-1, 1)
    gridt = torch.from_numpy(grid).float()
    ind_use = np.arange(S)
    times = []
    with torch.no_grad():
        for i in range(n_win):
            t0 = i * STEP_S
            t_a = time.perf_counter()
            lo = np.searchsorted(picks[:, 0], t0 - 2 * KERNEL_SIG_T)
            hi = np.searchsorted(picks[:, 0], t0 + max_t + 2 * KERNEL_SIG_T, side='right')
            Sl, Mk = go.input_scatter(picks[lo:hi], t0, ind_use, S, A[5].numpy(), trv, max_t, KERNEL_SIG_T, DT)
            y, x = go.forward_fixed_source(sd, torch.from_numpy(Sl), torch.from_numpy(Mk), A[2], A[3], attr, A[4], A[1],
                                           gridt, xq, tq, SCALE_REL, 3.0 * KERNEL_SIG_T)
            t_b = time.perf_counter()
            if i >= warmup:
                times.append(t_b - t_a)
    med = float(np.median(times))
    value = (1.0 / med) * Gs / G
    sample = ('oracle (torch-CPU fp32 port of module.py:999-1020 + process_utils.py:460-629), %d host threads, on all %d '
              'stations x the first %d of %d grid nodes (P=%d), %d windows after %d warm-up, median %.3f s/window on the '
              'sample; %s' % (cores, S, Gs, G, S * Gs, len(times), warmup, med,
                              'value = sample windows/s x %d/%d (cost linear in P: EXTRAPOLATED to the full size)' % (Gs, G)
                              if Gs < G else 'full size, not extrapolated'))
    return dict(value=value, unit=UNIT, cores=cores, kind='port', sample=sample, sample_product_nodes=S * Gs,
                windows_timed=len(times), extrapolated=bool(Gs < G)), med


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    S, G, k_s, k_g = WORKLOADS[args.workload]
    # P = 5e6 per step when the whole run stays within a few minutes (~10 s per step at that size on 16 cores), else smaller
    n_win = max(args.steps + args.warmup, 1)
    cb, s_per_step = cpu_reference(args.workload, args.steps, args.warmup, sample_nodes=min(5.0e6, 1.25e8 / n_win))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / cb['value'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'stations': S, 'grid_nodes': G, 'product_nodes': S * G, 'k_sta': k_s,
                   'k_grid': k_g, 'parallelism': 'host threads (torch intra-op), rank 0 only'},
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---- this repo's CUDA path ---------------------------------------------------------------------------------------------

class Workload(object):
    """Device-resident state of one replica: network, graphs, travel times, a day of picks, model."""

    def __init__(self, name, dev, day_s=DAY_S):
        import torch
        from genie_b200 import synth
        from genie_b200.module import GCN_Detection_Network_extended
        from genie_b200.process_utils import InputExtractor, extract_inputs_adjacencies_cartesian
        S, G, k_s, k_g = WORKLOADS[name]
        self.S, self.G, self.P, self.dev = S, G, S * G, dev
        net = synth.Network(S, G, seed=0)
        self.net = net
        A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
        self.A_sta, self.A_src = A_sta, A_src
        sta_d = torch.from_numpy(net.sta).to(dev)
        grid_d = torch.from_numpy(net.grid).to(dev)
        trv = torch.empty((G, S, 2), dtype=torch.float32, device=dev)
        attr = torch.empty((G * S, 3), dtype=torch.float32, device=dev)
        for lo in range(0, G, 4096):                       # synth.Network.travel_times / read_in_offsets, on the device
            hi = min(G, lo + 4096)
            diff = grid_d[lo:hi, None, :] - sta_d[None, :, :]
            d = diff.norm(dim=2)
            trv[lo:hi, :, 0] = (d / synth.VP).float()
            trv[lo:hi, :, 1] = (d / synth.VS).float()
            attr[lo * S:hi * S] = (diff / SCALE_REL).reshape(-1, 3).float()
        torch.manual_seed(2)
        m = GCN_Detection_Network_extended(None, None, scale_rel=SCALE_REL, device=dev).eval()
        m.set_adjacencies_cartesian(A_sta, A_src, attr, S, G, device=dev)
        self.model = m
        self.max_t = net.max_moveout()
        self.ex = InputExtractor(m._plan, trv, np.arange(S), S, self.max_t, KERNEL_SIG_T, DT)
        self.picks = synth.make_picks(net, 0.0, day_s, seed=1)
        self.ex.set_day(self.picks)
        self.locs = sta_d.float()
        self.grid = grid_d.float().contiguous()
        self.xq = torch.from_numpy(_query_points(net, N_QUERY)).float().to(dev)
        self.tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
        self.n_windows = int((day_s - self.max_t) // STEP_S)
        # host-side staging for the end-to-end leg
        self.picks_host = torch.from_numpy(self.ex._day[1].cpu().numpy()).pin_memory()
        # two sets of pinned result buffers: the host takes window w - 1 while the device works on window w
        self.y_host = [torch.empty((G, self.tq.shape[0], 1), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.x_host = [torch.empty((N_QUERY, self.tq.shape[0], 1), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._e2e_ev, self._e2e_n = [None, None], 0

    def runners(self, use_graph):
        """The streaming fast path (genie_b200.streaming.WindowRunner): a1 fused into the front end + heads, one parameter
        block per window; `use_graph` replays the whole window as one CUDA graph (launch-bound sizes)."""
        from genie_b200.streaming import WindowRunner
        self.runner = WindowRunner(self.model, self.ex, self.locs, self.grid, self.xq, self.tq, use_graph=use_graph)
        self.runner_e2e = WindowRunner(self.model, self.ex, self.locs, self.grid, self.xq, self.tq, use_graph=use_graph,
                                       source='staged', max_window_picks=self.runner.max_window_picks)
        self.use_graph = self.runner.use_graph

    def window_resident(self, w):
        """Hot path with everything resident in HBM."""
        return self.runner.run(w * STEP_S)

    def window_e2e(self, w):
        """Public API with host buffers: H2D of the window's picks, D2H of y and x (process_continuous_days.py:797-805)."""
        import torch
        lo, hi = self.ex.window_rows(w * STEP_S)
        y, x = self.runner_e2e.run(w * STEP_S, self.picks_host[lo:hi])
        # as a streaming caller does it (process_continuous_days.py:797-805 consumes window w while w + 1 is computed): this
        # window's results are copied to one of two pinned buffer sets in stream order, and the host waits for the PREVIOUS
        # window's copy — so the launches of the next window are queued while the device is still busy with this one
        b = self._e2e_n & 1
        self._e2e_n += 1
        self.y_host[b].copy_(y, non_blocking=True)
        self.x_host[b].copy_(x, non_blocking=True)
        if self._e2e_ev[b] is None:
            self._e2e_ev[b] = torch.cuda.Event()
        self._e2e_ev[b].record()
        if self._e2e_ev[b ^ 1] is not None:
            self._e2e_ev[b ^ 1].synchronize()
        return (hi - lo) * 5 * 8 + self.runner_e2e.sz, (y.numel() + x.numel()) * 4

    def window_two_step(self, w):
        """The reference-shaped call sequence (extract_input -> Slice / Mask tensors -> forward_fixed_source)."""
        Slice, Mask = self.ex(w * STEP_S)
        return self.model.forward_fixed_source(Slice, Mask, None, None, None, self.locs, self.grid, self.xq, self.tq)


class ShardedWorkload(object):
    """One network sharded by grid nodes over all ranks (BASELINE.json configs[4]); same interface as Workload."""

    def __init__(self, name, dev, rank, world, day_s=DAY_S, storage='fp32', exchange='peer'):
        import torch
        from genie_b200 import synth
        from genie_b200.module import GCN_Detection_Network_extended
        from genie_b200.process_utils import InputExtractor, extract_inputs_adjacencies_cartesian
        from genie_b200.sharded import CudaBackend, GridPartition, PeerHalo, ShardedFrontEnd
        S, G, k_s, k_g = WORKLOADS[name]
        self.S, self.G, self.dev, self.rank, self.world = S, G, dev, rank, world
        net = synth.Network(S, G, seed=0)
        A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
        part = GridPartition(A_src, G, world)
        nodes = part.local_nodes(rank)
        self.n_local, self.n_owned = len(nodes), len(part.owned[rank])
        self.P = S * G                                            # whole-job product nodes
        sta_d = torch.from_numpy(net.sta).to(dev)
        grid_d = torch.from_numpy(net.grid).to(dev)
        loc_d = grid_d[torch.from_numpy(nodes).to(dev)]
        trv = torch.empty((self.n_local, S, 2), dtype=torch.float32, device=dev)
        attr = torch.empty((self.n_local * S, 3), dtype=torch.float32, device=dev)
        for lo in range(0, self.n_local, 4096):
            hi = min(self.n_local, lo + 4096)
            diff = loc_d[lo:hi, None, :] - sta_d[None, :, :]
            d = diff.norm(dim=2)
            trv[lo:hi, :, 0] = (d / synth.VP).float()
            trv[lo:hi, :, 1] = (d / synth.VS).float()
            attr[lo * S:hi * S] = (diff / SCALE_REL).reshape(-1, 3).float()
        torch.manual_seed(2)
        self.model = GCN_Detection_Network_extended(None, None, scale_rel=SCALE_REL, device=dev).eval()
        be = CudaBackend(self.model, A_sta, part.local_graph(rank), S, self.n_local, self.n_owned, attr, A_src, G, dev,
                         grid_groups=part.local_groups(rank))
        if storage == 'bf16':
            self.model._plan = be.plan
            self.model.set_storage('bf16')
        # halo rows: stored into the peers' memory by the layer-1 kernel itself (PeerHalo), or one all_to_all_single per window
        self.peer_halo = PeerHalo.create(part, rank, be.plan, S, dev) if (exchange == 'peer' and world > 1) else None
        self.fe = ShardedFrontEnd(part, rank, be, dev, peer_halo=self.peer_halo)
        self.halo_rows = [len(h) for h in part.halo]
        self.exchange_events = None
        self.max_t = net.max_moveout()
        self.ex = InputExtractor(be.plan, trv, np.arange(S), S, self.max_t, KERNEL_SIG_T, DT)
        self.picks = synth.make_picks(net, 0.0, day_s, seed=1)
        self.ex.set_day(self.picks)
        self.grid = grid_d.float().contiguous()
        self.xq = torch.from_numpy(_query_points(net, N_QUERY)).float().to(dev)
        self.tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
        self.n_windows = int((day_s - self.max_t) // STEP_S)
        self.picks_host = torch.from_numpy(self.ex._day[1].cpu().numpy()).pin_memory()
        self.y_host = torch.empty((G, self.tq.shape[0], 1), dtype=torch.float32).pin_memory()
        self.x_host = torch.empty((N_QUERY, self.tq.shape[0], 1), dtype=torch.float32).pin_memory()

    def _window(self, Slice, Mask):
        import torch
        from genie_b200.sharded import sharded_heads
        m = self.model
        with torch.no_grad():
            x_spatial, _ = self.fe.forward(Slice, Mask, self.grid, SCALE_REL, events=self.exchange_events)
            # the read-out heads are per grid node / query point: every rank computes its block of rows, two all-gathers
            y, x = sharded_heads(m, x_spatial, self.grid, self.xq, self.tq, self.rank, self.world)
            return (y, x) if self.rank == 0 else (None, None)

    def window_resident(self, w):
        Slice, Mask = self.ex(w * STEP_S)
        return self._window(Slice, Mask)

    def window_e2e(self, w):
        import torch
        lo, hi = self.ex.window_rows(w * STEP_S)
        picks = self.picks_host[lo:hi].to(self.dev, non_blocking=True)
        Slice, Mask = self.ex(w * STEP_S, picks)
        y, x = self._window(Slice, Mask)
        d2h = 0
        if y is not None:
            self.y_host.copy_(y, non_blocking=True)
            self.x_host.copy_(x, non_blocking=True)
            d2h = (y.numel() + x.numel()) * 4
        torch.cuda.current_stream().synchronize()
        return (hi - lo) * 5 * 8, d2h

    def use_collective_exchange(self):
        """Switch the halo rows back to the NCCL all-to-all (the comparison leg)."""
        if self.peer_halo is not None:
            self.fe.peer_halo = None
            self.fe.backend.plan.set_halo_export(None, None, None, None, None)

    def close(self):
        if self.peer_halo is not None:
            self.peer_halo.close()
            self.peer_halo = None


def closure_parity(wl, w, n_clusters=5, cluster=4, tol=1e-4):
    """Oracle check of one of the timed windows at FULL size (outside every timed region): the window's a1 inputs, x_latent
    and Bipartite_ReadIn rows of >= 16 sampled grid nodes against the CPU oracle run on their 2-hop source-graph closure
    (oracle/closure_check.py), and y / x of the timed call against the oracle's SpatialAggregation + heads on the full grid.
    Element-wise (row-scaled) relative metric; the integer time-bin map must be equal."""
    import torch
    from oracle import closure_check as cc
    m, ex, S, G = wl.model, wl.ex, wl.S, wl.G
    t0 = w * STEP_S
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    targets = cc.sample_targets(G, n_clusters, cluster)
    lo, hi = ex.window_rows(t0)
    picks = ex._day[1][lo:hi].cpu().numpy()
    trv_of = lambda nodes: ex.trv_times[torch.from_numpy(nodes).to(wl.dev)].cpu().numpy()
    attr_of = lambda nodes: m._read_in_attr.view(G, S, 3)[torch.from_numpy(nodes).to(wl.dev)].reshape(-1, 3).cpu().numpy()
    t_a = time.time()
    want = cc.oracle_on_closure(sd, wl.A_sta, wl.A_src, S, G, targets, picks, t0, trv_of, attr_of, wl.max_t, KERNEL_SIG_T, DT)
    from genie_b200 import ops
    rows = torch.from_numpy(want['nodes']).to(wl.dev)
    y, x = wl.runner.run(t0)                                             # the timed call (leaves the window's parameter block)
    y, x = y.clone(), x.clone()
    r = wl.runner
    # the same fused kernels once more, this time materialising their inputs and intermediates
    if not r.fused:          # plans without tiling tables (C1): the runner is the two-step sequence itself
        Slice, Mask = ex(t0)
        _, latent, readin = m.front_end(Slice, Mask, wl.grid, want_latent=True, want_readin=True, locs_use_cart=wl.locs)
    else:
        _, latent, readin, Slice, Mask = ops.window_fwd(m._plan, m._packed_weights(wl.dev), r.wp_dev, r.max_window_picks, r.n_extra,
            r.picks, ex.sta_perm, ex.ind_use, ex.trv_times, r.series, r.n_ts_max, m._read_in_attr, wl.grid,
            float(m.scale_rel), want_inputs=True, want_latent=True, want_readin=True)
    S2, M2, tb = ex(t0, want_time_bin=True)                              # the stand-alone a1 kernels (integer time bins)
    y2, x2 = m.forward_fixed_source(S2, M2, None, None, None, wl.locs, wl.grid, wl.xq, wl.tq)
    fused_same = bool(torch.equal(S2, Slice) and torch.equal(M2, Mask) and torch.equal(y2, y) and torch.equal(x2, x))
    del S2, M2
    rep = cc.compare(want, tb[rows].cpu().numpy(), Slice[rows].cpu().numpy(), Mask[rows].cpu().numpy(),
                     latent[rows].cpu().numpy(), readin[torch.from_numpy(targets).to(wl.dev)].cpu().numpy())
    del latent, tb
    y_o, x_o = cc.oracle_tail(sd, readin.cpu(), wl.A_src, wl.grid.cpu(), wl.xq.cpu(), wl.tq.cpu(), SCALE_REL,
                              float(m.TemporalAttention.scale_t))
    rep['y_rel'], rep['x_rel'] = cc.global_rel(y.cpu().numpy(), y_o), cc.global_rel(x.cpu().numpy(), x_o)
    rep['window'], rep['picks_in_window'], rep['seconds'] = int(w), int(hi - lo), round(time.time() - t_a, 1)
    rep['tolerance'] = tol
    rep['fused_equals_two_step'] = fused_same      # genie_window_fwd vs extract_input + forward_fixed_source, bit for bit
    rep['ok'] = bool(rep.get('time_bin_equal') and rep.get('mask_equal') and rep['max_rel'] < tol and
                     rep['y_rel'] < tol and rep['x_rel'] < tol and fused_same)
    return rep


def bf16_mode(wl, windows, W, K, use_graph, args):
    """The SECOND mode (BASELINE.json configs[1]: bf16 inference): GENIE_STORAGE_BF16 keeps the gathered intermediate rows as
    bf16 (fp32 arithmetic).  Timed like `value` (same windows, resident inputs); its error against the oracle (sampled
    closure) and against the fp32 mode is reported beside it.  It does not satisfy, and is never used for, the 1e-4 parity
    bar."""
    import torch
    from genie_b200 import capi
    m = wl.model
    y32, x32 = wl.runner.run(windows[W] * STEP_S)
    y32, x32 = y32.clone(), x32.clone()
    m.set_storage('bf16')
    wl.runners(use_graph)
    for w in windows[:W]:
        wl.window_resident(w)
    torch.cuda.synchronize()
    capi.timing_enable(True)
    capi.timing_collect(reset=True)
    profile = os.environ.get('GENIE_BENCH_PROFILE') == 'bf16'      # ncu --profile-from-start off: this mode's timed loop only
    if profile:
        torch.cuda.profiler.start()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg.record()
    for w in windows[W:W + K]:
        wl.window_resident(w)
    end.record()
    torch.cuda.synchronize()
    if profile:
        torch.cuda.profiler.stop()
    ms = beg.elapsed_time(end)
    kt = capi.timing_collect(reset=True)
    capi.timing_enable(False)
    y16, x16 = wl.runner.run(windows[W] * STEP_S)
    from oracle import closure_check as cc
    out = {'value': K / (ms * 1e-3), 'ms_per_step': ms / K, 'unit': UNIT,
           'algorithmic_bytes_per_node': 404.0,
           'window_roofline_frac': 404.0 * wl.P / (ms / K * 1e-3) / 1e9 / _peaks()[0],
           'kernels_ms_per_step': {k: round(v[0] / K, 4) for k, v in sorted(kt.items()) if v[1] > 0},
           'y_rel_vs_fp32': cc.global_rel(y16.cpu().numpy(), y32.cpu().numpy()),
           'x_rel_vs_fp32': cc.global_rel(x16.cpu().numpy(), x32.cpu().numpy())}
    if not args.no_parity_check:
        rep = closure_parity(wl, windows[W], tol=2e-2)
        out['max_rel_vs_oracle'] = max(rep['max_rel'], rep['y_rel'], rep['x_rel'])
        out['parity_check'] = rep
    m.set_storage('fp32')
    wl.runners(use_graph)
    return out


def sharded_leg(args, dev, rank, world, dist, name='c5_2000x200000_sharded'):
    """One network too large for one GPU (C5: 2000 stations x 200000 grid nodes, P = 4e8) sharded by grid nodes over all
    ranks (genie_b200/sharded.py): every rank works on the SAME window, one halo exchange (all-to-all of the layer-2 message
    rows, NCCL over NVLink) + one all-gather of read-in rows per window.  Timed on the device, max over ranks."""
    import torch
    from genie_b200 import capi
    S, G, k_s, k_g = WORKLOADS[name]
    t_a = time.time()
    wl = ShardedWorkload(name, dev, rank, world, day_s=min(args.day_seconds, 3600.0), storage=args.sharded_storage,
                         exchange=args.sharded_exchange)
    torch.cuda.synchronize()
    setup = time.time() - t_a
    K, W = max(2, min(args.steps, 6)), 3
    peak = _peaks()[0]

    def timed():
        for w in range(W):
            wl.window_resident(w)
        dist.barrier()
        torch.cuda.synchronize()
        wl.exchange_events = []
        capi.timing_enable(True)
        capi.timing_collect(reset=True)
        beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        beg.record()
        for w in range(W, W + K):
            wl.window_resident(w)
        end.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = beg.elapsed_time(end)
        kt = capi.timing_collect(reset=True)
        capi.timing_enable(False)
        ex_ms = sum(a.elapsed_time(b) for a, b in wl.exchange_events)
        wl.exchange_events = None
        kern_ms = sum(v[0] for v in kt.values())
        l1_ms = kt.get('da_layer1_s_kernel', (0.0, 0))[0]
        t = torch.tensor([ms, ex_ms, kern_ms, float(wl.fe.exchange_bytes), float(wl.n_local), float(wl.n_owned), l1_ms],
                         dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # per-rank picture: where the ranks differ (ms per step: wait at the exchange / fence, then the kernels by name)
        names = ['da_init_kernel', 'src_mean32_kernel', 'da_layer1_s_kernel', 'src_mean16_kernel', 'da_layer2_s_kernel',
                 'input_gather_kernel', 'sa_main_kernel', 'heads_grid_kernel', 'heads_query_kernel']
        mine = torch.tensor([ex_ms / K] + [kt.get(n, (0.0, 0))[0] / K for n in names], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {'columns': ['exchange_wait'] + names, 'ms_per_step': [[round(float(v), 3) for v in r] for r in allr]}
        ms = float(tmax[0])
        return {'value': K / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / K, 'per_rank': per_rank,
                'exchange_bytes_per_step_sum': float(t[3]), 'exchange_ms_per_step_max': float(tmax[1]) / K,
                'exchange_share_of_step': float(tmax[1]) / ms, 'library_kernels_share_of_step_rank_max': float(tmax[2]) / ms,
                'layer1_station_pass_ms_per_step_max': float(tmax[6]) / K,
                'product_nodes_per_s': S * G * K / (ms * 1e-3),
                'window_roofline_frac': BYTES_PER_NODE_WINDOW * S * G / (ms / K * 1e-3) / 1e9 / (peak * world)}, tmax

    main, tmax = timed()
    peer = wl.peer_halo is not None
    rep = {'workload': name, 'stations': S, 'grid_nodes': G, 'product_nodes': S * G, 'n_gpus': world, 'steps': K, 'warmup': W,
           'storage': args.sharded_storage, 'scaling': 'strong',
           'halo_grid_nodes_per_rank': wl.halo_rows, 'owned_grid_nodes_max': int(tmax[5]), 'local_grid_nodes_max': int(tmax[4]),
           'exchange': ('peer stores: the layer-1 station pass writes the v_b rows of boundary grid nodes straight into the peers\' '
                        'landing buffers over NVLink (genie_plan_set_halo_export); exchange_ms = the one-element all-reduce that '
                        'orders layer 2 after every rank\'s layer 1') if peer else
                       'all_to_all_single (v_b halo rows), NCCL',
           'collectives': ('all_reduce of one element (fence) + ' if peer else 'all_to_all_single (v_b halo rows) + ') +
                          'all_gather_into_tensor (read-in rows) per window, NCCL',
           'setup_s': round(setup, 1)}
    rep.update(main)
    if peer:
        wl.use_collective_exchange()                       # the same windows with the NCCL all-to-all, for comparison
        rep['with_nccl_all_to_all'] = timed()[0]
    wl.close()
    del wl
    torch.cuda.empty_cache()
    return rep


def run_genie(args):
    import torch
    import torch.distributed as dist
    from genie_b200 import capi
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device — genie_b200 has no CPU path (use --impl reference for the host baseline)')
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    S, G, k_s, k_g = WORKLOADS[args.workload]
    t_setup = time.time()
    sharded = args.workload.endswith('_sharded')
    if sharded:
        if world < 2:
            raise RuntimeError('bench.py: %s is the grid-sharded workload, launch it with --gpus >= 2' % args.workload)
        wl = ShardedWorkload(args.workload, dev, rank, world, day_s=args.day_seconds, storage=args.sharded_storage,
                             exchange=args.sharded_exchange)
        # every rank works on the SAME window (its shard of the grid): fixed total work, strong scaling
        windows = [i % wl.n_windows for i in range(2 * (args.steps + args.warmup))]
        units = 1
    else:
        wl = Workload(args.workload, dev, day_s=args.day_seconds)
        # whole-window CUDA graphs where launches dominate (C1 / C2); at C4 the kernels are > 99 % of a step and the library's
        # per-kernel events (skipped under capture) are wanted inside the timed region
        use_graph = args.graph == 'on' or (args.graph == 'auto' and S * G < 5000000)
        wl.runners(use_graph)
        # rank r streams windows r, r + world, ...: windows are independent, no data-path collective
        windows = [(rank + i * world) % wl.n_windows for i in range(2 * (args.steps + args.warmup))]
        units = world
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: resident inputs (value) -------------------------------------------------------------------------------
    for w in windows[:W]:
        wl.window_resident(w)
    barrier()
    capi.timing_enable(True)
    capi.timing_collect(reset=True)
    n0 = capi.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    profile = os.environ.get('GENIE_BENCH_PROFILE') == '1'     # ncu --profile-from-start off: capture the timed loop only
    if profile:
        torch.cuda.profiler.start()
    # Small networks (C1 / C2) fit in the 126 MB L2: every window is then timed on its own, with a 256 MB write in between
    # (outside the timed spans) so that no window starts with its tables cached by the previous one.
    flush = (not sharded) and wl.P * 1400.0 < 2.5e8
    if flush:
        scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        spans = []
        for w in windows[W:W + K]:
            scratch.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            wl.window_resident(w)
            b.record()
            spans.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in spans)
    else:
        beg.record()
        for w in windows[W:W + K]:
            wl.window_resident(w)
        end.record()
        barrier()
        ms = beg.elapsed_time(end)
    if profile:
        torch.cuda.profiler.stop()
    launches = capi.launch_count() - n0
    kt = capi.timing_collect(reset=True)
    capi.timing_enable(False)
    graph = (not sharded) and wl.use_graph
    if graph:
        # graph replays bypass the library's launch counter and per-kernel events: count the launches of one eager window and
        # take the per-kernel times from K eager windows OUTSIDE the timed region (reported as such)
        from genie_b200.streaming import WindowRunner
        eager = WindowRunner(wl.model, wl.ex, wl.locs, wl.grid, wl.xq, wl.tq, use_graph=False)
        eager.run(windows[0] * STEP_S)
        n0 = capi.launch_count()
        eager.run(windows[0] * STEP_S)
        launches = (capi.launch_count() - n0) * K
        capi.timing_enable(True)
        capi.timing_collect(reset=True)
        for w in windows[W:W + K]:
            eager.run(w * STEP_S)
        torch.cuda.synchronize()
        kt = capi.timing_collect(reset=True)
        capi.timing_enable(False)
    # ---- leg 2: end to end with host buffers (e2e) ----------------------------------------------------------------------
    for w in windows[W + K:W + K + W]:
        wl.window_e2e(w)
    barrier()
    beg2, end2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg2.record()
    h2d = d2h = 0
    for w in windows[W + K + W:W + K + W + K]:
        a, b = wl.window_e2e(w)
        h2d += a
        d2h += b
    end2.record()
    barrier()
    ms2 = beg2.elapsed_time(end2)
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        t = torch.tensor([ms, ms2], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms2 = float(t[0]), float(t[1])
        cnt = torch.tensor([launches, h2d, d2h], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches = int(cnt[0])
        if sharded:
            h2d, d2h = int(cnt[1]), int(cnt[2])
    if rank == 0:
        peak, peak_src = _peaks()
        timed = {k: v for k, v in kt.items() if v[1] > 0}
        dom = max(timed, key=lambda k: timed[k][0])
        dom_ms = timed[dom][0] / timed[dom][1]
        launch_nodes = wl.n_owned * wl.S if sharded else wl.P          # product nodes one launch (rank 0) processes
        per_node = BYTES_PER_NODE_KERNEL.get(dom, BYTES_PER_NODE_WINDOW)
        if dom == 'da_layer1_s_kernel' and not sharded and getattr(wl.runner, 'fused', False):
            per_node = 512.0
        dom_bytes = per_node * launch_nodes
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        kernel_total = sum(v[0] for v in timed.values())
        line = {
            'metric': METRIC if not sharded else 'time-windows/sec, %d stations x %d grid nodes sharded over the GPUs' % (S, G),
            'value': units * K / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong' if sharded else 'weak',
            'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': args.workload, 'stations': S, 'grid_nodes': G, 'product_nodes': S * G, 'k_sta': k_s,
                       'k_grid': k_g, 'kernel_sig_t': KERNEL_SIG_T, 'dt': DT, 'window_step_s': STEP_S,
                       'n_query': N_QUERY, 'n_t_query': 9, 'picks_resident': int(wl.picks.shape[0]),
                       'parallelism': ('grid nodes sharded over %d ranks (halo rows per rank %s): one all-to-all of '
                                       'layer-2 message rows + one all-gather of read-in rows per window' % (
                                           world, wl.halo_rows)) if sharded else
                       'windows round-robin over %d replica(s), no data-path collective' % world,
                       'l2_policy': ('working set below 2 x L2: every window timed on its own, a 256 MB write between windows '
                                     '(outside the timed spans) flushes the 126 MB L2') if flush else
                       'inputs larger than L2: every window streams %.1f GB of node features through HBM (L2 = 126 MB), no '
                       'explicit flush' % (BYTES_PER_NODE_WINDOW * wl.P / 1e9),
                       'setup_s': round(t_setup, 1),
                       'cuda_graph': bool(graph), 'kernel_times_from': 'eager windows outside the timed region (graph replays '
                       'carry no events)' if graph else 'library events inside the timed region',
                       'api': 'streaming.WindowRunner: genie_window_fwd (a1 fused into the front end) + genie_heads_*'
                       if not sharded else 'sharded.ShardedFrontEnd'},
            'e2e': {'value': units * K / (ms2 * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d // K,
                    'd2h_bytes_per_step': d2h // K, 'ms_per_step': ms2 / K,
                    'host_sync': 'every step copies its picks from pinned host memory and its y, x back to pinned host memory; '
                                 'the host waits for window w - 1 while window w runs (two pinned result sets)'
                                 if not sharded else 'every step: picks H2D, y, x D2H, stream synchronize'},
            'gpu_launches': launches,
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak,
                         # the committed ncu capture is of the C4 workload on one GPU
                         'traffic': _traffic(dom) if (args.workload == 'c4_1000x50000_dense' and not sharded) else None,
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': dom_bytes, 'ms_per_launch': dom_ms,
                         'kernel_share_of_step': timed[dom][0] / ms,
                         'window': {'algorithmic_bytes': BYTES_PER_NODE_WINDOW * wl.P,
                                    'achieved': BYTES_PER_NODE_WINDOW * wl.P / (ms / K * 1e-3) / 1e9,
                                    'frac': BYTES_PER_NODE_WINDOW * wl.P / (ms / K * 1e-3) / 1e9 / (peak * (world if sharded else 1))},
                         'kernels_ms_per_step': {k: round(v[0] / K, 4) for k, v in sorted(timed.items())},
                         'library_kernels_share_of_step': kernel_total / ms},
            'clocks': clocks,
        }
        if not sharded and not args.no_parity_check:
            line['parity_check'] = closure_parity(wl, windows[W])          # the first timed window, against the CPU oracle
        if not sharded and not args.no_bf16 and wl.runner.fused:
            line['modes'] = {'fp32_parity': {'value': line['value'], 'ms_per_step': line['ms_per_step'],
                                             'max_rel_vs_oracle': (line.get('parity_check') or {}).get('max_rel')},
                             'bf16_storage': bf16_mode(wl, windows, W, K, use_graph, args)}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference(args.workload, 5, 1, sample_nodes=2.0e6)[0]   # ~25 s of CPU work
    if world > 1 and not sharded and not args.no_sharded_leg:
        # BASELINE.json configs[4]: the C5 network sharded by grid nodes over the same ranks, after the replica legs
        del wl
        torch.cuda.empty_cache()
        rep = sharded_leg(args, dev, rank, world, dist)
        if rank == 0:
            line['sharded'] = rep
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='genie', choices=['genie', 'reference'])
    ap.add_argument('--workload', default='c4_1000x50000_dense', choices=sorted(WORKLOADS))
    ap.add_argument('--day-seconds', type=float, default=DAY_S)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity-check', action='store_true')
    ap.add_argument('--no-bf16', action='store_true', help='skip the bf16-storage second mode')
    ap.add_argument('--no-sharded-leg', action='store_true', help='N > 1: skip the grid-sharded C5 leg after the replica legs')
    ap.add_argument('--sharded-exchange', default='peer', choices=['peer', 'nccl'],
                    help='halo rows of the sharded leg: peer stores from inside the layer-1 kernel, or an NCCL all-to-all')
    ap.add_argument('--sharded-storage', default='fp32', choices=['fp32', 'bf16'])
    ap.add_argument('--graph', default='auto', choices=['auto', 'on', 'off'], help='replay each window as one CUDA graph')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'genie' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_genie(args)


if __name__ == '__main__':
    main()
