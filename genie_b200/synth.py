"""Seeded synthetic seismic networks and pick streams (SURVEY.md §8d "Synthetic inputs").

Everything is generated directly in Cartesian metres — the GNN only ever sees `locs_use_cart` / `x_temp_cuda_cart`
(module.py:908) — with a homogeneous-velocity travel-time table, Poisson events and uniform false picks at roughly
the Ferndale example's pick rate.  Host-side numpy only; used by tests, `bench.py` and `__graft_entry__.smoke()`.
"""
import numpy as np

VP, VS = 6000.0, 3464.0          # m/s
WIDTH_KM = {10: 60.0, 100: 200.0, 1000: 650.0, 2000: 900.0}


def _morton_order(xyz, bits=10):
    """Sort key along a 3-D Morton (Z-order) curve so that consecutive grid nodes are spatial neighbours."""
    lo, hi = xyz.min(0), xyz.max(0)
    q = np.minimum(((xyz - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    key = np.zeros(len(xyz), dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return np.argsort(key, kind='stable')


class Network(object):
    """Stations, source grid, travel times and normalised station-grid offsets of one synthetic network."""

    def __init__(self, n_sta, n_grid, seed=0, width_km=None, depth_km=40.0, morton=True):
        rng = np.random.default_rng(seed)
        w = 1000.0 * (width_km if width_km is not None else WIDTH_KM.get(n_sta, 20.0 * np.sqrt(n_sta)))
        self.width = w
        sta = np.stack((rng.uniform(0, w, n_sta), rng.uniform(0, w, n_sta), rng.uniform(0, 2000.0, n_sta)), axis=1)
        grid = np.stack((rng.uniform(0, w, n_grid), rng.uniform(0, w, n_grid),
                         rng.uniform(-1000.0 * depth_km, 0.0, n_grid)), axis=1)
        if morton:
            sta = sta[_morton_order(np.stack((sta[:, 0], sta[:, 1], np.zeros(n_sta)), axis=1))]
            grid = grid[_morton_order(grid)]
        self.sta = sta.astype(np.float64)            # [S,3] m
        self.grid = grid.astype(np.float64)          # [G,3] m
        self.S, self.G = n_sta, n_grid

    def travel_times(self, lo=0, hi=None):
        """fp32 [g, S, 2] P and S travel times (s) for grid nodes lo..hi."""
        hi = self.G if hi is None else hi
        d = np.linalg.norm(self.grid[lo:hi, None, :] - self.sta[None, :, :], axis=2)
        return np.stack((d / VP, d / VS), axis=2).astype(np.float32)

    def max_moveout(self):
        diag = np.sqrt(2 * self.width ** 2 + 42000.0 ** 2)
        return float(np.ceil(diag / VS / 10.0) * 10.0)

    def read_in_offsets(self, scale, lo=0, hi=None):
        """fp32 [(hi-lo)*S, 3]: (grid - station)/scale, the `A_src_in_edges.x` of process_continuous_days.py:630."""
        hi = self.G if hi is None else hi
        off = (self.grid[lo:hi, None, :] - self.sta[None, :, :]) / scale
        return off.reshape(-1, 3).astype(np.float32)


def make_picks(net, t_start, t_end, seed=1, events_per_3h=50.0, false_per_sta_min=1.0, max_dist_km=150.0,
               noise_frac=0.025):
    """Picks [n,5] float64 (time, station, amp, prob, phase) sorted by time, train_config.yaml:26,43 style."""
    rng = np.random.default_rng(seed)
    dur = t_end - t_start
    n_ev = rng.poisson(events_per_3h * dur / 10800.0)
    rows = []
    for _ in range(int(n_ev)):
        src = np.array([rng.uniform(0, net.width), rng.uniform(0, net.width), rng.uniform(-40000.0, 0.0)])
        t_org = rng.uniform(t_start, t_end)
        d = np.linalg.norm(net.sta - src[None, :], axis=1)
        near = np.where(d < 1000.0 * max_dist_km)[0]
        for ph, v in ((0, VP), (1, VS)):
            use = near[rng.uniform(size=len(near)) < 0.8]
            tt = d[use] / v
            t = t_org + tt + rng.normal(0.0, 1.0, len(use)) * noise_frac * tt
            rows.append(np.stack((t, use.astype(np.float64), np.ones(len(use)), np.ones(len(use)),
                                  np.full(len(use), float(ph))), axis=1))
    n_false = rng.poisson(false_per_sta_min * net.S * dur / 60.0)
    rows.append(np.stack((rng.uniform(t_start, t_end, n_false), rng.integers(0, net.S, n_false).astype(np.float64),
                          np.ones(n_false), np.ones(n_false), rng.integers(0, 2, n_false).astype(np.float64)), axis=1))
    P = np.concatenate(rows, axis=0)
    P = P[(P[:, 0] >= t_start) & (P[:, 0] < t_end)]
    return P[np.argsort(P[:, 0], kind='stable')]
