"""The caller-side streaming loop of `process_continuous_days.py:757-813` with everything resident on the device
(SURVEY.md §8f rank 4).

The reference, per origin-time sample: `extract_input_from_data` on the host, a fresh host->device copy of Slice / Mask
(:797), `forward_fixed_source`, a device->host copy of the query prediction and `Out_2[:, ip_need] += ...` in numpy (:802-805).
Here the day's pick table, the travel times and the solution array `Out_2` stay in HBM: per window one input-scatter launch pair
(genie_input_scatter_fwd), the front end + heads (genie_frontend_fwd, genie_heads_*) and one stacking launch
(genie_stack_output_fwd); the host only computes the window's row range in the pick table and the nine column indices.
One copy of `Out_2` [Q, n_steps] comes back at the end of the day.
"""
import ctypes

import numpy as np
import torch

from . import capi

N_OVERLAP = {'full': 1.0, 'partial': 3.0, 'half': 2.0}          # process_continuous_days.py:369-379


def nearest_index(sorted_vals, t):
    """Index of the entry of the ascending array `sorted_vals` nearest to every t (cKDTree.query on a 1-D set, :765, :793);
    the lower index on an exact tie."""
    sorted_vals = np.asarray(sorted_vals, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    hi = np.clip(np.searchsorted(sorted_vals, t, side='left'), 0, len(sorted_vals) - 1)
    lo = np.clip(hi - 1, 0, len(sorted_vals) - 1)
    return np.where(np.abs(sorted_vals[lo] - t) <= np.abs(sorted_vals[hi] - t), lo, hi)


class WindowRunner(object):
    """One window of the streaming loop (process_continuous_days.py:776-805: extract_input_from_data + forward_fixed_source) as
    ONE device-side unit: genie_window_fwd (a1 fused into the front end, Slice / Mask never reach HBM) + the read-out heads,
    captured once in a CUDA graph and replayed per window.  Everything that changes between windows is the 96-byte parameter
    block (capi.WindowParams: t0, the time axis, the window's row range in the resident pick table); the host writes it into
    a ring of pinned slots and one stream-ordered copy per window moves it to the device block the graph reads.

    mz: GCN_Detection_Network_extended with a CARTESIAN plan that has tiling tables; extractor: the plan's InputExtractor.
    source='resident': the picks come from the extractor's resident day table (`set_day`);
    source='staged':   the window's picks are copied per window from pinned host memory (`run(t0, picks_host=...)`) into a
                       device staging area of `max_window_picks` rows — the end-to-end path with host buffers.
    run() returns (y [G,T,1], x [Q,T,1]); with use_graph=True these are the graph's static output tensors (overwritten by the
    next run).  Call refresh() after changing model weights or query points."""

    RING = 64

    def __init__(self, mz, extractor, locs_use_cart, x_grid_cart, x_query_cart, t_query, use_graph=True, source='resident',
                 max_window_picks=None):
        from . import ops
        self.mz, self.ex, self.ops = mz, extractor, ops
        self.locs, self.grid, self.xq, self.tq = locs_use_cart, x_grid_cart, x_query_cart, t_query
        self.use_graph, self.source = bool(use_graph), source
        plan = mz._plan
        if plan is None or plan is not extractor.plan:
            raise capi.GenieError('WindowRunner: the model and the extractor must share one GraphPlan')
        # genie_window_fwd needs a dense plan with tiling tables; other plans (station in-degree > 16, sub-graph mode)
        # run the reference-shaped two-step sequence eagerly: extract_input -> Slice / Mask -> forward_fixed_source
        self.fused = plan.mode == capi.GRAPH_CARTESIAN and plan.tiles is not None and extractor.node_sta is None
        if not self.fused:
            self.use_graph = False
        if mz.use_absolute_pos and locs_use_cart is None:
            raise capi.GenieError('use_absolute_pos: WindowRunner needs locs_use_cart')
        self.dev = dev = plan.device
        self.sz = ctypes.sizeof(capi.WindowParams)
        self.wp_host = torch.zeros((self.RING, self.sz), dtype=torch.uint8).pin_memory()
        self.wp_dev = torch.zeros(self.sz, dtype=torch.uint8, device=dev)
        self._slot, self._ring_events = 0, [None] * self.RING
        prm = extractor.params(0.0)
        self.n_extra = int(prm.n_extra)
        self.n_ts_max = int(prm.n_ts) + 2                    # len(arange) can differ by one between windows (fp64 rounding)
        self.series = torch.empty((2 * extractor.n_sta_use * self.n_ts_max,), dtype=torch.float32, device=dev)
        if source == 'resident':
            if extractor._day is None:
                raise capi.GenieError("WindowRunner(source='resident') needs extractor.set_day(picks)")
            times = extractor._day[0]
            if max_window_picks is None:                      # the busiest window of the day, from the sorted pick times
                span = extractor.max_t + 4.0 * extractor.kernel_sig_t
                hi = np.searchsorted(times, times + span, side='right')
                max_window_picks = int((hi - np.arange(len(times))).max()) if len(times) else 0
            self.picks = extractor._day[1]
        elif source == 'staged':
            if not max_window_picks:
                raise capi.GenieError("WindowRunner(source='staged') needs max_window_picks")
            self.picks = torch.zeros((int(max_window_picks), 5), dtype=torch.float64, device=dev)
        else:
            raise ValueError(source)
        self.max_window_picks = int(max_window_picks)
        self._graph, self._out = None, None
        self.windows = 0

    def refresh(self):
        """Drop the captured graph (weights, query points or t_query changed)."""
        self._graph, self._out = None, None

    def _device_work(self):
        mz, ex, ops = self.mz, self.ex, self.ops
        init_relaid = mz._update_init_terms(self.locs, self.grid) if mz.use_absolute_pos else None
        packed = mz._packed_weights(self.dev, init_relaid)
        x_spatial = ops.window_fwd(mz._plan, packed, self.wp_dev, self.max_window_picks, self.n_extra, self.picks, ex.sta_perm,
                                   ex.ind_use, ex.trv_times, self.series, self.n_ts_max, mz._read_in_attr, self.grid,
                                   float(mz.scale_rel))[0]
        return mz._heads(x_spatial, self.grid, self.xq, self.tq)

    def _post_params(self, t0, lo, hi):
        """Window parameters -> next pinned ring slot -> (stream-ordered) device block."""
        wp = capi.WindowParams()
        wp.prm = self.ex.params(t0)
        if wp.prm.n_ts > self.n_ts_max or wp.prm.n_extra != self.n_extra:
            raise capi.GenieError('WindowRunner: the window time axis does not fit the series scratch')
        if hi - lo > self.max_window_picks:
            raise capi.GenieError('WindowRunner: %d picks in the window, max_window_picks = %d' % (hi - lo, self.max_window_picks))
        wp.pick_lo, wp.pick_hi = int(lo), int(hi)
        slot = self._slot
        ev = self._ring_events[slot]
        if ev is not None:
            ev.synchronize()                                  # the copy that last read this slot has executed
        ctypes.memmove(self.wp_host[slot].data_ptr(), ctypes.addressof(wp), self.sz)
        self.wp_dev.copy_(self.wp_host[slot], non_blocking=True)
        if self._ring_events[slot] is None:
            self._ring_events[slot] = torch.cuda.Event()
        self._ring_events[slot].record()
        self._slot = (slot + 1) % self.RING

    def run(self, t0, picks_host=None):
        t0 = float(t0)
        if not self.fused:
            picks = None
            if self.source != 'resident':
                n = int(picks_host.shape[0])
                self.picks[:n].copy_(picks_host, non_blocking=True)
                picks = self.picks[:n]
            self.windows += 1
            Slice, Mask = self.ex(t0, picks)
            return self.mz.forward_fixed_source(Slice, Mask, None, None, None, self.locs, self.grid, self.xq, self.tq)
        with torch.no_grad():
            if self.source == 'resident':
                lo, hi = self.ex.window_rows(t0)
            else:
                n = int(picks_host.shape[0])
                if n > self.max_window_picks:
                    raise capi.GenieError('WindowRunner: %d picks in the window, max_window_picks = %d' % (n, self.max_window_picks))
                if n:
                    self.picks[:n].copy_(picks_host, non_blocking=True)
                lo, hi = 0, n
            self._post_params(t0, lo, hi)
            self.windows += 1
            if not self.use_graph:
                return self._device_work()
            if self._graph is None:
                self._device_work()                           # warm-up: packs weights, builds caches, sets kernel attributes
                torch.cuda.current_stream(self.dev).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._out = self._device_work()
                self._graph = g
            self._graph.replay()
            return self._out


class DayProcessor(object):
    """`mz`: a GCN_Detection_Network_extended with its adjacencies set; `extractor`: an InputExtractor with the day's picks
    resident (`set_day`).  `run(tsteps, tsteps_abs)` returns Out_2 [Q, len(tsteps_abs)] on the device."""

    def __init__(self, mz, extractor, locs_use_cart, x_grid_cart, X_query_cart, t_win=6.0, dt_win=0.75, step_size='half',
                 n_scale_x_grid=1, pick_t_win=10.0, use_runner=True, use_graph=True):
        self.mz, self.ex = mz, extractor
        self.locs, self.grid, self.xq = locs_use_cart, x_grid_cart, X_query_cart
        self.t_win, self.dt_win, self.step_size = float(t_win), float(dt_win), step_size
        self.scale = 1.0 / (N_OVERLAP[step_size] * float(n_scale_x_grid))
        self.pick_t_win = float(pick_t_win)                       # extract_pick_inputs_from_data's t_win (process_utils.py:644)
        dev = X_query_cart.device
        self.t_rel = np.arange(-self.t_win / 2.0, self.t_win / 2.0 + self.dt_win, self.dt_win)       # :534, :793
        self.tq = torch.from_numpy(self.t_rel).reshape(-1, 1).float().to(dev)
        self.windows_done, self.windows_skipped = 0, 0
        self.use_runner, self.use_graph, self._run = use_runner, use_graph, None

    def _runner(self):
        """WindowRunner for dense plans with tiling tables; None -> the two-step path."""
        if self.use_runner is False:
            return None
        plan = self.mz._plan
        if plan is None or plan is not self.ex.plan or plan.mode != capi.GRAPH_CARTESIAN or plan.tiles is None or \
                self.ex.node_sta is not None or self.ex._day is None:
            return None
        if self._run is None:
            self._run = WindowRunner(self.mz, self.ex, self.locs, self.grid, self.xq, self.tq, use_graph=self.use_graph)
        return self._run

    def window_pick_count(self, t0):
        """len(lp_times[i0]) of the reference (:787): picks of used stations inside the input window (process_utils.py:476)
        and inside extract_pick_inputs_from_data's ball (process_utils.py:665)."""
        ex = self.ex
        times, cum = ex._day[0], ex.used_cum()
        sig2, mt = 2.0 * ex.kernel_sig_t, ex.max_t
        lo = max(np.searchsorted(times, t0 - sig2, side='right'), np.searchsorted(times, t0 - self.pick_t_win, side='left'))
        hi = min(np.searchsorted(times, t0 + mt + sig2, side='left'),
                 np.searchsorted(times, t0 + mt + self.pick_t_win, side='right'))
        return int(cum[hi] - cum[lo]) if hi > lo else 0

    def run(self, tsteps, tsteps_abs, out=None):
        """tsteps: origin-time samples to process (already thinned by the caller's min-pick rule, :725-748); tsteps_abs: the
        fixed solution grid (:411).  Windows without picks are skipped as in the reference (:787)."""
        dev = self.xq.device
        Q, n_abs = int(self.xq.shape[0]), len(tsteps_abs)
        tsteps_abs = np.asarray(tsteps_abs, dtype=np.float64)
        if out is None:
            out = torch.zeros((Q, n_abs), dtype=torch.float32, device=dev)
        T = len(self.t_rel)
        n_use = T - 1 if self.step_size == 'half' else T                                              # :802-805
        lib = capi.load()
        centre = nearest_index(tsteps_abs, np.asarray(tsteps, dtype=np.float64))                    # :765
        cols_all = nearest_index(tsteps_abs, tsteps_abs[centre][:, None] + self.t_rel[None, :])     # :793
        cols_dev = torch.from_numpy(cols_all.astype(np.int32)).to(dev)
        runner = self._runner()
        for i0, t0 in enumerate(np.asarray(tsteps, dtype=np.float64)):
            if self.window_pick_count(float(t0)) == 0:
                self.windows_skipped += 1
                continue
            if runner is not None:                 # a1 fused into the front end, the window replayed as one CUDA graph
                _, x = runner.run(float(t0))
            else:
                Slice, Mask = self.ex(float(t0))
                _, x = self.mz.forward_fixed_source(Slice, Mask, None, None, None, self.locs, self.grid, self.xq, self.tq)
            x = x.reshape(Q, T)
            with torch.cuda.device(dev):
                capi.check(lib.genie_stack_output_fwd(capi.dptr(x, torch.float32, 'x'), Q, T, n_use,
                                                      capi.dptr(cols_dev[i0], torch.int32, 'cols'),
                                                      ctypes.c_float(self.scale), capi.dptr(out, torch.float32, 'Out_2'),
                                                      n_abs, capi.stream_ptr(dev)))
            self.windows_done += 1
        return out

    def run_distributed(self, tsteps, tsteps_abs, group=None):
        """One day over the ranks of a torch.distributed group (one process per GPU): windows are independent, so rank r
        takes the origin times r, r + R, ... (no data-path exchange while processing), and the ranks' partial stacks are
        summed by ONE all-reduce at the end of the day (NCCL over NVLink on a multi-GPU box; Q x n_steps floats).  Every
        rank returns the full Out_2.  The sum order differs from the sequential loop's: equal within fp32 rounding."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        out = self.run(np.asarray(tsteps)[rank::world], tsteps_abs)
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def sparse(out, thresh=0.01):
        """Out_2_sparse of :812-813: rows (query index, time index, value) of the entries above the threshold."""
        iz = torch.nonzero(out > thresh, as_tuple=False)
        return torch.cat((iz.double(), out[iz[:, 0], iz[:, 1]].double().view(-1, 1)), dim=1)
