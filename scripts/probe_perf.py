"""Quick timing probe of the front end at BASELINE configs (development aid; bench.py is the judged benchmark)."""
import sys
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import ops, synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian, InputExtractor


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def run(S, G, k_s, k_g, label):
    dev = torch.device('cuda:0')
    t0 = time.time()
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    P = S * G
    g = torch.Generator(device=dev).manual_seed(1)
    Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
    pos = torch.from_numpy(net.grid).float().to(dev)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    packed = m._packed_weights(dev)
    torch.cuda.synchronize()
    print('%s: S=%d G=%d P=%d set-up %.1fs workspace %.2f GB' % (label, S, G, P, time.time() - t0,
                                                                   plan.workspace_bytes / 1e9), flush=True)
    out = torch.empty((G, 30), device=dev)
    med, best = timed(lambda: ops.frontend_fwd(plan, packed, Slice, Mask, attr, pos, 30000.0, out=out))
    print('  front end (fused call): median %.3f ms best %.3f ms -> %.1f windows/s; 764 B/node => %.0f GB/s' % (
        med, best, 1e3 / med, 764.0 * P / med / 1e6), flush=True)
    capi.timing_enable(True)
    capi.timing_collect(reset=True)
    for _ in range(3):
        ops.frontend_fwd(plan, packed, Slice, Mask, attr, pos, 30000.0, out=out)
    torch.cuda.synchronize()
    for k, (ms, n) in sorted(capi.timing_collect(reset=True).items()):
        if n:
            print('    %-28s %8.3f ms x %d' % (k, ms / n, n // 3), flush=True)
    capi.timing_enable(False)
    # stage by stage
    med1, _ = timed(lambda: ops.data_aggregation_fwd(plan, packed, Slice, Mask))
    print('  data_aggregation_fwd (K1+K2+K3 w/ latent store): %.3f ms' % med1, flush=True)
    # input scatter
    max_t = net.max_moveout()
    trv = torch.from_numpy(net.travel_times()).to(dev) if P <= 2e7 else None
    if trv is None:
        chunks = [torch.from_numpy(net.travel_times(lo, min(G, lo + 2000))).to(dev) for lo in range(0, G, 2000)]
        trv = torch.cat(chunks, 0)
    ex = InputExtractor(plan, trv, np.arange(S), S, max_t, 3.0, 0.3)
    Pk = synth.make_picks(net, 0.0, max_t + 600.0, seed=1)
    ex.set_day(Pk)
    medi, _ = timed(lambda: ex(300.0))
    lo, hi = ex.window_rows(300.0)
    print('  input scatter: %.3f ms (%d picks in window, n_ts=%d)' % (medi, hi - lo, ex.params(300.0).n_ts), flush=True)
    Q = 10000
    rng = np.random.default_rng(0)
    xq = torch.from_numpy(np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q),
                                    rng.uniform(-40000, 0, Q)), 1)).float().to(dev)
    tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
    locs = torch.from_numpy(net.sta).float().to(dev)
    m.set_adjacencies_cartesian(A_sta, A_src, attr, S, G, device=dev)
    Sl, Mk = ex(300.0)
    medf, _ = timed(lambda: m.forward_fixed_source(Sl, Mk, None, None, None, locs, pos, xq, tq))
    print('  forward_fixed_source (front end + torch heads, Q=%d): %.3f ms' % (Q, medf), flush=True)
    with torch.no_grad():
        xs = out
        medh, _ = timed(lambda: (m.TemporalAttention(m.SpatialDirect(xs), tq),
                                 m.TemporalAttention(m.SpatialAttention(xs, xq, pos), tq)))
    print('  heads only (torch): %.3f ms' % medh, flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['c2', 'c4']
    if 'c2' in which:
        run(100, 5000, 15, 15, 'C2')
    if 'c4s' in which:
        run(1000, 5000, 15, 15, 'C4/10')
    if 'c4' in which:
        run(1000, 50000, 15, 15, 'C4')
