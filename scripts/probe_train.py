"""Timing probe of one training sample (forward with gradients + backward + Adam step) at BASELINE.json configs[2]'s
network (100 x 5000, k = 15 / 15), with the share of the product-graph message passing kernel (genie_kron_spmm_fwd)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian, product_edge_lists
from probe_perf import timed


class Data(object):
    def __init__(self, x, edge_index):
        self.x, self.edge_index = x, edge_index


def run(S, G, label, n_arv=400, n_src=8, Q=2500, batch=32):
    dev = torch.device('cuda:0')
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15, device=dev)
    A_ps, A_pg, A_sip, A_sis = (a.to(dev) for a in product_edge_lists(A_sta.cpu(), A_src.cpu(), S, G))
    P = S * G
    g = torch.Generator(device=dev).manual_seed(1)
    Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
    pos = torch.from_numpy(net.grid).float().to(dev)
    locs = torch.from_numpy(net.sta).float().to(dev)
    tlatent = torch.from_numpy(net.travel_times()).to(dev).reshape(-1, 2).float()
    max_t = float(tlatent.max())
    dt_partition = torch.arange(-6.0, max_t + 6.6, 0.6, device=dev)
    l_dt = dt_partition.numel()
    rng = np.random.default_rng(3)
    sta_of = np.repeat(np.arange(S), l_dt * 10)
    A_p = torch.from_numpy(rng.integers(0, G, S * l_dt * 10) * S + sta_of).to(dev)
    A_s = torch.from_numpy(rng.integers(0, G, S * l_dt * 10) * S + sta_of).to(dev)
    m = GCN_Detection_Network_extended(None, None, device=dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    xq = torch.from_numpy(np.stack((rng.uniform(0, net.width, Q), rng.uniform(0, net.width, Q),
                                    rng.uniform(-40000, 0, Q)), 1)).float().to(dev)
    tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
    tpick = torch.from_numpy(rng.uniform(0.0, max_t, n_arv)).float().to(dev)
    ipick = torch.from_numpy(rng.integers(0, S, n_arv)).long().to(dev)
    phase = torch.from_numpy(rng.integers(0, 2, n_arv)).long().reshape(-1, 1).to(dev)
    isrc = rng.choice(G, n_src, replace=False)
    x_src = pos[isrc]
    trv_q = tlatent.view(G, S, 2)[isrc]
    tqs = torch.zeros(n_src, device=dev)
    graphs = (A_ps, A_pg, Data(attr, A_sip), Data(attr, A_sip.flip(0).contiguous()), A_sis, A_src, A_p, A_s, dt_partition, tlatent)
    window = (tpick, ipick, phase, locs, pos, xq, x_src, tq, tqs, trv_q)
    lbl = [torch.rand((G, 9), device=dev), torch.rand((Q, 9), device=dev), torch.rand((n_src, n_arv), device=dev),
           torch.rand((n_src, n_arv), device=dev)]
    mse = torch.nn.MSELoss()

    def sample():
        out = m(Slice, Mask, *graphs, *window)
        loss = sum(w * mse(o[:, :, 0], l) for w, o, l in zip((0.1, 0.4, 0.25, 0.25), out, lbl)) / batch
        loss.backward()

    def step():                      # train_GENIE_model.py:1756-1830: n_batch samples accumulate, then one optimiser step
        opt.zero_grad()
        for _ in range(batch):
            sample()
        opt.step()
    med, best = timed(sample, n=10, warm=3)
    print('%s: S=%d G=%d P=%d picks=%d sources=%d: one sample forward+backward %.2f ms (best %.2f)' % (label, S, G, P, n_arv,
                                                                                                       n_src, med, best), flush=True)
    t0 = time.time()
    meds, _ = timed(step, n=3, warm=1)
    print('  training step (batch %d + Adam): %.1f ms -> %.1f windows/s' % (batch, meds, 1e3 * batch / meds), flush=True)
    capi.timing_enable(True)
    capi.timing_collect(reset=True)
    sample()
    torch.cuda.synchronize()
    for k, (ms, n) in sorted(capi.timing_collect(reset=True).items()):
        if n:
            print('    %-24s %8.3f ms total in %d launches per sample' % (k, ms, n), flush=True)
    capi.timing_enable(False)
    if '--profile' in sys.argv:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                sample()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=28, max_name_column_width=60), flush=True)
    with torch.no_grad():
        m.eval()
        medi, _ = timed(lambda: m(Slice, Mask, *graphs, *window), n=10, warm=3)
    print('  the same call under torch.no_grad() (fused inference kernels): %.2f ms' % medi, flush=True)


if __name__ == '__main__':
    run(100, 5000, 'C3 (configs[2])')
