"""Timing probe of forward_fixed (front end + heads + association branch) next to
forward_fixed_source at BASELINE configs; per-kernel device times from the library's own cudaEvents."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
from probe_perf import timed


def run(S, G, label, n_arv=600, n_src=2, Q=10000):
    dev = torch.device('cuda:0')
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    P = S * G
    g = torch.Generator(device=dev).manual_seed(1)
    Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
    pos = torch.from_numpy(net.grid).float().to(dev)
    locs = torch.from_numpy(net.sta).float().to(dev)
    tlatent = torch.cat([torch.from_numpy(net.travel_times(lo, min(G, lo + 2000))).to(dev) for lo in range(0, G, 2000)],
                        0).reshape(-1, 2).float()
    max_t = float(tlatent.max())
    dt_partition = torch.arange(-6.0, max_t + 6.6, 0.6, device=dev)
    l_dt = dt_partition.numel()
    rng = np.random.default_rng(3)
    sta_of = np.repeat(np.arange(S), l_dt * 10)
    A_p = torch.from_numpy(rng.integers(0, G, S * l_dt * 10) * S + sta_of).to(dev)
    A_s = torch.from_numpy(rng.integers(0, G, S * l_dt * 10) * S + sta_of).to(dev)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    m.set_adjacencies_cartesian(A_sta, A_src, attr, S, G, device=dev, A_edges_p=A_p, A_edges_s=A_s,
                                dt_partition=dt_partition, tlatent=tlatent)
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
    f_src = lambda: m.forward_fixed_source(Slice, Mask, None, None, None, locs, pos, xq, tq)
    f_fix = lambda: m.forward_fixed(Slice, Mask, tpick, ipick, phase, locs, pos, xq, x_src, tq, tqs, trv_q)
    med_s = 0.0 if os.environ.get('GENIE_PROBE_ASSOC_ONLY') == '1' else timed(f_src)[0]      # (ncu captures: forward_fixed first)
    med_f, _ = timed(f_fix)
    print('%s: S=%d G=%d P=%d picks=%d sources=%d: forward_fixed_source %.3f ms, forward_fixed %.3f ms (association adds %.3f ms)'
          % (label, S, G, P, n_arv, n_src, med_s, med_f, med_f - med_s), flush=True)
    capi.timing_enable(True)
    capi.timing_collect(reset=True)
    for _ in range(3):
        f_fix()
    torch.cuda.synchronize()
    for k, (ms, n) in sorted(capi.timing_collect(reset=True).items()):
        if n:
            print('    %-28s %8.3f ms' % (k, ms / n), flush=True)
    capi.timing_enable(False)
    # algorithmic bytes per product node of the association kernels: init 12+120+16+4 + 3*128,
    # layer 1 3*128+16 + 128+2*64, layer 2 128+2*64 + 128
    print('    algorithmic bytes/node: init 536, layer1 656, layer2 384 -> %.2f GB per call' % (1576.0 * P / 1e9), flush=True)


if __name__ == '__main__':
    which = sys.argv[1:] or ['c2', 'c4']
    if 'c2' in which:
        run(100, 5000, 'C2')
    if 'c4s' in which:
        run(1000, 5000, 'C4/10')
    if 'c4' in which:
        run(1000, 50000, 'C4')
