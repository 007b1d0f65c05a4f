"""Dev aid: does numbering the stations tile-major (every station tile = a contiguous range of station ids, i.e. contiguous
rows of every internal tensor) change the kernel times at C4?  Same network, same graph, stations relabelled."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import ops, synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan, bisection_groups, csr_by_destination, TILE_M
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian


def main():
    dev = torch.device('cuda:0')
    S, G = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    rp, cl = csr_by_destination(A_sta, S)
    ptr, nodes = bisection_groups(rp, cl, S, TILE_M)
    order = nodes.astype(np.int64)                       # new station s' = old station order[s']
    pos = np.empty(S, dtype=np.int64)
    pos[order] = np.arange(S)
    A_sta2 = torch.from_numpy(pos[A_sta.numpy()])         # the same graph, relabelled
    P = S * G
    g = torch.Generator(device=dev).manual_seed(1)
    Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
    posg = torch.from_numpy(net.grid).float().to(dev)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    packed = m._packed_weights(dev)
    capi.timing_enable(True)
    import genie_b200.plan as gp
    real = gp.bisection_groups
    for name, A in (('caller order', A_sta), ('tile-major order', A_sta2)):
        if A is A_sta2:                                  # the relabelled stations: the same groups, now ranges of ids
            gp.bisection_groups = lambda rp_, cl_, n, size: (ptr, np.arange(S, dtype=np.int32)) if (n == S and size == TILE_M) \
                else real(rp_, cl_, n, size)
        plan = GraphPlan.cartesian(A, A_src, S, G, device=dev)
        gp.bisection_groups = real
        rows, meta = plan.tiles['rows'].cpu().numpy(), plan.tiles['meta'].cpu().numpy()
        contig = all(np.array_equal(rows[t, :meta[t, 0]], np.arange(rows[t, 0], rows[t, 0] + meta[t, 0])) for t in range(len(meta)))
        for _ in range(2):
            ops.frontend_fwd(plan, packed, Slice, Mask, attr, posg, 30000.0)
        torch.cuda.synchronize()
        capi.timing_collect(reset=True)
        for _ in range(5):
            ops.frontend_fwd(plan, packed, Slice, Mask, attr, posg, 30000.0)
        torch.cuda.synchronize()
        kt = capi.timing_collect(reset=True)
        print('%s: tiles %d, own rows contiguous: %s' % (name, len(meta), contig))
        for k, (ms, n) in sorted(kt.items()):
            if n:
                print('    %-28s %8.3f ms' % (k, ms / n), flush=True)
        del plan
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
