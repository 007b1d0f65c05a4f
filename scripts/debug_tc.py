"""Dev aid: tensor-core layer-1 path vs the generic kernels on the same inputs (GPU only)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import ops, synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian

TC_OK_INDEX = None

def tc_ok_index():
    # layout.h: TC_BASE + TC_SCAL + TCS_OK ; derive from the packed size: PACKED = TC_BASE + TC_FLOATS, TC_SCAL = TC_FLOATS - 8
    n = int(capi.load().genie_frontend_packed_floats())
    return n - 8

def run(S, G, k_s, k_g, seed):
    dev = torch.device('cuda:0')
    net = synth.Network(S, G, seed=seed)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
    rng = np.random.default_rng(seed)
    P = S * G
    Slice = torch.from_numpy((rng.random((P, 4)) * (rng.random((P, 4)) < 0.35)).astype(np.float32)).to(dev)
    Mask = (Slice.abs() > 0.01).float()
    torch.manual_seed(2)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    packed = m._packed_weights(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    lat_tc = ops.data_aggregation_fwd(plan, packed, Slice, Mask)
    torch.cuda.synchronize()
    i = tc_ok_index()
    assert float(packed[i]) == 1.0, float(packed[i])
    packed[i] = 0.0
    lat_g = ops.data_aggregation_fwd(plan, packed, Slice, Mask)
    packed[i] = 1.0
    torch.cuda.synchronize()
    err = (lat_tc - lat_g).abs().max(dim=1)[0] / lat_g.abs().max()
    bad = torch.nonzero(err > 1e-5).flatten().cpu().numpy()
    print('S=%d G=%d: max rel err %.3e; rows > 1e-5: %d of %d' % (S, G, float(err.max()), len(bad), P))
    if len(bad):
        g, s = bad // S, bad % S
        print('  bad g: min %d max %d unique %d ; first 20 g:' % (g.min(), g.max(), len(np.unique(g))), np.unique(g)[:20])
        print('  bad s hist (bins of 16):', np.bincount(s // 16, minlength=(S + 15) // 16))
        ug, cnt = np.unique(g, return_counts=True)
        print('  bad rows per bad g: min %d max %d' % (cnt.min(), cnt.max()))
        print('  it index (g // 148) of bad g:', np.unique(ug // 148)[:40])
        print('  cta (g %% 148) of bad g:', np.unique(ug % 148)[:40])
        e2 = (lat_tc - lat_g).abs() / lat_g.abs().max()
        print('  bad channels hist:', (e2[bad] > 1e-5).sum(0).cpu().numpy())

if __name__ == '__main__':
    run(100, 500, 15, 15, 11)
    run(100, 5000, 15, 15, 31)
    run(1000, 1500, 15, 15, 21)


def ws_views(plan):
    P = plan.n_prod
    ws = plan.workspace()
    al = lambda n: (n + 255) // 256 * 256
    off = 0
    out = {}
    for name, w in (('tr0', 32), ('zc', 32), ('va', 16), ('vb', 16)):
        nbytes = P * w * 4
        out[name] = ws[off:off + nbytes].view(torch.float32).view(P, w)
        off += al(nbytes)
    return out


def run2(S, G, k_s, k_g, seed, reps=3):
    dev = torch.device('cuda:0')
    net = synth.Network(S, G, seed=seed)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, k_s, k_g)
    rng = np.random.default_rng(seed)
    P = S * G
    Slice = torch.from_numpy((rng.random((P, 4)) * (rng.random((P, 4)) < 0.35)).astype(np.float32)).to(dev)
    Mask = (Slice.abs() > 0.01).float()
    torch.manual_seed(2)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    packed = m._packed_weights(dev)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    i = tc_ok_index()
    packed[i] = 0.0
    ops.data_aggregation_fwd(plan, packed, Slice, Mask)
    torch.cuda.synchronize()
    ref = {k: v.clone() for k, v in ws_views(plan).items()}
    packed[i] = 1.0
    for rep in range(reps):
        ops.data_aggregation_fwd(plan, packed, Slice, Mask)
        torch.cuda.synchronize()
        got = ws_views(plan)
        for k in ('zc', 'va', 'vb'):
            e = (got[k] - ref[k]).abs() / ref[k].abs().max()
            bad = torch.nonzero(e.max(dim=1)[0] > 1e-5).flatten().cpu().numpy()
            print('rep %d %s: max rel err %.3e bad rows %d' % (rep, k, float(e.max()), len(bad)))
            if len(bad) and k == 'zc':
                g, s = bad // S, bad % S
                print('   g:', np.unique(g)[:10], ' s:', s[:40], ' cols bad:', (e[bad] > 1e-5).sum(0).cpu().numpy())
                srcdeg = plan.src_rowptr[1:] - plan.src_rowptr[:-1]
                print('   src deg of bad g:', srcdeg[np.unique(g)[:10]].cpu().numpy(), 'cta', np.unique(g)[:10] % 148, 'it', np.unique(g)[:10] // 148)
                print('   sample got', got[k][bad[0]].cpu().numpy()[:8], 'ref', ref[k][bad[0]].cpu().numpy()[:8])


if __name__ == '__main__':
    run2(100, 5000, 15, 15, 31, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
