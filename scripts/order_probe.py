"""Dev aid: time the DataAggregation kernels at C4 under different tile (grid-node) processing orders."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import ops, synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan, csr_by_destination, locality_order
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian


def morton(q):
    key = np.zeros(len(q), dtype=np.uint64)
    for b in range(12):
        for d in range(q.shape[1]):
            key |= ((q[:, d].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(q.shape[1] * b + d)
    return np.argsort(key, kind='stable').astype(np.int32)


def main():
    dev = torch.device('cuda:0')
    S, G = 1000, 50000
    net = synth.Network(S, G, seed=0)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    P = S * G
    g = torch.Generator(device=dev).manual_seed(1)
    Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
    Mask = (Slice.abs() > 0.01).float()
    torch.manual_seed(2)
    m = GCN_Detection_Network_extended(None, None, device=dev).eval()
    packed = m._packed_weights(dev)
    xyz = net.grid
    rp, col = csr_by_destination(A_src, G)
    orders = {
        'identity (synth anisotropic Morton)': np.arange(G, dtype=np.int32),
        'rcm': locality_order(rp, col, G),
        'morton3d isotropic 5km': morton(np.floor((xyz - xyz.min(0)) / 5000.0).astype(np.int64)),
        'morton3d isotropic 10km': morton(np.floor((xyz - xyz.min(0)) / 10000.0).astype(np.int64)),
        'morton2d xy 10km': morton(np.floor((xyz[:, :2] - xyz[:, :2].min(0)) / 10000.0).astype(np.int64)),
        'morton2d xy 20km': morton(np.floor((xyz[:, :2] - xyz[:, :2].min(0)) / 20000.0).astype(np.int64)),
        'stripes x(40km) then y': np.lexsort((xyz[:, 1], np.floor(xyz[:, 0] / 40000.0))).astype(np.int32),
        'random': np.random.default_rng(0).permutation(G).astype(np.int32),
    }
    src, dst = A_src[0].numpy(), A_src[1].numpy()
    capi.timing_enable(True)
    ref = None
    for name, order in orders.items():
        pos = np.empty(G, dtype=np.int64); pos[order] = np.arange(G)
        d = np.abs(pos[src] - pos[dst])
        plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev, grid_order=order)
        for _ in range(2):
            lat = ops.data_aggregation_fwd(plan, packed, Slice, Mask)
        torch.cuda.synchronize()
        capi.timing_collect(reset=True)
        for _ in range(3):
            lat = ops.data_aggregation_fwd(plan, packed, Slice, Mask)
        torch.cuda.synchronize()
        kt = capi.timing_collect(reset=True)
        if ref is None:
            ref = lat.clone()
        same = bool(torch.equal(ref, lat))
        print('%-38s d50 %6d d90 %6d d99 %6d | layer1_tc %.2f ms  layer2 %.2f ms  init %.2f ms | same result %s' % (
            name, np.median(d), np.percentile(d, 90), np.percentile(d, 99), kt['da_layer1_tc_kernel'][0] / 3,
            kt['da_layer2_readin_kernel'][0] / 3, kt['da_init_kernel'][0] / 3, same), flush=True)
        del plan


if __name__ == '__main__':
    main()
