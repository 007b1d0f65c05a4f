"""Dev aid: data_aggregation_association / data_aggregation, fused per-node kernels vs torch ops, on random inputs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genie_b200.training as T
from genie_b200 import synth
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
dev = torch.device('cuda:0')
S, G = 18, 160
net = synth.Network(S, G, seed=1)
A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 8, 15)
plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
kg_sta, kg_src = T.build_kron_graphs(plan)
m = GCN_Detection_Network_extended(None, None, device=dev).train()
P = S * G
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
s_in, lat = rnd(P, 15).requires_grad_(True), rnd(P, 30)
m1, m2 = (rnd(P, 1) > 0).float(), (rnd(P, 4) > 0).float()
gy = rnd(P, 30)
res = {}
for tag, rows in (('torch', 10 ** 9), ('fused', 0)):
    T.MLP_MIN_ROWS = rows
    m.zero_grad(); s_in.grad = None
    inter = {}
    orig = T._agg_layer
    def agg(da, l_a, l_b, act, tr, msg_a, msg_b, kg1, kg2, ta, tb, _o=orig, _i=inter):
        for nm, x in (('tr', tr), ('msg_a', msg_a), ('msg_b', msg_b)):
            if x.requires_grad:
                x.retain_grad(); _i['%s_%d' % (nm, len(_i) // 3)] = x
        return _o(da, l_a, l_b, act, tr, msg_a, msg_b, kg1, kg2, ta, tb)
    T._agg_layer = agg
    out = T.data_aggregation_association(m.DataAggregationAssociationPhase, s_in, lat, m1, m2, kg_sta, kg_src)
    T._agg_layer = orig
    out.backward(gy)
    res[tag] = dict(out=out.detach().clone(), s=s_in.grad.clone(), **{k: v.grad.clone() for k, v in inter.items()},
                    **{k: p.grad.clone() for k, p in m.DataAggregationAssociationPhase.named_parameters() if p.grad is not None})
for k in res['torch']:
    a, b = res['fused'][k], res['torch'][k]
    sc = float(b.abs().max())
    print('%-28s rel %.3e |ref| %.3e' % (k, float((a - b).abs().max()) / max(sc, 1e-30), sc))
