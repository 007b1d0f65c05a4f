"""Dev aid: per-phase clock64() trace of CTA 0 of da_layer1_s_kernel (genie_debug_trace).  Slots per tile:
 0 mma: operands ready   1 mma: stage B issued   2 mma: epilogue B done   3 stage C issued   4 epilogue C done   5 stage D issued
 6 epi: stage B done     7 epi: B epilogue done  8 epi: stage C done      9 C epilogue done  10 stage D done     11 D epilogue done
12 gather: buffer full  13 converted            14 gathered              15 operand slot free 16 operands written
17 producer: buffer free 18 producer: copies issued"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genie_b200 import ops, synth, capi
from genie_b200.module import GCN_Detection_Network_extended
from genie_b200.plan import GraphPlan
from genie_b200.process_utils import extract_inputs_adjacencies_cartesian

dev = torch.device('cuda:0')
S, G = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
net = synth.Network(S, G, seed=0)
A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
if len(sys.argv) > 2:
    plan.set_storage(sys.argv[2])          # 'bf16': the fast storage mode
P = S * G
g = torch.Generator(device=dev).manual_seed(1)
Slice = torch.rand((P, 4), device=dev, generator=g) * (torch.rand((P, 4), device=dev, generator=g) < 0.3)
Mask = (Slice.abs() > 0.01).float()
attr = torch.rand((P, 3), device=dev, generator=g) - 0.5
pos = torch.from_numpy(net.grid).float().to(dev)
m = GCN_Detection_Network_extended(None, None, device=dev).eval()
packed = m._packed_weights(dev)
NTR = 160
trace = torch.zeros((NTR, 48), dtype=torch.int64, device=dev)
for _ in range(2):
    ops.frontend_fwd(plan, packed, Slice, Mask, attr, pos, 30000.0)
capi.check(capi.load().genie_debug_trace(ctypes.c_void_p(trace.data_ptr()), NTR))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.frontend_fwd(plan, packed, Slice, Mask, attr, pos, 30000.0)
e1.record()
torch.cuda.synchronize()
print('front end %.2f ms' % e0.elapsed_time(e1))
capi.check(capi.load().genie_debug_trace(None, 0))
tt = trace.cpu().numpy().astype(np.float64).reshape(NTR, 2, 24)
t = tt[:, 0, :]
n = min(NTR, (plan.tiles['n_tiles'] * G + 147) // 148 // 2 - int(os.environ.get('GENIE_TRACE_START', 0))) - 2
t = t[8:n]
print('tiles traced', len(t), 'cycles per tile (mma slot 0 to next slot 0): %.0f' % np.mean(np.diff(t[:, 0])))
def d(a, b, name):
    print('  %-52s %8.0f' % (name, np.mean(t[:, b] - t[:, a])))
d(0, 1, 'mma: issue stage B (39 MMAs)')
d(1, 6, 'stage B execution after issue (d_full seen by epilogue)')
d(6, 7, 'epilogue B (ld 64, prelu, split, st 128 + bias st 96)')
d(7, 2, 'mma wake-up after epilogue B')
d(2, 3, 'mma: issue stage C (24 MMAs)')
d(3, 8, 'stage C execution after issue')
d(8, 9, 'epilogue C (ld 96, st 128, store zc)')
d(9, 4, 'mma wake-up after epilogue C')
d(4, 5, 'mma: issue stage D (24 MMAs)')
d(5, 10, 'stage D execution after issue')
d(10, 11, 'epilogue D (ld 32, store va vb)')
d(12, 14, 'gather: 16 neighbour rows + unrotate')
d(14, 15, 'gather: wait for operand slot')
d(15, 16, 'gather: write STA operand to TMEM')
d(11, 19, 'epilogue: OWN / SRC operands of the next tile (incl. wait for buffer)')
d(17, 18, 'producer: issue cp.async + landing')
print('  %-52s %8.0f' % ('gather: wait for buffer full (after prev operands written)', np.mean(t[1:, 12] - t[:-1, 16])))
print('  %-52s %8.0f' % ('mma: wait for operands after prev stage D issue', np.mean(t[1:, 0] - t[:-1, 5])))
print('  %-52s %8.0f' % ('producer: fill latency (issue -> full seen by gather)', np.mean(t[:, 12] - t[:, 18])))

# timeline of a few tiles, both pipelines (cycles relative to the first event shown)
names = {17: 'P  buffer free', 18: 'P  buffer full', 12: 'G  gather start', 14: 'G  gather done', 16: 'G  STA written', 0: 'M  stage B issue',
         6: 'E  eB start', 7: 'E  eB done', 8: 'E  eC start', 9: 'E  eC done', 10: 'E  eD start', 11: 'E  eD done', 19: 'E  next OWN/SRC written'}
ev = []
for k in range(40, 43):
    for q in range(2):
        for s_, nm in names.items():
            if tt[k, q, s_] > 0:
                ev.append((tt[k, q, s_], 'tile %d pipe %d  %s' % (k, q, nm)))
ev.sort()
for tm_, nm in ev:
    print('%9.0f  %s' % (tm_ - ev[0][0], nm))
