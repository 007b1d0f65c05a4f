"""Dev aid: per-parameter gradient of one training sample, fused per-node layer kernels vs the torch-op path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import genie_b200.training as training
from conftest import load_golden
from test_gpu_parity import _assoc_setup
dev = torch.device('cuda:0')
name = 'assoc_18of20x160'
d, sd = load_golden(name)
m, graphs, window, locs, grid = _assoc_setup(d, sd, dev, name)
m.train()
t = lambda k: torch.from_numpy(d[k]).to(dev)
rng = np.random.default_rng(5)
lbl = [torch.from_numpy(rng.uniform(0, 1, d[k].shape[:2]).astype(np.float32)).to(dev) for k in ('y', 'x', 'arv_p', 'arv_s')]
mse = torch.nn.MSELoss()
res = {}
for tag, rows in (('torch', 10 ** 9), ('fused', 0)):
    training.MLP_MIN_ROWS = rows
    m.zero_grad()
    out = m(t('Slice'), t('Mask'), *graphs, *window)
    loss = sum(w * mse(o[:, :, 0], l) for w, o, l in zip((0.1, 0.4, 0.25, 0.25), out, lbl))
    loss.backward()
    res[tag] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    print(tag, 'loss', float(loss))
for k in res['torch']:
    a, b = res['fused'][k], res['torch'][k]
    sc = float(b.abs().max())
    if sc > 0:
        e = float((a - b).abs().max()) / sc
        if e > 1e-4:
            print('%-60s rel %.3e  |g| %.3e' % (k, e, sc))
