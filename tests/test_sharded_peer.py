"""Grid sharding with the halo rows travelling over peer memory (genie_b200.sharded.PeerHalo): two ranks = two processes sharing
cuda:0 (gloo rendezvous on 127.0.0.1), landing buffers mapped across the processes with the library's inter-process handles.
`ShardedFrontEnd.forward` must reproduce the unsharded front end on both ranks, window after window.  Needs a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err, REPO

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import torch.distributed as dist
    from genie_b200 import ops, synth
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.plan import GraphPlan
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    from genie_b200.sharded import CudaBackend, GridPartition, PeerHalo, ShardedFrontEnd, sharded_heads
    from oracle import genie_oracle as go
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cuda:0')
    S, G = 150, 800
    net = synth.Network(S, G, seed=5)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 15, 15)
    m = GCN_Detection_Network_extended(None, None, device=dev)
    m.load_state_dict(go.init_state(seed=7), strict=False)
    m.eval()
    grid = torch.from_numpy(net.grid).float().to(dev)
    g = torch.Generator().manual_seed(3)
    attr = torch.rand((S * G, 3), generator=g) - 0.5
    part = GridPartition(A_src, G, world)
    nd = torch.from_numpy(part.local_nodes(rank))
    loc = lambda x: x.view(G, S, -1).index_select(0, nd).reshape(len(nd) * S, -1).contiguous().to(dev)
    be = CudaBackend(m, A_sta, part.local_graph(rank), S, len(nd), len(part.owned[rank]), loc(attr), A_src, G, dev,
                     grid_groups=part.local_groups(rank))
    halo = PeerHalo(part, rank, be.plan, S, dev)
    fe = ShardedFrontEnd(part, rank, be, dev, peer_halo=halo)
    plan = GraphPlan.cartesian(A_sta, A_src, S, G, device=dev)
    errs = []
    for w in range(3):                                                      # several windows: buffer re-use across windows
        Slice = torch.rand((S * G, 4), generator=g) * (torch.rand((S * G, 4), generator=g) < 0.3)
        Mask = (Slice.abs() > 0.01).float()
        xs, r = fe.forward(loc(Slice), loc(Mask), grid, 30000.0)
        want_xs, _, want_r = ops.frontend_fwd(plan, m._packed_weights(dev), Slice.to(dev), Mask.to(dev), attr.to(dev), grid,
                                              30000.0, want_readin=True)
        errs.append((rel_err(r.cpu().numpy(), want_r.cpu().numpy()), rel_err(xs.cpu().numpy(), want_xs.cpu().numpy())))
    # the read-out heads split by rows over the ranks against the single-device heads (ragged blocks: 800 / 2, 77 / 2 rows)
    xq = torch.rand((77, 3), generator=g).to(dev) * torch.tensor([net.width, net.width, -40000.0], device=dev)
    tq = torch.arange(-3.0, 3.01, 0.75, device=dev).reshape(-1, 1)
    with torch.no_grad():
        y_s, x_s = sharded_heads(m, xs, grid, xq, tq, rank, world)
        y_1, x_1 = m._heads(xs, grid, xq, tq)
    errs.append((rel_err(y_s.cpu().numpy(), y_1.cpu().numpy()), rel_err(x_s.cpu().numpy(), x_1.cpu().numpy())))
    assert y_s.shape == y_1.shape and x_s.shape == x_1.shape
    ret[rank] = (errs, fe.exchange_bytes, int(halo.exp_row.numel()))
    halo.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_front_end_two_processes_peer_stores():
    if not torch.cuda.is_available():
        pytest.fail('needs a CUDA device')
    import socket
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        with socket.socket() as sk:                      # a free rendezvous port on the loopback interface
            sk.bind(('127.0.0.1', 0))
            port = sk.getsockname()[1]
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        for r in range(2):
            errs, nbytes, n_exp = ret[r]
            assert n_exp > 0 and nbytes == n_exp * 150 * 64
            for e_r, e_x in errs:
                assert e_r < 1e-5 and e_x < 1e-5, (r, errs)
