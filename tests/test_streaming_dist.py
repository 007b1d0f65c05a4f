"""DayProcessor.run_distributed: two ranks (two processes sharing cuda:0, gloo rendezvous on 127.0.0.1) split the day's windows
round-robin and all-reduce their partial Out_2 stacks; the result must equal the single-process loop.  Needs a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, REPO

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tsteps, tsteps_abs, ret):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import torch.distributed as dist
    from genie_b200.module import GCN_Detection_Network_extended
    from genie_b200.process_utils import InputExtractor
    from genie_b200.streaming import DayProcessor
    from oracle import genie_oracle as go
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cuda:0')
    d, sd = load_golden('assoc_10x100')
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src = go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']), int(d['k_spc']))[:2]
    m = GCN_Detection_Network_extended(None, None, scale_rel=float(d['scale_rel']), device=dev)
    m.load_state_dict(sd)
    m.TemporalAttention.scale_t = float(d['scale_t'])
    m.eval()
    m.set_adjacencies_cartesian(A_sta, A_src, torch.from_numpy(d['read_in_attr']).to(dev), S, G, device=dev)
    ex = InputExtractor(m._plan, d['trv_times'], d['ind_use'], d['sta'].shape[0], float(d['max_t']), float(d['kernel_sig_t']),
                        float(d['dt']))
    ex.set_day(d['picks'])
    dp = DayProcessor(m, ex, torch.from_numpy(d['sta'][d['ind_use']]).float().to(dev), torch.from_numpy(d['grid']).float().to(dev),
                      torch.from_numpy(d['x_query']).float().to(dev))
    out = dp.run_distributed(tsteps, tsteps_abs)
    if rank == 0:
        single = DayProcessor(m, ex, dp.locs, dp.grid, dp.xq).run(tsteps, tsteps_abs)
        ret['dist'], ret['single'], ret['done'] = out.cpu().numpy(), single.cpu().numpy(), dp.windows_done
    dist.barrier()
    dist.destroy_process_group()


def test_day_processor_two_ranks_allreduce_equals_single_process():
    if not torch.cuda.is_available():
        pytest.fail('needs a CUDA device')
    import torch.multiprocessing as mp
    tsteps = np.arange(0.0, 60.0, 3.0)
    tsteps_abs = np.arange(-3.0, 70.0 + 0.75, 0.75)
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        import socket
        with socket.socket() as sk:                      # a free rendezvous port on the loopback interface
            sk.bind(('127.0.0.1', 0))
            port = sk.getsockname()[1]
        procs = [ctx.Process(target=_worker, args=(r, 2, port, tsteps, tsteps_abs, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        a, b = ret['dist'], ret['single']
        assert 0 < ret['done'] <= len(tsteps) // 2 + 1
        assert np.abs(b).max() > 0 and rel_err(a, b) < 1e-6
