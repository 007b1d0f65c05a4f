"""Grid sharding (genie_b200/sharded.py) with world_size 2 over gloo on the CPU: partition, halo exchange and all-gather
logic.  The compute steps are supplied by a test-only backend built on the CPU oracle; the product backend (CudaBackend)
runs the same ShardedFrontEnd code path on GPUs (tests/test_gpu_parity.py::test_sharded_front_end_single_process)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err
from oracle import genie_oracle as go


class OracleBackend(object):
    """Same interface as genie_b200.sharded.CudaBackend, arithmetic from oracle/genie_oracle.py (module.py:85-98, 224-249)."""

    def __init__(self, sd, A_sta, A_src_local, S, n_local, n_owned, attr_local, A_src_global):
        from genie_b200.process_utils import product_edge_lists
        self.sd, self.S, self.n_local, self.n_owned = sd, S, n_local, n_owned
        self.A_ps, self.A_pg, self.A_sip, _ = product_edge_lists(A_sta, A_src_local, S, n_local)
        self.attr, self.A_src_global = attr_local, A_src_global

    def _agg(self, edges, x):
        return go.propagate_mean(x.index_select(0, edges[0]), edges[1], self.n_local * self.S)

    def layer1(self, Slice, Mask):
        sd, p = self.sd, 'DataAggregation.'
        tr0 = go._prelu(sd, p + 'activate', go._lin(sd, p + 'init_trns', torch.cat((Slice, Mask), dim=-1)))
        tr1 = go._lin(sd, p + 'l1_t1_2', torch.cat((tr0, self._agg(self.A_ps, go._prelu(sd, p + 'activate11', tr0)), Mask), 1))
        tr2 = go._lin(sd, p + 'l1_t2_2', torch.cat((tr0, self._agg(self.A_pg, go._prelu(sd, p + 'activate12', tr0)), Mask), 1))
        self.tr = go._prelu(sd, p + 'activate1', torch.cat((tr1, tr2), dim=1))
        self.Mask = Mask
        b = go._prelu(sd, p + 'activate22', go._lin(sd, p + 'l2_t2_1', self.tr))
        self.msg = b.view(self.n_local, self.S * 30).clone()
        self.msg[self.n_owned:] = float('nan')            # halo rows must come from their owners

    def message_rows(self):
        return self.msg

    def layer2_readin(self):
        sd, p = self.sd, 'DataAggregation.'
        assert not torch.isnan(self.msg).any()
        a = go._prelu(sd, p + 'activate21', go._lin(sd, p + 'l2_t1_1', self.tr))
        b = self.msg.view(-1, 30)
        o1 = go._lin(sd, p + 'l2_t1_2', torch.cat((self.tr, self._agg(self.A_ps, a), self.Mask), dim=1))
        o2 = go._lin(sd, p + 'l2_t2_2', torch.cat((self.tr, self._agg(self.A_pg, b), self.Mask), dim=1))
        x_latent = go._prelu(sd, p + 'activate2', torch.cat((o1, o2), dim=1))
        n = self.n_owned * self.S
        return go.bipartite_read_in(sd, 'Bipartite_ReadIn.', x_latent[:n], self.attr[:n], self.A_sip[:, :n], self.Mask[:n])

    def spatial(self, read_in, pos, scale_rel):
        x = read_in
        for i in (1, 2, 3):
            x = go.spatial_aggregation(self.sd, 'SpatialAggregation%d.' % i, x, self.A_src_global, pos, scale_rel)
        return x


def _case(S=12, G=260, seed=9):
    from genie_b200 import synth
    from genie_b200.process_utils import extract_inputs_adjacencies_cartesian
    net = synth.Network(S, G, seed=seed)
    A_sta, A_src = extract_inputs_adjacencies_cartesian(net.sta, net.grid, 6, 9)
    rng = np.random.default_rng(seed)
    Slice = torch.from_numpy((rng.random((S * G, 4)) * (rng.random((S * G, 4)) < 0.4)).astype(np.float32))
    Mask = (Slice.abs() > 0.01).float()
    attr = torch.from_numpy(net.read_in_offsets(30000.0))
    return net, A_sta, A_src, Slice, Mask, attr


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from genie_b200.sharded import GridPartition, ShardedFrontEnd
        net, A_sta, A_src, Slice, Mask, attr = _case()
        S, G = net.S, net.G
        sd = go.init_state(seed=6)
        part = GridPartition(A_src, G, world, group_size=16)
        nodes = torch.from_numpy(part.local_nodes(rank))
        loc = lambda x: x.view(G, S, -1).index_select(0, nodes).reshape(len(nodes) * S, -1).contiguous()
        be = OracleBackend(sd, A_sta, part.local_graph(rank), S, len(nodes), len(part.owned[rank]), loc(attr), A_src)
        fe = ShardedFrontEnd(part, rank, be, 'cpu')
        pos = torch.from_numpy(net.grid).float()
        xs, read_in = fe.forward(loc(Slice), loc(Mask), pos, 30000.0)
        if rank == 0:
            torch.save(dict(xs=xs, read_in=read_in, halo=[len(h) for h in part.halo]), out)
    finally:
        dist.destroy_process_group()


def test_sharded_front_end_world2_gloo(tmp_path):
    from genie_b200.process_utils import product_edge_lists
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    net, A_sta, A_src, Slice, Mask, attr = _case()
    sd = go.init_state(seed=6)
    A_ps, A_pg, A_sip, _ = product_edge_lists(A_sta, A_src, net.S, net.G)
    want, parts = go.front_end(sd, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, torch.from_numpy(net.grid).float(),
                               30000.0, return_parts=True)
    assert min(got['halo']) > 0                                   # the exchange really moved rows
    assert rel_err(got['read_in'].numpy(), parts['read_in'].numpy()) < 1e-5
    assert rel_err(got['xs'].numpy(), want.numpy()) < 1e-5


@pytest.mark.parametrize('world', [2, 3, 8])
def test_grid_partition_invariants(world):
    from genie_b200.sharded import GridPartition
    net, A_sta, A_src, _, _, _ = _case(G=700)
    part = GridPartition(A_src, net.G, world, group_size=16)
    owned = np.concatenate(part.owned)
    assert np.array_equal(np.sort(owned), np.arange(net.G))       # every grid node has exactly one owner
    total_send = 0
    for r in range(world):
        nodes = part.local_nodes(r)
        assert len(np.unique(nodes)) == len(nodes)
        e = part.local_graph(r).numpy()
        assert e[1].max() < len(part.owned[r]) and e[0].max() < len(nodes)
        # local in-edges of owned nodes are exactly the global ones
        src_g = nodes[e[0]]
        tgt_g = nodes[e[1]]
        glob = A_src.numpy()
        mask = np.isin(glob[1], part.owned[r])
        assert sorted(zip(src_g.tolist(), tgt_g.tolist())) == sorted(zip(glob[0][mask].tolist(), glob[1][mask].tolist()))
        send_rows, send_counts, recv_counts = part.exchange_lists(r)
        assert sum(recv_counts) == len(part.halo[r]) and recv_counts[r] == 0 and send_counts[r] == 0
        total_send += sum(send_counts)
    assert total_send == sum(len(h) for h in part.halo)
