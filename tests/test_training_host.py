"""Host logic of the training path (genie_b200/training.py) without a GPU: the operator graph + the forward / transposed CSR
matrices of `MeanAggregate`, with the one CUDA call (`ops.kron_spmm`, checked on the GPU in test_gpu_parity.py) replaced by
a dense restatement of the same CSR product.  Loss and every parameter gradient must match the oracle's autograd."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import genie_oracle as go


def _dense_kron_spmm(kg, csr, x):
    rowptr, col, val = csr
    n = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    A = torch.zeros((n, n), dtype=torch.float64).index_put_((rows, col.long()), val.double(), accumulate=True)
    C = x.shape[1]
    if kg.mode == 0:      # out[g, s] = sum_s' A[s, s'] x[g, s']
        return torch.einsum('st,gtc->gsc', A, x.double().view(kg.n_grid, kg.n_sta, C)).reshape(-1, C).float()
    if kg.mode == 1:      # out[g, s] = sum_g' A[g, g'] x[g', s]
        return torch.einsum('gh,hsc->gsc', A, x.double().view(kg.n_grid, kg.n_sta, C)).reshape(-1, C).float()
    return (A @ x.double()).float()


class _FakePlan(object):
    """The attributes of GraphPlan the training path reads, built on the CPU."""

    def __init__(self, A_sta, A_src, S, G):
        from genie_b200.plan import csr_by_destination
        self.mode, self.n_sta, self.n_grid, self.n_prod, self.device = 0, S, G, S * G, torch.device('cpu')
        self.sta_rowptr, self.sta_col = csr_by_destination(A_sta, S)
        self.src_rowptr, self.src_col = csr_by_destination(A_src, G)

    def node_grid_index(self):
        return torch.arange(self.n_prod) // self.n_sta


@pytest.mark.parametrize('name', ['assoc_10x100', 'assoc_14of16x120_edges', 'assoc_14of16x120_abspos'])
def test_training_forward_and_gradients_match_oracle(name, monkeypatch):
    from genie_b200 import ops, training, module as gm
    from test_oracle_golden import assoc_inputs, assoc_variant
    d, sd = load_golden(name)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']),
                                                                       int(d['k_spc']))
    kw = assoc_inputs(d)
    pos_rel, abs_pos = assoc_variant(d, name, A_ps, A_pg, A_sis)
    Slice, Mask = torch.from_numpy(d['Slice']), torch.from_numpy(d['Mask'])
    attr = torch.from_numpy(d['read_in_attr'])
    grid = torch.from_numpy(d['grid']).float()
    locs = torch.from_numpy(d['sta'][d['ind_use']]).float()
    # oracle: the same state as leaf tensors
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = go.forward_fixed(sdo, Slice, Mask, A_ps, A_pg, attr, A_sip, A_src, grid, scale_rel=float(d['scale_rel']),
                            scale_t=float(d['scale_t']), eps=float(d['eps']), pos_rel=pos_rel, abs_pos=abs_pos, **kw)
    wts = [torch.linspace(0.5, 1.5, w.numel()).view_as(w) for w in want]
    loss_o = sum((w * t).sum() for w, t in zip(want, wts))
    loss_o.backward()
    # product: training path with the CUDA call replaced by its dense restatement
    monkeypatch.setattr(ops, 'kron_spmm', _dense_kron_spmm)
    monkeypatch.setattr(training, 'LIN_SPLIT_MIN_ROWS', 300)        # the fixtures' P ~ 10^3: take the split-K weight gradient
    monkeypatch.setattr(gm, 'knn_query_edges', lambda xc, xq, k: go.knn(xc / 1000.0, xq / 1000.0, k).flip(0))
    m = gm.GCN_Detection_Network_extended(None, None, scale_rel=float(d['scale_rel']), device='cpu',
                                          updated_model=name.endswith('_edges'), use_absolute_pos=name.endswith('_abspos'))
    m.load_state_dict(sd)
    m.TemporalAttention.scale_t = float(d['scale_t'])
    m.train()
    m._plan = _FakePlan(A_sta, A_src, S, G)
    m._read_in_attr = attr
    m.A_src_in_sta = A_sis
    if m.updated_model:
        m._set_edge_means(locs, grid, None)
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))          # the CPU guard of forward_train
    got = training.forward_train(m, Slice, Mask, None, A_src, kw['A_edges_p'], kw['A_edges_s'], kw['dt_partition'],
                                 kw['tlatent'], kw['tpick'], kw['ipick'], kw['phase_label'], locs, grid, kw['x_query_cart'],
                                 kw['x_query_src_cart'], kw['t_query'], kw['tq_sample'], kw['trv_out_q'])
    for a, b, key in zip(got, want, ('y', 'x', 'arv_p', 'arv_s')):
        assert rel_err(a.detach().numpy(), b.detach().numpy()) < 1e-5, key
        assert rel_err(a.detach().numpy(), d[key]) < 1e-5, key
    loss = sum((w * t).sum() for w, t in zip(got, wts))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_o.detach())) < 1e-5 * abs(float(loss_o.detach()))
    n_checked = 0
    for k, p in m.named_parameters():
        go_ = sdo[k].grad
        if go_ is None or p.grad is None:
            assert (go_ is None or not go_.any()) and (p.grad is None or not p.grad.any()), k
            continue
        assert rel_err(p.grad.numpy(), go_.numpy()) < 2e-4, k
        n_checked += 1
    assert n_checked > 100
