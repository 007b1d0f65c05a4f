"""Host side of the association branch, runnable without a GPU:

* the pick-sized source-arrival attention (`Arrivals`, plain torch in the product) against the unmodified reference;
* the packed weight blob of the association kernels: a numpy interpreter that follows the KERNELS' algebra (K-major rows,
  the linearity split of the last layer, the [o1 0 | o2 0] row layout) reads the blob through genie_assoc_layout's offsets
  and must reproduce the reference's intermediate tensors.  The CUDA kernels themselves are checked in test_gpu_parity.py.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

ASSOC = ['assoc_10x100', 'assoc_18of20x160']


def _cpu_model(sd, d):
    from genie_b200.module import GCN_Detection_Network_extended
    m = GCN_Detection_Network_extended(None, None, scale_rel=float(d['scale_rel']), device='cpu')
    m.load_state_dict(sd)
    m.eval()
    return m


@pytest.mark.parametrize('name', ASSOC)
def test_arrivals_attention_matches_reference(name):
    d, sd = load_golden(name)
    m = _cpu_model(sd, d)
    t = torch.from_numpy
    with torch.no_grad():
        arv = m.Arrivals(t(d['x_query_src']).float(), t(d['tq_sample']).float(), t(d['x_src']), t(d['trv_out_q']).float(),
                         None, t(d['arv_p_embed']), t(d['arv_s_embed']), t(d['tpick']).float(), t(d['ipick']).long(),
                         t(d['phase_label']).long().reshape(-1, 1))
    assert arv.shape == (len(d['tq_sample']), len(d['tpick']), 2)
    assert rel_err(arv[:, :, 0:1].numpy(), d['arv_p']) < 5e-6
    assert rel_err(arv[:, :, 1:2].numpy(), d['arv_s']) < 5e-6


def test_arrivals_attention_edge_cases():
    """No picks; a source whose origin time puts every edge (incl. the null arrival) outside the 2 eps window."""
    d, sd = load_golden(ASSOC[0])
    m = _cpu_model(sd, d)
    t = torch.from_numpy
    with torch.no_grad():
        out = m.Arrivals.forward_merged(torch.tensor([1000.0]), t(d['x_src'])[:1], t(d['trv_out_q']).float()[:1],
                                        torch.zeros(len(d['tpick']) + 1, 30), t(d['tpick']).float(), t(d['ipick']).long(),
                                        t(d['phase_label']).long())
        base = m.Arrivals.proj_2(m.Arrivals.activate4(m.Arrivals.proj_1(torch.zeros(1, 15))))
        assert torch.allclose(out, base.view(1, 1, 2).expand_as(out))
        out0 = m.Arrivals.forward_merged(torch.tensor([0.0]), t(d['x_src'])[:1], t(d['trv_out_q']).float()[:1],
                                         torch.zeros(1, 30), torch.zeros(0), torch.zeros(0, dtype=torch.long),
                                         torch.zeros(0, 1))
        assert out0.shape == (1, 0, 2)


def _prelu(x, a):
    return np.where(x >= 0, x, a * x)


def _mean_over(A, x, n):
    out = np.zeros((n, x.shape[1]), dtype=np.float64)
    np.add.at(out, A[1], x[A[0]])
    cnt = np.bincount(A[1], minlength=n).astype(np.float64)
    return out / np.maximum(cnt, 1.0)[:, None]


@pytest.mark.parametrize('name', ASSOC)
def test_assoc_packed_blob_reproduces_reference(name):
    from genie_b200 import ops
    from oracle import genie_oracle as go
    d, sd = load_golden(name)
    m = _cpu_model(sd, d)
    aw = ops.AssocWeights('cpu')
    assert ops.AssocWeights.supported(m)
    buf = aw.update(m).numpy().astype(np.float64)
    o = aw.off

    def mat(nm, n_in, ld, n_out):
        return buf[o[nm]:o[nm] + n_in * ld].reshape(n_in, ld)[:, :n_out]

    def vec(nm, n):
        return buf[o[nm]:o[nm] + n]

    sl = vec('SL', 16)
    S, G = len(d['ind_use']), d['grid'].shape[0]
    A_sta, A_src, A_ps, A_pg, A_sip, A_sis = go.build_adjacencies_dense(d['sta'][d['ind_use']], d['grid'], int(d['k_sta']),
                                                                       int(d['k_spc']))
    A_ps, A_pg = A_ps.numpy(), A_pg.numpy()
    P = S * G
    grid_of = np.arange(P) // S
    # assoc_grid_pre_kernel
    yl = _prelu(d['x_spatial'].astype(np.float64) @ mat('SD_W', 30, 32, 30) + vec('SD_B', 30), sl[0])
    assert rel_err(yl, d['y_latent']) < 1e-5
    yfc1 = yl @ mat('RO_WY', 30, 32, 30) + vec('RO_B1', 30)
    mask_out = (d['y'][:, :, 0].max(1) > np.float32(0.01)).astype(np.float64)
    # assoc_init_kernel
    mo = mask_out[grid_of][:, None]
    h = yfc1[grid_of] + d['read_in_attr'].astype(np.float64) @ mat('RO_WA', 3, 32, 30)
    s0 = _prelu((mo * _prelu(h, sl[1])) @ mat('RO_W2', 30, 16, 15) + vec('RO_B2', 15), sl[2])
    assert rel_err(s0, d['assoc_s0']) < 1e-5
    x0 = np.concatenate((s0, d['x_latent'], mo, d['Mask']), axis=1)
    tr = _prelu(x0 @ mat('AI_W', 50, 32, 30) + vec('AI_B', 30), sl[3])
    a1 = _prelu(tr @ mat('M11_W', 30, 32, 30) + vec('M11_B', 30), sl[4])
    a2 = _prelu(tr @ mat('M12_W', 30, 32, 30) + vec('M12_B', 30), sl[5])
    # assoc_layer1_kernel
    mask5 = np.concatenate((mo, d['Mask']), axis=1)
    f1 = np.concatenate((tr, _mean_over(A_ps, a1, P), mask5), axis=1)
    f2 = np.concatenate((tr, _mean_over(A_pg, a2, P), mask5), axis=1)
    trb = _prelu(np.concatenate((f1 @ mat('W11', 65, 32, 30) + vec('B11', 30), f2 @ mat('W12', 65, 32, 30) + vec('B12', 30)),
                                axis=1), sl[6])
    va = _prelu(trb @ mat('W21A', 60, 32, 30) + vec('B21A', 30), sl[7]) @ mat('WVA', 30, 16, 15)
    vb = _prelu(trb @ mat('W22A', 60, 32, 30) + vec('B22A', 30), sl[8]) @ mat('WVB', 30, 16, 15)
    fc = np.concatenate((trb, mask5), axis=1)
    ca = fc @ mat('WCA', 65, 16, 15) + vec('BCA', 15)
    cb = fc @ mat('WCB', 65, 16, 15) + vec('BCB', 15)
    # assoc_layer2_kernel: rows [o1 0 | o2 0]
    rows = np.zeros((P, 32))
    rows[:, 0:15] = _prelu(ca + _mean_over(A_ps, va, P), sl[9])
    rows[:, 16:31] = _prelu(cb + _mean_over(A_pg, vb, P), sl[9])
    s = np.concatenate((rows[:, 0:15], rows[:, 16:31]), axis=1)
    assert rel_err(s, d['assoc_s']) < 1e-5
    # assoc_collapse_kernel
    eps = float(d['eps'])
    dtp = d['dt_partition'].astype(np.float32)
    dt0, dstep = dtp[0], np.float32(dtp[1] - dtp[0])
    tp = d['tpick'].astype(np.float32)
    l_dt, k = len(dtp), 10
    for ph, (pre, key, edges) in enumerate((('CP', 'arv_p_embed', d['A_edges_p']), ('CS', 'arv_s_embed', d['A_edges_s']))):
        W1, B1 = mat(pre + '_W1', 32, 32, 30), vec(pre + '_B1', 30)
        W2, B2 = mat(pre + '_W2', 30, 16, 15), vec(pre + '_B2', 15)
        out = np.zeros((len(tp), 15))
        for a in range(len(tp)):
            tq = int(np.floor((tp[a] - dt0) / dstep))
            base = (int(d['ipick'][a]) * l_dt + tq) * k
            acc, cnt = np.zeros(30), 0
            for e in range(k):
                j = int(edges[base + e])
                t_rel = np.float32(tp[a] - np.float32(d['tlatent'][j, ph]))
                if not abs(t_rel) < 2.0 * eps:
                    continue
                cnt += 1
                col = [c if c < 15 else c + 1 for c in range(30)]
                x = np.concatenate((rows[j, col], [t_rel / eps, float(d['phase_label'][a])]))
                acc += _prelu(x @ W1 + B1, sl[10 + 2 * ph])
            mean = acc / cnt if cnt else acc
            out[a] = _prelu(mean @ W2 + B2, sl[11 + 2 * ph])
        assert rel_err(out, d[key]) < 1e-5, key


def test_association_needs_cuda_tensors():
    """No CPU path: the product raises instead of computing the association branch on host tensors."""
    from genie_b200 import capi
    d, sd = load_golden(ASSOC[0])
    m = _cpu_model(sd, d)
    with pytest.raises((capi.GenieError, RuntimeError)):
        m.forward_fixed(torch.zeros(10, 4), torch.zeros(10, 4), None, None, None, torch.zeros(2, 3), torch.zeros(5, 3),
                        torch.zeros(1, 3), torch.zeros(1, 3), torch.zeros(3, 1), torch.zeros(1), torch.zeros(1, 2, 2))
