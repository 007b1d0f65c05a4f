"""CPU oracle for GraphDD's location network (TEST INFRASTRUCTURE — never on the product path).

A functional restatement, in plain torch-CPU fp32, of the second consumer of the DataAggregation kernel family
(SURVEY.md §8f rank 4): `Relocation/train_double_difference_model.py:333-505` —

  * `DataAggregation` (:333-388): the product-graph aggregation of the detection model, but every message passes through a
    per-EDGE layer first, `merge_edges([x_j | (pos_j - pos_i) / scale_rel])` (:386-388), before the mean;
  * `BipartiteGraphOperator` / `BipartiteGraphOperatorSta` (:390-436): per-edge `fc1` (Linear-PReLU-Linear) of
    `[x_j | mask_j | (pos_i - pos_j) / scale_rel]`, mean onto the sources / the stations;
  * `GNN_Location.forward` (:484-505): five DataAggregation blocks, the two read-outs and the small projection heads.

Pinned by tests/golden/graphdd_*.npz, which oracle/gen_golden.py `graphdd` produced by executing the reference's class
definitions (read from the reference file at generation time; the file is a script with top-level code and cannot be
imported) through oracle/refshim.  State: flat dict with the reference's state_dict key names.  The third-party message
passing semantics (flow source -> target, `_j` = edge_index[0], `_i` = edge_index[1], mean = sum / max(count, 1), output rows =
size[1]) are the ones oracle/genie_oracle.py restates.
"""
import torch

from .genie_oracle import _lin, propagate_mean


def _prelu_w(sd, name, x):
    a = sd[name].reshape(())
    return torch.where(x >= 0, x, a * x)


def data_aggregation(sd, pre, tr, mask, A_in_sta, A_in_src, A_src_in_sta, pos_loc, pos_src, scale_rel=30.0):
    """:357-388.  `pre` = state_dict prefix, e.g. 'DataAggregation1.'."""
    n = tr.shape[0]

    def agg(edges, x, edge_attr):                                                          # propagate + message (:386-388)
        m = _lin(sd, pre + 'merge_edges.0', torch.cat((x.index_select(0, edges[0]), edge_attr), dim=1))
        return propagate_mean(_prelu_w(sd, pre + 'merge_edges.1.weight', m), edges[1], n)

    tr = _prelu_w(sd, pre + 'activate.weight', _lin(sd, pre + 'init_trns', torch.cat((tr, mask), dim=-1)))     # :359-360
    sta_of, src_of = A_src_in_sta[0], A_src_in_sta[1]
    pos_rel_sta = (pos_loc[sta_of[A_in_sta[0]]] / 1000.0 - pos_loc[sta_of[A_in_sta[1]]] / 1000.0) / scale_rel   # :364
    pos_rel_src = (pos_src[src_of[A_in_src[0]]] / 1000.0 - pos_src[src_of[A_in_src[1]]] / 1000.0) / scale_rel   # :365
    a11 = _prelu_w(sd, pre + 'activate11.weight', tr)
    a12 = _prelu_w(sd, pre + 'activate12.weight', tr)
    tr1 = _lin(sd, pre + 'l1_t1_2', torch.cat((tr, agg(A_in_sta, a11, pos_rel_sta), mask), dim=1))              # :368
    tr2 = _lin(sd, pre + 'l1_t2_2', torch.cat((tr, agg(A_in_src, a12, pos_rel_src), mask), dim=1))              # :369
    tr = _prelu_w(sd, pre + 'activate1.weight', torch.cat((tr1, tr2), dim=1))
    a21 = _prelu_w(sd, pre + 'activate21.weight', _lin(sd, pre + 'l2_t1_1', tr))
    a22 = _prelu_w(sd, pre + 'activate22.weight', _lin(sd, pre + 'l2_t2_1', tr))
    tr1 = _lin(sd, pre + 'l2_t1_2', torch.cat((tr, agg(A_in_sta, a21, pos_rel_sta), mask), dim=1))              # :372
    tr2 = _lin(sd, pre + 'l2_t2_2', torch.cat((tr, agg(A_in_src, a22, pos_rel_src), mask), dim=1))              # :373
    return _prelu_w(sd, pre + 'activate2.weight', torch.cat((tr1, tr2), dim=1))


def bipartite_read_out(sd, pre, x, mask, edges, pos_j, pos_i, n_out, scale_rel=30e3):
    """:390-436: per-edge fc1 on [x_j | mask_j | (pos_i - pos_j) / scale_rel], activate1, mean onto the n_out targets, fc2,
    activate2.  pos_j: [N,3] positions indexed by edges[0]; pos_i: [n_out,3] indexed by edges[1]."""
    xm = torch.cat((x, mask), dim=1)
    e = torch.cat((xm.index_select(0, edges[0]), (pos_i.index_select(0, edges[1]) - pos_j.index_select(0, edges[0])) / scale_rel),
                  dim=1)
    h = _lin(sd, pre + 'fc1.2', _prelu_w(sd, pre + 'fc1.1.weight', _lin(sd, pre + 'fc1.0', e)))
    h = _prelu_w(sd, pre + 'activate1.weight', h)
    return _prelu_w(sd, pre + 'activate2.weight', _lin(sd, pre + 'fc2', propagate_mean(h, edges[1], n_out)))


def _seq(sd, pre, x):
    """nn.Sequential(Linear, PReLU, Linear)."""
    return _lin(sd, pre + '2', _prelu_w(sd, pre + '1.weight', _lin(sd, pre + '0', x)))


def gnn_location(sd, x, mask, A_in_pick, A_in_src, A_src_in_product, A_sta_in_product, A_src_in_sta, locs_cart, srcs_cart,
                 memory=None, scale_fixed=5000.0):
    """GNN_Location.forward (:484-505) -> (scale * proj(x1), proj_t(x1), proj_c(x2), x)."""
    if memory is not None:                                                                  # use_memory (:486-488)
        mask = _seq(sd, 'embed_inpt.', torch.cat((mask, memory[A_src_in_sta[1]]), dim=1))
        x = torch.cat((x, memory[A_src_in_sta[1]]), dim=1)
    else:
        mask = _seq(sd, 'embed_inpt.', mask)
    for i in range(1, 6):
        x = data_aggregation(sd, 'DataAggregation%d.' % i, x, mask, A_in_pick, A_in_src, A_src_in_sta, locs_cart, srcs_cart)
    x1 = bipartite_read_out(sd, 'BipartiteReadOut1.', x, mask, A_src_in_product, locs_cart[A_src_in_sta[0]], srcs_cart,
                            srcs_cart.shape[0])                                             # :498, :407
    x2 = bipartite_read_out(sd, 'BipartiteReadOut2.', x, mask, A_sta_in_product, srcs_cart[A_src_in_sta[1]], locs_cart,
                            locs_cart.shape[0])                                             # :499, :430
    if memory is not None:
        x1 = _seq(sd, 'merge_data.', torch.cat((x1, _seq(sd, 'proj_memory.', memory)), dim=1))                  # :501-503
    return scale_fixed * _seq(sd, 'proj.', x1), _seq(sd, 'proj_t.', x1), _seq(sd, 'proj_c.', x2), x
