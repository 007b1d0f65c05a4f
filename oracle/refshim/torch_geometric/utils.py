"""torch_geometric.utils stand-ins (remove_self_loops, subgraph, softmax, degree, ...)."""
import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, keep]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[keep]


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros(n, dtype=dtype if dtype is not None else torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype, device=index.device))


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    if not isinstance(subset, torch.Tensor):
        subset = torch.as_tensor(subset, dtype=torch.long)
    subset = subset.to(edge_index.device)
    n = num_nodes
    if n is None:
        n = int(max(int(edge_index.max()) if edge_index.numel() else -1, int(subset.max()) if subset.numel() else -1)) + 1
    if subset.dtype == torch.bool:
        node_mask = subset
        subset = node_mask.nonzero().view(-1)
    else:
        node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
        node_mask[subset] = True
    keep = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, keep]
    ea = edge_attr[keep] if edge_attr is not None else None
    if relabel_nodes:
        relabel = torch.zeros(n, dtype=torch.long, device=edge_index.device)
        relabel[subset] = torch.arange(subset.numel(), device=edge_index.device)
        ei = relabel[ei]
    return ei, ea


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    assert dim == 0
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    mx = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src.detach(), reduce='amax', include_self=True)
    out = (src - mx.gather(0, idx)).exp()
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, out)
    return out / (den.gather(0, idx) + 1e-16)


def to_undirected(edge_index, *a, **k):
    ei = torch.cat((edge_index, edge_index.flip(0)), dim=1)
    return torch.unique(ei, dim=1)


def to_networkx(*a, **k):
    raise NotImplementedError('refshim: to_networkx is not on the hot path')
