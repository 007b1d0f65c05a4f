"""torch_geometric.nn.MessagePassing stand-in (documented PyG semantics, CPU/torch only).

flow = source_to_target: j = edge_index[0] (message source), i = edge_index[1] (target / aggregation index).
"""
import inspect
import torch

_SPECIAL = ('edge_index', 'index', 'size', 'size_i', 'size_j', 'ptr', 'dim_size')


def _scatter(src, index, dim, dim_size, reduce):
    from torch_scatter import scatter
    return scatter(src, index, dim=dim, dim_size=dim_size, reduce=reduce)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):
        super().__init__()
        assert flow == 'source_to_target'
        self.aggr = 'sum' if aggr == 'add' else aggr
        self.node_dim = node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    # default message
    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs

    def _set_size(self, size, dim, src):
        n = src.size(self.node_dim)
        if size[dim] is None:
            size[dim] = n
        elif size[dim] != n:
            raise ValueError('refshim: inconsistent node count %d vs %d' % (size[dim], n))

    def propagate(self, edge_index, size=None, **kwargs):
        size = [None, None] if size is None else list(size)
        i, j = 1, 0
        coll = {}
        for arg in self._msg_params:
            if arg in _SPECIAL:
                continue
            if arg[-2:] not in ('_i', '_j'):
                coll[arg] = kwargs[arg] if arg in kwargs else inspect.signature(self.message).parameters[arg].default
                continue
            dim = j if arg[-2:] == '_j' else i
            data = kwargs[arg[:-2]]
            if isinstance(data, (tuple, list)):
                assert len(data) == 2
                if isinstance(data[1 - dim], torch.Tensor):
                    self._set_size(size, 1 - dim, data[1 - dim])
                data = data[dim]
            if isinstance(data, torch.Tensor):
                self._set_size(size, dim, data)
                data = data.index_select(self.node_dim, edge_index[dim])
            coll[arg] = data
        if 'edge_index' in self._msg_params:
            coll['edge_index'] = edge_index
        if 'index' in self._msg_params:
            coll['index'] = edge_index[i]
        if 'size_i' in self._msg_params:
            coll['size_i'] = size[i] if size[i] is not None else size[j]
        if 'size_j' in self._msg_params:
            coll['size_j'] = size[j] if size[j] is not None else size[i]
        dim_size = size[i] if size[i] is not None else size[j]
        msg = self.message(**coll)
        out = _scatter(msg, edge_index[i], self.node_dim, dim_size, self.aggr)
        return self.update(out)
