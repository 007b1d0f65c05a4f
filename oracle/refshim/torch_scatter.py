"""torch_scatter.scatter stand-in: sum / mean / max with zero-filled empty bins."""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    assert out is None
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    index = index.long()
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ('sum', 'add'):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(dim, index, src)
    if reduce == 'mean':
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(dim, index, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(
            0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device)).clamp_(min=1)
        view = [1] * src.dim()
        view[dim] = dim_size
        return s / cnt.view(view)
    if reduce == 'max':
        view = [1] * src.dim()
        view[dim] = index.numel()
        idx = index.view(view).expand_as(src)
        res = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
        res = res.scatter_reduce(dim, idx, src, reduce='amax', include_self=True)
        return torch.where(torch.isinf(res) & (res < 0), torch.zeros_like(res), res)
    raise ValueError(reduce)
