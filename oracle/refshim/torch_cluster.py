"""torch_cluster.knn stand-in: for each point of y the k nearest points of x, nearest first.

Returns LongTensor [2, |y|*k]: row 0 = index into y, row 1 = index into x (torch_cluster convention).
"""
import numpy as np
import torch
from scipy.spatial import cKDTree


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1):
    assert batch_x is None and batch_y is None and not cosine
    xn = x.detach().cpu().double().numpy()
    yn = y.detach().cpu().double().numpy()
    k = int(min(k, xn.shape[0]))
    _, ind = cKDTree(xn).query(yn, k=k)
    ind = ind.reshape(yn.shape[0], k)
    row = np.repeat(np.arange(yn.shape[0]), k)
    return torch.from_numpy(np.stack((row, ind.reshape(-1)), axis=0)).long().to(x.device)
