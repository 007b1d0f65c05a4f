"""Operator-level Python wrappers over the C-ABI (one function per entry point of include/genie_b200.h).

Every function takes CUDA tensors, allocates its outputs with torch and launches on torch's current stream.
"""
import ctypes
import math

import numpy as np
import torch

from . import capi

F32 = torch.float32


def _lin(ws, field, mod):
    lin = getattr(ws, field)
    lin.weight = capi.dptr(mod.weight.detach(), F32, field + '.weight')
    lin.bias = capi.dptr(mod.bias.detach(), F32, field + '.bias')


class PackedWeights(object):
    """Kernel-layout copy of the front-end parameters; re-packed whenever a source tensor changes."""

    def __init__(self, device):
        self.device = torch.device(device)
        n = int(capi.load().genie_frontend_packed_floats())
        self.buf = torch.zeros(n, dtype=F32, device=self.device)
        self._key = None

    @staticmethod
    def _sources(model):
        da, ri = model.DataAggregation, model.Bipartite_ReadIn
        sas = (model.SpatialAggregation1, model.SpatialAggregation2, model.SpatialAggregation3)
        ts = [da.init_trns, da.l1_t1_2, da.l1_t2_2, da.l2_t1_1, da.l2_t2_1, da.l2_t1_2, da.l2_t2_2, da.activate,
              da.activate11, da.activate12, da.activate1, da.activate21, da.activate22, da.activate2, ri.fc1, ri.fc2,
              ri.activate1, ri.activate2]
        for sa in sas:
            ts += [sa.fc1, sa.fc2, sa.fglobal, sa.activate1, sa.activate2, sa.activate3]
        return da, ri, sas, ts

    def update(self, model, relaid=None, init_relaid=None):
        """`relaid`: updated model definition only — (l1_t1_2, l1_t2_2, l2_t1_2, l2_t2_2) weights without their four
        edge-feature columns ([30,64] / [15,94]); those columns live in the plan's edge-term tables."""
        da, ri, sas, mods = self._sources(model)
        if getattr(self, '_plist_for', None) is not model:           # nn.Module.parameters() walks the tree: ~70 us per call
            self._plist, self._plist_for = [p for m in mods for p in m.parameters()], model
        key = tuple((p.data_ptr(), p._version) for p in self._plist)
        if relaid is not None:
            key = key + tuple(t.data_ptr() for t in relaid)
        if init_relaid is not None:                 # use_absolute_pos: init_trns without its six position columns [30,8]
            key = key + (init_relaid.data_ptr(),)
        if key == self._key:
            return self.buf
        for m in mods:
            for p in m.parameters():
                if p.device != self.device or p.dtype != F32 or not p.is_contiguous():
                    raise capi.GenieError('front-end parameters must be contiguous fp32 tensors on %s' % self.device)
        ws = capi.FrontendWeights()
        for f, m in (('da_init_trns', da.init_trns), ('da_l1_t1_2', da.l1_t1_2), ('da_l1_t2_2', da.l1_t2_2),
                     ('da_l2_t1_1', da.l2_t1_1), ('da_l2_t2_1', da.l2_t2_1), ('da_l2_t1_2', da.l2_t1_2),
                     ('da_l2_t2_2', da.l2_t2_2), ('ri_fc1', ri.fc1), ('ri_fc2', ri.fc2)):
            _lin(ws, f, m)
        if init_relaid is not None:
            ws.da_init_trns.weight = capi.dptr(init_relaid, F32, 'da_init_trns.weight (re-laid)')
            self._init_relaid = init_relaid         # keep the tensor alive
        elif da.init_trns.weight.shape[1] != 8:
            raise capi.GenieError('init_trns of a use_absolute_pos model needs its re-laid copy')
        if relaid is not None:
            for f, t in zip(('da_l1_t1_2', 'da_l1_t2_2', 'da_l2_t1_2', 'da_l2_t2_2'), relaid):
                getattr(ws, f).weight = capi.dptr(t, F32, f + '.weight (re-laid)')
            self._relaid = relaid                   # keep the tensors alive
        else:
            for m, n_in in ((da.l1_t1_2, 64), (da.l1_t2_2, 64), (da.l2_t1_2, 94), (da.l2_t2_2, 94)):
                if m.weight.shape[1] != n_in:
                    raise capi.GenieError('DataAggregation weights of the updated model need their re-laid copies')
        for f, m in (('da_activate', da.activate), ('da_activate11', da.activate11), ('da_activate12', da.activate12),
                     ('da_activate1', da.activate1), ('da_activate21', da.activate21),
                     ('da_activate22', da.activate22), ('da_activate2', da.activate2),
                     ('ri_activate1', ri.activate1), ('ri_activate2', ri.activate2)):
            setattr(ws, f, capi.dptr(m.weight.detach(), F32, f))
        for i, sa in enumerate(sas):
            _lin(ws.sa[i], 'fc1', sa.fc1)
            _lin(ws.sa[i], 'fc2', sa.fc2)
            _lin(ws.sa[i], 'fglobal', sa.fglobal)
            for a in ('activate1', 'activate2', 'activate3'):
                setattr(ws.sa[i], a, capi.dptr(getattr(sa, a).weight.detach(), F32, a))
        with torch.cuda.device(self.device):
            capi.check(capi.load().genie_frontend_pack_weights(ctypes.byref(ws), capi.dptr(self.buf),
                                                               capi.stream_ptr(self.device)))
        self._key = key
        return self.buf


class HeadsWeights(object):
    """Kernel-layout copy of the read-out head parameters (SpatialDirect, TemporalAttention, SpatialAttention) and the
    t_query branch of TemporalAttention folded into f_context_2 (include/genie_b200.h, genie_heads_*).  Re-packed whenever
    a parameter changes; the fold is cached per t_query tensor."""
    NAMES = ('SD_W', 'SD_B', 'SD_SL', 'TA_WC1', 'TA_BC1', 'TA_WV1', 'TA_BV1', 'TA_WV2', 'TA_BV2', 'TA_WP1', 'TA_BP1',
             'TA_WP2', 'TA_BP2', 'TA_SL', 'SA_WQ', 'SA_BQ', 'SA_WC', 'SA_BC', 'SA_WV', 'SA_BV', 'SA_WP', 'SA_BP', 'SA_SL')

    def __init__(self, device):
        self.device = torch.device(device)
        lib = capi.load()
        off = (ctypes.c_int32 * len(self.NAMES))()
        capi.check(lib.genie_heads_layout(off, len(self.NAMES)))
        self.off = dict(zip(self.NAMES, [int(o) for o in off]))
        self.buf = torch.zeros(int(lib.genie_heads_packed_floats()), dtype=F32, device=self.device)
        self._key, self._fold = None, None

    @staticmethod
    def supported(model):
        sd, ta, sa = model.SpatialDirect, model.TemporalAttention, model.SpatialAttention
        return (tuple(sd.f_direct.weight.shape) == (30, 30) and tuple(ta.f_context_1.weight.shape) == (30, 30) and
                tuple(ta.f_context_2.weight.shape) == (75, 30) and tuple(ta.proj_2.weight.shape) == (1, 30) and
                ta.n_heads == 5 and ta.n_latent == 15 and tuple(sa.f_context.weight.shape) == (75, 33) and
                tuple(sa.proj.weight.shape) == (30, 15) and sa.n_heads == 5 and sa.n_latent == 15)

    def _mat(self, name, lin_w, ld):
        w = lin_w.detach().t().contiguous()                         # [n_in, n_out]
        n_in, n_out = w.shape
        o = self.off[name]
        self.buf[o:o + n_in * ld].view(n_in, ld)[:, :n_out] = w

    def _vec(self, name, v):
        v = v.detach().reshape(-1)
        self.buf[self.off[name]:self.off[name] + v.numel()] = v

    def update(self, model, t_query):
        sd, ta, sa = model.SpatialDirect, model.TemporalAttention, model.SpatialAttention
        mods = (sd, ta, sa)
        if getattr(self, '_plist_for', None) is not model:
            self._plist, self._plist_for = [p for m in mods for p in m.parameters()], model
        key = tuple((p.data_ptr(), p._version) for p in self._plist)
        if key != self._key:
            with torch.no_grad():
                self.buf.zero_()
                self._mat('SD_W', sd.f_direct.weight, 32); self._vec('SD_B', sd.f_direct.bias)
                self._vec('SD_SL', sd.activate.weight)
                self._mat('TA_WC1', ta.f_context_1.weight, 32); self._vec('TA_BC1', ta.f_context_1.bias)
                self._mat('TA_WV1', ta.f_values_1.weight, 32); self._vec('TA_BV1', ta.f_values_1.bias)
                self._mat('TA_WV2', ta.f_values_2.weight, 76); self._vec('TA_BV2', ta.f_values_2.bias)
                self._mat('TA_WP1', ta.proj_1.weight, 32); self._vec('TA_BP1', ta.proj_1.bias)
                self._vec('TA_WP2', ta.proj_2.weight); self._vec('TA_BP2', ta.proj_2.bias)
                self._vec('TA_SL', torch.cat([a.weight.detach().reshape(1) for a in
                                              (ta.activate1, ta.activate2, ta.activate4, ta.activate5)]))
                self._mat('SA_WQ', sa.f_queries.weight, 76); self._vec('SA_BQ', sa.f_queries.bias)
                self._mat('SA_WC', sa.f_context.weight, 76); self._vec('SA_BC', sa.f_context.bias)
                self._mat('SA_WV', sa.f_values.weight, 76); self._vec('SA_BV', sa.f_values.bias)
                self._mat('SA_WP', sa.proj.weight, 32); self._vec('SA_BP', sa.proj.bias)
                self._vec('SA_SL', torch.cat([sa.activate1.weight.detach().reshape(1), sa.activate2.weight.detach().reshape(1)]))
            self._key, self._fold = key, None
        fkey = (t_query.data_ptr(), t_query._version, tuple(t_query.shape), float(ta.scale_t))
        if self._fold is None or self._fold[0] != fkey:
            with torch.no_grad():
                T = t_query.shape[0]
                q = ta.temporal_query_2(ta.activate3(ta.temporal_query_1(t_query.reshape(-1, 1).float() / ta.scale_t)))
                q = q.view(T, 5, 15)
                w2 = ta.f_context_2.weight.detach().view(5, 15, 30)
                b2 = ta.f_context_2.bias.detach().view(5, 15)
                A = torch.einsum('thl,hlk->thk', q, w2) / ta.scale                  # [T, 5, 30]
                a0 = torch.einsum('thl,hl->th', q, b2) / ta.scale                   # [T, 5]
                fold = torch.zeros(T * 5 * 32 + T * 5, dtype=F32, device=self.device)
                fold[:T * 5 * 32].view(T * 5, 32)[:, :30] = A.reshape(T * 5, 30)
                fold[T * 5 * 32:] = a0.reshape(-1)
            self._fold = (fkey, fold, T, t_query)
        return self.buf, self._fold[1], self._fold[2]


HEADS_T_MAX = 25               # query times per kernel launch (T * 5 <= 128 columns of the folded query table)


def heads_fwd(hw, packed, fold, T, x_spatial, x_context, x_query, nbr, scale_rel, grid_rows=None, query_rows=None):
    """y [G,T,1], x [Q,T,1] of forward_fixed_source (module.py:1015-1020) in two kernels (per block of 25 query times).
    grid_rows = (g0, g1) / query_rows = (q0, q1): only those rows of y / x (the heads split over the ranks of a sharded run)."""
    dev = x_spatial.device
    x_spatial = _f32c(x_spatial, 'x_spatial')
    if grid_rows is not None or query_rows is not None:
        g0, g1 = grid_rows
        q0, q1 = query_rows
        y = torch.empty((g1 - g0, T, 1), dtype=F32, device=dev)
        x = torch.empty((q1 - q0, T, 1), dtype=F32, device=dev)
        if T > HEADS_T_MAX:
            raise capi.GenieError('heads_fwd: row ranges support at most %d query times' % HEADS_T_MAX)
        lib = capi.load()
        with torch.cuda.device(dev):
            if g1 > g0:
                capi.check(lib.genie_heads_grid_fwd(capi.dptr(packed, F32), capi.dptr(fold, F32), int(T),
                                                    capi.dptr(x_spatial[g0:g1], F32), int(x_spatial.stride(0)), int(g1 - g0),
                                                    capi.dptr(y), None, capi.stream_ptr(dev)))
            if q1 > q0:
                capi.check(lib.genie_heads_query_fwd(
                    capi.dptr(packed, F32), capi.dptr(fold, F32), int(T), capi.dptr(x_spatial, F32), int(x_spatial.stride(0)),
                    capi.dptr(_f32c(x_context, 'x_context'), F32), capi.dptr(_f32c(x_query, 'x_query')[q0:q1], F32),
                    capi.dptr(nbr[q0:q1], torch.int64), int(nbr.shape[1]), int(q1 - q0), float(scale_rel), capi.dptr(x), None,
                    capi.stream_ptr(dev)))
        return y, x
    G, Q = x_spatial.shape[0], x_query.shape[0]
    if T > HEADS_T_MAX:
        ys, xs = [], []
        for t0 in range(0, T, HEADS_T_MAX):
            t1 = min(T, t0 + HEADS_T_MAX)
            part = torch.cat((fold[t0 * 5 * 32:t1 * 5 * 32], fold[T * 5 * 32 + t0 * 5:T * 5 * 32 + t1 * 5])).contiguous()
            yp, xp = heads_fwd(hw, packed, part, t1 - t0, x_spatial, x_context, x_query, nbr, scale_rel)
            ys.append(yp)
            xs.append(xp)
        return torch.cat(ys, dim=1), torch.cat(xs, dim=1)
    y = torch.empty((G, T, 1), dtype=F32, device=dev)
    x = torch.empty((Q, T, 1), dtype=F32, device=dev)
    # the query-independent halves of SpatialAttention's per-edge layers, once per context node (GENIE_HEADS_PROJ_LD); worth
    # it whenever the queries touch more (query, neighbour) pairs than there are context nodes — always, in practice
    proj = torch.empty((G, capi.HEADS_PROJ_LD), dtype=F32, device=dev) if Q * max(int(nbr.shape[1]), 1) >= G else None
    lib = capi.load()
    with torch.cuda.device(dev):
        capi.check(lib.genie_heads_grid_fwd(capi.dptr(packed, F32), capi.dptr(fold, F32), int(T), capi.dptr(x_spatial, F32),
                                            int(x_spatial.stride(0)), int(G), capi.dptr(y), capi.dptr(proj),
                                            capi.stream_ptr(dev)))
        if Q > 0:
            capi.check(lib.genie_heads_query_fwd(
                capi.dptr(packed, F32), capi.dptr(fold, F32), int(T), capi.dptr(x_spatial, F32), int(x_spatial.stride(0)),
                capi.dptr(_f32c(x_context, 'x_context'), F32), capi.dptr(_f32c(x_query, 'x_query'), F32),
                capi.dptr(nbr, torch.int64), int(nbr.shape[1]), int(Q), float(scale_rel), capi.dptr(x), capi.dptr(proj),
                capi.stream_ptr(dev)))
    return y, x


class AssocWeights(object):
    """Kernel-layout copy of the association-branch parameters (BipartiteGraphReadOutOperator,
    DataAggregationAssociationPhase, LocalSliceLgCollapse{P,S}, + SpatialDirect for y_latent); see include/genie_b200.h,
    genie_assoc_*.  Re-packed whenever a parameter changes."""
    NAMES = ('SD_W', 'SD_B', 'RO_WY', 'RO_B1', 'RO_WA', 'RO_W2', 'RO_B2', 'AI_W', 'AI_B', 'M11_W', 'M11_B', 'M12_W', 'M12_B',
             'W11', 'W12', 'B11', 'B12', 'W21A', 'W22A', 'B21A', 'B22A', 'WVA', 'WVB', 'WCA', 'WCB', 'BCA', 'BCB',
             'CP_W1', 'CP_B1', 'CP_W2', 'CP_B2', 'CS_W1', 'CS_B1', 'CS_W2', 'CS_B2', 'SL')

    def __init__(self, device):
        self.device = torch.device(device)
        lib = capi.load()
        off = (ctypes.c_int32 * len(self.NAMES))()
        capi.check(lib.genie_assoc_layout(off, len(self.NAMES)))
        self.off = dict(zip(self.NAMES, [int(o) for o in off]))
        self.buf = torch.zeros(int(lib.genie_assoc_packed_floats()), dtype=F32, device=self.device)
        self._key = None

    @staticmethod
    def supported(model):
        ro, da = model.BipartiteGraphReadOutOperator, model.DataAggregationAssociationPhase
        cp, cs = model.LocalSliceLgCollapseP, model.LocalSliceLgCollapseS
        return (tuple(model.SpatialDirect.f_direct.weight.shape) == (30, 30) and tuple(ro.fc1.weight.shape) == (30, 33) and
                tuple(ro.fc2.weight.shape) == (15, 30) and tuple(da.init_trns.weight.shape) in ((30, 50), (30, 56)) and
                tuple(da.l1_t1_2.weight.shape) in ((30, 65), (30, 69)) and
                tuple(da.l2_t1_2.weight.shape) in ((15, 95), (15, 99)) and
                tuple(cp.fc1.weight.shape) == (30, 32) and tuple(cs.fc2.weight.shape) == (15, 30))

    def _mat(self, name, w, ld):
        """w: [n_out, n_in] slice of an nn.Linear weight -> K-major rows of `ld` floats."""
        w = w.detach().t().contiguous()
        n_in, n_out = w.shape
        o = self.off[name]
        self.buf[o:o + n_in * ld].view(n_in, ld)[:, :n_out] = w

    def _vec(self, name, v):
        v = v.detach().reshape(-1)
        self.buf[self.off[name]:self.off[name] + v.numel()] = v

    def update(self, model, relaid=None):
        """`relaid`: model variants only — dict of re-laid weight copies (`init_trns` [30,50] without the six position columns of
        use_absolute_pos; `l1_t1_2`, `l1_t2_2` [30,65] and `l2_t1_2`, `l2_t2_2` [15,95] without the four edge-feature columns
        of the updated model definition); those columns live in the tables of genie_assoc_set_terms."""
        relaid = relaid or {}
        sd, ro, da = model.SpatialDirect, model.BipartiteGraphReadOutOperator, model.DataAggregationAssociationPhase
        cp, cs = model.LocalSliceLgCollapseP, model.LocalSliceLgCollapseS
        mods = (sd, ro, da, cp, cs)
        if getattr(self, '_plist_for', None) is not model:
            self._plist, self._plist_for = [p for m in mods for p in m.parameters()], model
        key = tuple((p.data_ptr(), p._version) for p in self._plist) + tuple((k, relaid[k].data_ptr()) for k in sorted(relaid))
        if key == self._key:
            return self.buf
        for p in self._plist:
            if p.device != self.device or p.dtype != F32:
                raise capi.GenieError('association parameters must be fp32 tensors on %s' % self.device)
        w = lambda name: relaid[name] if name in relaid else getattr(da, name).weight
        for name, n_in in (('init_trns', 50), ('l1_t1_2', 65), ('l1_t2_2', 65), ('l2_t1_2', 95), ('l2_t2_2', 95)):
            if w(name).shape[1] != n_in:
                raise capi.GenieError('association weights of a model variant need their re-laid copies (%s)' % name)
        self._relaid = relaid                                  # keep the tensors alive
        with torch.no_grad():
            self.buf.zero_()
            self._mat('SD_W', sd.f_direct.weight, 32); self._vec('SD_B', sd.f_direct.bias)
            self._mat('RO_WY', ro.fc1.weight[:, 0:30], 32); self._vec('RO_B1', ro.fc1.bias)
            self._mat('RO_WA', ro.fc1.weight[:, 30:33], 32)
            self._mat('RO_W2', ro.fc2.weight, 16); self._vec('RO_B2', ro.fc2.bias)
            self._mat('AI_W', w('init_trns'), 32); self._vec('AI_B', da.init_trns.bias)
            self._mat('M11_W', da.l1_t1_1.weight, 32); self._vec('M11_B', da.l1_t1_1.bias)
            self._mat('M12_W', da.l1_t2_1.weight, 32); self._vec('M12_B', da.l1_t2_1.bias)
            self._mat('W11', w('l1_t1_2'), 32); self._vec('B11', da.l1_t1_2.bias)
            self._mat('W12', w('l1_t2_2'), 32); self._vec('B12', da.l1_t2_2.bias)
            self._mat('W21A', da.l2_t1_1.weight, 32); self._vec('B21A', da.l2_t1_1.bias)
            self._mat('W22A', da.l2_t2_1.weight, 32); self._vec('B22A', da.l2_t2_1.bias)
            for wv, wc, bc, nm in (('WVA', 'WCA', 'BCA', 'l2_t1_2'), ('WVB', 'WCB', 'BCB', 'l2_t2_2')):
                self._mat(wv, w(nm)[:, 60:90], 16)
                self._mat(wc, torch.cat((w(nm)[:, 0:60], w(nm)[:, 90:95]), dim=1), 16)
                self._vec(bc, getattr(da, nm).bias)
            for pre, m in (('CP', cp), ('CS', cs)):
                self._mat(pre + '_W1', m.fc1.weight, 32); self._vec(pre + '_B1', m.fc1.bias)
                self._mat(pre + '_W2', m.fc2.weight, 16); self._vec(pre + '_B2', m.fc2.bias)
            self._vec('SL', torch.cat([a.weight.detach().reshape(1) for a in (
                sd.activate, ro.activate1, ro.activate2, da.activate, da.activate11, da.activate12, da.activate1,
                da.activate21, da.activate22, da.activate2, cp.activate1, cp.activate2, cs.activate1, cs.activate2)]))
        self._key = key
        return self.buf


def assoc_workspace(plan):
    ws = getattr(plan, '_assoc_workspace', None)
    if ws is None:
        n = int(capi.load().genie_assoc_workspace_bytes(plan.handle))
        ws = torch.empty(max(n, 256), dtype=torch.uint8, device=plan.device)
        plan._assoc_workspace = ws
    return ws


def assoc_product_fwd(plan, packed, x_spatial, y, edge_attr, x_latent, Mask, mask_thresh=0.01, want_parts=False):
    """mask_out, BipartiteGraphReadOutOperator and DataAggregationAssociationPhase (module.py:983-987).  Returns the branch
    output s as a [P,32] view into the association workspace (columns 0-14 and 16-30 hold s[:, :15] and s[:, 15:]; valid
    until the next call on this plan), and with want_parts also s0 [P,15] and mask_out [G]."""
    dev = plan.device
    x_spatial, y = _f32c(x_spatial, 'x_spatial'), _f32c(y, 'y')
    edge_attr, x_latent, Mask = _f32c(edge_attr, 'attr'), _f32c(x_latent, 'x_latent'), _f32c(Mask, 'Mask')
    G, P = plan.n_grid, plan.n_prod
    if x_spatial.shape[0] != G or y.shape[0] != G or x_latent.shape != (P, 30) or Mask.shape != (P, 4) or \
            edge_attr.shape != (P, 3):
        raise capi.GenieError('association: x_spatial/y must have G = %d rows, x_latent [P,30], Mask [P,4], attr [P,3]' % G)
    T = int(y.numel() // max(G, 1))
    ws = assoc_workspace(plan)
    s0 = torch.empty((P, 15), dtype=F32, device=dev) if want_parts else None
    mask_out = torch.empty((G,), dtype=F32, device=dev) if want_parts else None
    ptr = ctypes.c_void_p()
    with torch.cuda.device(dev):
        capi.check(capi.load().genie_assoc_product_fwd(
            plan.handle, capi.dptr(packed, F32), capi.dptr(x_spatial, F32), int(x_spatial.stride(0)), capi.dptr(y, F32), T,
            ctypes.c_float(mask_thresh), capi.dptr(edge_attr, F32), capi.dptr(x_latent, F32), capi.dptr(Mask, F32),
            capi.dptr(ws), capi.dptr(s0), capi.dptr(mask_out), ctypes.byref(ptr), capi.stream_ptr(dev)))
    off = ptr.value - ws.data_ptr()
    s_rows = ws[off:off + P * 32 * 4].view(F32).view(P, 32)
    return (s_rows, s0, mask_out) if want_parts else s_rows


def assoc_collapse_fwd(packed, s_rows, A_edges_p, A_edges_s, dt_partition, tlatent, tpick, ipick, phase_label, n_sta, eps,
                       k_infer=10):
    """LocalSliceLgCollapseP / S (module.py:624-653) for every pick -> arrival [n_arv + 1, 30] (last row: the null arrival)."""
    dev = s_rows.device
    n_arv = int(tpick.shape[0])
    if dt_partition.numel() < 2:
        raise capi.GenieError('dt_partition needs at least two entries')
    dtp = dt_partition.to(dev, F32)
    dt0, dt_step = float(dtp[0]), float(dtp[1] - dtp[0])          # fp32 difference, as module.py:626 forms it
    A_edges_p = A_edges_p.to(dev, torch.int64).contiguous()
    A_edges_s = A_edges_s.to(dev, torch.int64).contiguous()
    if A_edges_p.numel() != A_edges_s.numel():
        raise capi.GenieError('A_edges_p and A_edges_s must have the same length')
    tlatent = _f32c(tlatent.to(dev), 'tlatent')
    if tlatent.shape != (s_rows.shape[0], 2):
        raise capi.GenieError('tlatent must be [P,2]')
    tpick = _f32c(tpick.to(dev).reshape(-1), 'tpick')
    ipick = ipick.to(dev, torch.int64).reshape(-1).contiguous()
    phase = _f32c(phase_label.to(dev).reshape(-1), 'phase_label')
    arrival = torch.empty((n_arv + 1, 30), dtype=F32, device=dev)
    with torch.cuda.device(dev):
        capi.check(capi.load().genie_assoc_collapse_fwd(
            capi.dptr(packed, F32), capi.dptr(s_rows, F32), int(s_rows.shape[0]), capi.dptr(A_edges_p, torch.int64),
            capi.dptr(A_edges_s, torch.int64), int(A_edges_p.numel()), capi.dptr(tlatent, F32),
            capi.dptr(tpick, F32) if n_arv else None, capi.dptr(ipick, torch.int64) if n_arv else None,
            capi.dptr(phase, F32) if n_arv else None, n_arv, int(n_sta), int(dtp.numel()), int(k_infer),
            ctypes.c_float(dt0), ctypes.c_float(dt_step), ctypes.c_float(eps), capi.dptr(arrival), capi.stream_ptr(dev)))
    return arrival


def kron_spmm(kg, csr, x):
    """out[i] = sum_e val[e] * x[nbr(i, e)] over one edge type of the product graph (genie_kron_spmm_fwd): the mean
    aggregation of `propagate` (csr = kg.fwd) or its gradient (csr = kg.rev).  x: fp32 CUDA [P, C]."""
    if x.dtype != F32:
        x = x.float()
    x = x.contiguous()
    if x.dim() != 2 or x.shape[0] != kg.n_prod:
        raise capi.GenieError('kron_spmm: x must be [P, C] with P = %d' % kg.n_prod)
    rowptr, col, val = csr
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        capi.check(capi.load().genie_kron_spmm_fwd(
            int(kg.mode), int(kg.n_sta), int(kg.n_grid), int(kg.n_prod), capi.dptr(rowptr, torch.int64, 'rowptr'),
            capi.dptr(col, torch.int32, 'col'), capi.dptr(val, F32, 'val'), capi.dptr(x, F32, 'x'), int(x.shape[1]),
            int(x.shape[1]), capi.dptr(out), int(x.shape[1]), capi.stream_ptr(x.device)))
    return out


def knn(x, y, k):
    """torch_cluster.knn(x, y, k) on the device (genie_knn_fwd): int64 [n_y, k], the k rows of x nearest to every row of y,
    nearest first.  x, y: fp32 CUDA tensors [n, 3] (kilometres, as the reference passes them)."""
    x, y = _f32c(x, 'x'), _f32c(y, 'y')
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != 3 or y.shape[1] != 3:
        raise capi.GenieError('knn: points must be [n, 3]')
    k = int(min(k, x.shape[0]))
    out = torch.empty((y.shape[0], k), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(capi.load().genie_knn_fwd(capi.dptr(x, F32, 'x'), int(x.shape[0]), capi.dptr(y, F32, 'y') if y.shape[0] else None,
                                             int(y.shape[0]), k, capi.dptr(out) if y.shape[0] else None,
                                             capi.stream_ptr(x.device)))
    return out


def _f32c(t, name):
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


def data_aggregation_fwd(plan, packed, Slice, Mask):
    """DataAggregation.forward (module.py:85-98): Slice, Mask [P,4] -> x_latent [P,30]."""
    Slice, Mask = _f32c(Slice, 'Slice'), _f32c(Mask, 'Mask')
    out = torch.empty((plan.n_prod, 30), dtype=F32, device=plan.device)
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_data_aggregation_fwd(
            plan.handle, capi.dptr(packed, F32), capi.dptr(Slice, F32, 'Slice'), capi.dptr(Mask, F32, 'Mask'),
            capi.dptr(out), capi.dptr(plan.workspace()), capi.stream_ptr(plan.device)))
    return out


def bipartite_readin_fwd(plan, packed, x_latent, edge_attr, Mask):
    """BipartiteGraphOperator.forward (module.py:224-229): -> [G,15]."""
    x_latent, edge_attr, Mask = _f32c(x_latent, 'x'), _f32c(edge_attr, 'attr'), _f32c(Mask, 'Mask')
    out = torch.empty((plan.n_grid, 15), dtype=F32, device=plan.device)
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_bipartite_readin_fwd(
            plan.handle, capi.dptr(packed, F32), capi.dptr(x_latent, F32), capi.dptr(edge_attr, F32),
            capi.dptr(Mask, F32), capi.dptr(out), capi.dptr(plan.workspace()), capi.stream_ptr(plan.device)))
    return out


def spatial_aggregation_fwd(plan, packed, layer, x, pos, scale_rel):
    """SpatialAggregation.forward (module.py:243-249), layer = 0,1,2: [G,C] -> [G,30]."""
    x, pos = _f32c(x, 'x'), _f32c(pos, 'pos')
    if x.shape[1] != (15 if layer == 0 else 30):
        raise capi.GenieError('SpatialAggregation%d expects %d input channels' % (layer + 1, 15 if layer == 0 else 30))
    out = torch.empty((plan.n_grid, 30), dtype=F32, device=plan.device)
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_spatial_aggregation_fwd(
            plan.handle, capi.dptr(packed, F32), int(layer), capi.dptr(x, F32), capi.dptr(pos, F32),
            ctypes.c_float(scale_rel), capi.dptr(out), capi.dptr(plan.workspace()), capi.stream_ptr(plan.device)))
    return out


def frontend_fwd(plan, packed, Slice, Mask, edge_attr, pos, scale_rel, want_latent=False, want_readin=False, out=None):
    """DataAggregation -> Bipartite_ReadIn -> SpatialAggregation1..3 (module.py:1010-1014) in one call."""
    Slice, Mask = _f32c(Slice, 'Slice'), _f32c(Mask, 'Mask')
    edge_attr, pos = _f32c(edge_attr, 'attr'), _f32c(pos, 'pos')
    if Slice.shape != (plan.n_prod, 4) or Mask.shape != (plan.n_prod, 4) or edge_attr.shape != (plan.n_prod, 3):
        raise capi.GenieError('Slice/Mask must be [P,4] and the read-in edge features [P,3] with P = %d' % plan.n_prod)
    if pos.shape != (plan.n_grid, 3):
        raise capi.GenieError('grid positions must be [G,3] with G = %d' % plan.n_grid)
    x_spatial = out if out is not None else torch.empty((plan.n_grid, 30), dtype=F32, device=plan.device)
    latent = torch.empty((plan.n_prod, 30), dtype=F32, device=plan.device) if want_latent else None
    readin = torch.empty((plan.n_grid, 15), dtype=F32, device=plan.device) if want_readin else None
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_frontend_fwd(
            plan.handle, capi.dptr(packed, F32), capi.dptr(Slice, F32), capi.dptr(Mask, F32),
            capi.dptr(edge_attr, F32), capi.dptr(pos, F32), ctypes.c_float(scale_rel), capi.dptr(latent),
            capi.dptr(readin), capi.dptr(x_spatial), capi.dptr(plan.workspace()), capi.stream_ptr(plan.device)))
    return x_spatial, latent, readin


# ---- the front end in two halves (grid-sharded plans, genie_b200/sharded.py) -----------------------------------------------

def da_layer1_fwd(plan, packed, Slice, Mask):
    """DataAggregation up to the layer-2 messages (module.py:88-93); results stay in the plan's workspace."""
    Slice, Mask = _f32c(Slice, 'Slice'), _f32c(Mask, 'Mask')
    if Slice.shape != (plan.n_prod, 4) or Mask.shape != (plan.n_prod, 4):
        raise capi.GenieError('Slice/Mask must be [P,4] with P = %d' % plan.n_prod)
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_da_layer1_fwd(plan.handle, capi.dptr(packed, F32), capi.dptr(Slice, F32),
                                                   capi.dptr(Mask, F32), capi.dptr(plan.workspace()),
                                                   capi.stream_ptr(plan.device)))


def message_rows(plan):
    """View [n_grid, row] of the layer-2 source messages v_b inside the workspace, one row per grid node (n_sta x 16
    channels: fp32, or bf16 carried as uint8 in the bf16 storage mode — the halo exchange only moves the bytes)."""
    ws = plan.workspace()
    ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
    capi.check(capi.load().genie_workspace_region(plan.handle, capi.dptr(ws), 0, ctypes.byref(ptr), ctypes.byref(nbytes)))
    off = ptr.value - ws.data_ptr()
    raw = ws[off:off + nbytes.value]
    if plan.storage == 'bf16':
        return raw.view(plan.n_grid, plan.n_sta * 32)
    return raw.view(F32).view(plan.n_grid, plan.n_sta * 16)


def da_layer2_readin_fwd(plan, packed, Mask, edge_attr, want_latent=False):
    """Rest of DataAggregation (module.py:94-98) + Bipartite_ReadIn (module.py:224-229) for the owned grid nodes."""
    Mask, edge_attr = _f32c(Mask, 'Mask'), _f32c(edge_attr, 'attr')
    n_own = plan.n_grid_owned or plan.n_grid
    out = torch.empty((n_own, 15), dtype=F32, device=plan.device)
    latent = torch.empty((plan.n_prod, 30), dtype=F32, device=plan.device) if want_latent else None
    with torch.cuda.device(plan.device):
        capi.check(capi.load().genie_da_layer2_readin_fwd(plan.handle, capi.dptr(packed, F32), capi.dptr(Mask, F32),
                                                          capi.dptr(edge_attr, F32), capi.dptr(latent), capi.dptr(out),
                                                          capi.dptr(plan.workspace()), capi.stream_ptr(plan.device)))
    return out, latent


# ---- a1 ----------------------------------------------------------------------------------------------------------------

def input_params(t0, max_t, kernel_sig_t, dt, n_locs, n_sta_use, use_sign_input=False):
    """The fp64 scalars the reference derives with numpy (process_utils.py:500-502, 520), by the same expressions."""
    t0, max_t, kernel_sig_t, dt = float(t0), float(max_t), float(kernel_sig_t), float(dt)
    t_offset = 3.0 * kernel_sig_t
    start = t0 - t_offset
    stop = t0 + max_t + t_offset + dt
    prm = capi.InputParams()
    prm.t0, prm.max_t, prm.kernel_sig_t, prm.dt = t0, max_t, kernel_sig_t, dt
    prm.ref0 = start
    prm.ref_step = (start + dt) - start                 # numpy.arange fills start + i*((start+dt)-start)
    prm.n_ts = int(math.ceil((stop - start) / dt))      # len(numpy.arange(start, stop, dt))
    prm.n_extra = int(np.ceil(3 * kernel_sig_t / dt))
    prm.n_locs, prm.n_sta_use = int(n_locs), int(n_sta_use)
    prm.use_sign_input = 1 if use_sign_input else 0
    return prm


def input_scatter_fwd(plan, prm, picks, sta_perm, ind_use, trv_times, node_sta=None, node_grid=None,
                      want_time_bin=False, series=None):
    """extract_input_from_data (process_utils.py:460-629) on the device: picks [n,5] fp64 -> Slice, Mask [P,4]."""
    dev = plan.device
    if picks.dtype != torch.float64:
        raise capi.GenieError('picks must be float64 (pick times lose precision in fp32)')
    picks = picks.contiguous()
    if trv_times.shape[-1] != 2 or trv_times.shape[0] != plan.n_grid or trv_times.shape[1] != prm.n_locs:
        raise capi.GenieError('trv_times must be [G, n_locs, 2]')
    Slice = torch.empty((plan.n_prod, 4), dtype=F32, device=dev)
    Mask = torch.empty((plan.n_prod, 4), dtype=F32, device=dev)
    if series is None:
        series = torch.empty((2, prm.n_sta_use, prm.n_ts), dtype=F32, device=dev)
    tb = torch.empty((plan.n_prod, 2), dtype=torch.int64, device=dev) if want_time_bin else None
    with torch.cuda.device(dev):
        capi.check(capi.load().genie_input_scatter_fwd(
            plan.handle, ctypes.byref(prm), capi.dptr(picks, torch.float64, 'picks'), int(picks.shape[0]),
            capi.dptr(sta_perm, torch.int32, 'sta_perm'), capi.dptr(ind_use, torch.int32, 'ind_use'),
            capi.dptr(_f32c(trv_times, 'trv_times'), F32), capi.dptr(node_sta, torch.int32, 'node_sta'),
            capi.dptr(node_grid, torch.int32, 'node_grid'), capi.dptr(series, F32), capi.dptr(Slice), capi.dptr(Mask),
            capi.dptr(tb), capi.stream_ptr(dev)))
    return Slice, Mask, tb, series


def window_fwd(plan, packed, wp_dev, max_window_picks, n_extra, picks, sta_perm, ind_use, trv_times, series, n_ts_max,
               edge_attr, pos, scale_rel, want_inputs=False, want_latent=False, want_readin=False, out=None):
    """genie_window_fwd: a1 fused into the front end for one window whose parameters sit in the device block `wp_dev`
    (uint8 CUDA tensor holding a capi.WindowParams).  Returns (x_spatial, latent, readin, Slice, Mask)."""
    dev = plan.device
    x_spatial = out if out is not None else torch.empty((plan.n_grid, 30), dtype=F32, device=dev)
    latent = torch.empty((plan.n_prod, 30), dtype=F32, device=dev) if want_latent else None
    readin = torch.empty((plan.n_grid, 15), dtype=F32, device=dev) if want_readin else None
    Slice = torch.empty((plan.n_prod, 4), dtype=F32, device=dev) if want_inputs else None
    Mask = torch.empty((plan.n_prod, 4), dtype=F32, device=dev) if want_inputs else None
    with torch.cuda.device(dev):
        capi.check(capi.load().genie_window_fwd(
            plan.handle, capi.dptr(packed, F32), capi.dptr(wp_dev, torch.uint8, 'window params'), int(max_window_picks),
            int(n_extra), capi.dptr(picks, torch.float64, 'picks') if max_window_picks else None,
            capi.dptr(sta_perm, torch.int32, 'sta_perm'), capi.dptr(ind_use, torch.int32, 'ind_use'),
            capi.dptr(trv_times, F32, 'trv_times'), capi.dptr(series, F32, 'series'), int(n_ts_max),
            capi.dptr(edge_attr, F32, 'attr'), capi.dptr(pos, F32, 'pos'), ctypes.c_float(scale_rel), capi.dptr(Slice),
            capi.dptr(Mask), capi.dptr(latent), capi.dptr(readin), capi.dptr(x_spatial), capi.dptr(plan.workspace()),
            capi.stream_ptr(dev)))
    return x_spatial, latent, readin, Slice, Mask


# ---- per-node dense layers of the training path (genie_node_mlp_fwd / genie_node_mlp_bwd) -------------------------------------

def _mlp_desc(parts, weight, bias, slope):
    n = int(parts[0].shape[0])
    if not (1 <= len(parts) <= 4):
        raise capi.GenieError('node_mlp: 1..4 input parts')
    d = capi.MlpDesc()
    d.n_rows, d.n_parts, d.n_out = n, len(parts), int(weight.shape[0])
    n_in = 0
    for i, x in enumerate(parts):
        if x.dim() != 2 or x.shape[0] != n or x.dtype != F32 or x.stride(1) != 1:
            raise capi.GenieError('node_mlp: parts must be fp32 [n, w] tensors with unit column stride')
        d.width[i], d.ld[i] = int(x.shape[1]), int(x.stride(0)) if n > 1 else int(x.shape[1])
        d.x[i] = capi.dptr(x, F32, 'part %d' % i) if n else None
        n_in += int(x.shape[1])
    if tuple(weight.shape) != (d.n_out, n_in) or not weight.is_contiguous():
        raise capi.GenieError('node_mlp: weight must be contiguous [n_out, %d]' % n_in)
    d.weight = capi.dptr(weight, F32, 'weight')
    d.bias = capi.dptr(bias, F32, 'bias') if bias is not None else None
    d.slope = capi.dptr(slope, F32, 'slope') if slope is not None else None
    return d, n_in


def node_mlp_supported(parts, weight):
    return len(parts) <= 4 and weight.shape[0] <= 32 and sum(int(x.shape[1]) for x in parts) <= 104


def node_mlp_fwd(parts, weight, bias, slope, want_mask=True):
    """y = PReLU_slope(Linear([parts...])) without materialising the concatenation (slope None: no activation); also the
    per-node sign bits of the pre-activation that node_mlp_bwd needs (int32 [n], None without a slope)."""
    d, _ = _mlp_desc(parts, weight, bias, slope)
    y = torch.empty((d.n_rows, d.n_out), dtype=F32, device=weight.device)
    neg = torch.empty((d.n_rows,), dtype=torch.int32, device=weight.device) if (slope is not None and want_mask) else None
    with torch.cuda.device(weight.device):
        capi.check(capi.load().genie_node_mlp_fwd(ctypes.byref(d), capi.dptr(y), d.n_out, capi.dptr(neg),
                                                  capi.stream_ptr(weight.device)))
    return y, neg


def node_mlp_bwd(parts, weight, bias, slope, y, neg, gy, need_gx):
    """Gradients of node_mlp_fwd: (gx per part or None, gW, gb, gslope)."""
    d, n_in = _mlp_desc(parts, weight, bias, slope)
    dev = weight.device
    gy = gy.contiguous()
    lib = capi.load()
    with torch.cuda.device(dev):
        rows = int(lib.genie_node_mlp_partial_rows())
    pld = d.n_out * n_in + d.n_out + 1
    partial = torch.empty((rows, pld), dtype=F32, device=dev)
    gx = [torch.empty_like(x, memory_format=torch.contiguous_format) if (need and d.n_rows) else None
          for x, need in zip(parts, need_gx)]
    ptrs = (ctypes.c_void_p * 4)(*[(capi.dptr(g) if g is not None else None) for g in gx] + [None] * (4 - len(gx)))
    lds = (ctypes.c_int32 * 4)(*[(int(g.shape[1]) if g is not None else 0) for g in gx] + [0] * (4 - len(gx)))
    with torch.cuda.device(dev):
        capi.check(lib.genie_node_mlp_bwd(ctypes.byref(d), capi.dptr(y, F32) if slope is not None else None, d.n_out,
                                          capi.dptr(neg, torch.int32, 'neg_mask') if slope is not None else None,
                                          capi.dptr(gy, F32, 'gy'), d.n_out, ptrs, lds, capi.dptr(partial), capi.stream_ptr(dev)))
    tot = partial.sum(0)
    gW = tot[:d.n_out * n_in].view(d.n_out, n_in)
    gb = tot[d.n_out * n_in:d.n_out * n_in + d.n_out]
    ga = tot[d.n_out * n_in + d.n_out:]
    gx = [g if g is not None else (torch.zeros_like(x) if need else None) for g, x, need in zip(gx, parts, need_gx)]
    return gx, gW, gb, ga


def csr_apply(csr, x, n_out):
    """out[i, :] = sum_{e in row i} val[e] * x[col[e], :]  for an explicit CSR (rowptr int64 [n_out+1], col int32, val fp32):
    genie_kron_spmm_fwd in explicit mode with independent input / output row counts.  Gathers (one entry per row), segment
    means (val = 1 / row length) and their transposes are all instances; no atomics, bit-reproducible."""
    rowptr, col, val = csr
    if x.dtype != F32:
        x = x.float()
    x = x.contiguous()
    out = torch.empty((int(n_out), x.shape[1]), dtype=F32, device=x.device)
    if n_out == 0 or x.shape[1] == 0:
        return out
    with torch.cuda.device(x.device):
        capi.check(capi.load().genie_kron_spmm_fwd(
            2, 0, 0, int(n_out), capi.dptr(rowptr, torch.int64, 'rowptr'), capi.dptr(col, torch.int32, 'col'),
            capi.dptr(val, F32, 'val'), capi.dptr(x, F32, 'x') if x.shape[0] else capi.dptr(out), int(x.shape[1]), int(x.shape[1]),
            capi.dptr(out), int(x.shape[1]), capi.stream_ptr(x.device)))
    return out
