"""Zero-copy torch.Tensor front end of the `_dev` entry points (BASELINE.json north_star: "thin PyTorch extension over
a C-ABI").  PyTorch is plumbing here: it owns device memory and streams, every number is computed by
libcsplotch_b200.so through include/csplotch.h (`csplotch_batch_create_dev`, `csplotch_log_prob_grad_dev`); tensors are
handed over as raw device pointers on torch's current stream, nothing is copied through the host.

    model = api.Model(shared, "comp_splotch_stan_model")
    batch = torch_api.batch_from_tensor(model, counts_cuda_int32, pm, ps)       # counts stay in HBM
    lp, grad = torch_api.log_prob_grad(batch, theta_cuda_f64, gene_of_cuda_int32)

Reference call sites these stand in for: the per-gene `data_<g>.R` hand-over of /root/reference/bin/splotch:129-133
(counts) and the stanc-generated `log_prob` / `grad_log_prob` that CmdStan's sampler calls."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api
from ._lib import CsplotchError, check


def _torch():
    import torch
    return torch


def _check_cuda(t, dtype, name):
    torch = _torch()
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise CsplotchError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise CsplotchError("%s must have dtype %s" % (name, dtype))
    return t.contiguous()


def batch_from_tensor(model: api.Model, counts, beta_prior_mean=None, beta_prior_std=None, gene_id0: int = 0,
                      gene_ids=None) -> api.Batch:
    """counts: CUDA int32 tensor (G, N) already in HBM; priors: host arrays (G, K) for comp_splotch_stan_model."""
    torch = _torch()
    c = _check_cuda(counts, torch.int32, "counts").reshape(-1, model.N)
    G = int(c.shape[0])
    pm = ps = None
    if model.family == 1:
        if beta_prior_mean is None or beta_prior_std is None:
            raise CsplotchError("comp_splotch_stan_model needs beta_prior_mean / beta_prior_std")
        pm = np.ascontiguousarray(np.asarray(beta_prior_mean, dtype=np.float64).reshape(G, model.K))
        ps = np.ascontiguousarray(np.asarray(beta_prior_std, dtype=np.float64).reshape(G, model.K))
    b = api.Batch.__new__(api.Batch)
    b.model, b.lib, b.G = model, model.lib, G
    b.beta_prior_mean, b.beta_prior_std = pm, ps
    h = C.c_void_p()
    stream = torch.cuda.current_stream(c.device).cuda_stream
    check(model.lib.csplotch_batch_create_dev(model.h, C.c_void_p(c.data_ptr()),
                                              pm.ctypes.data_as(C.c_void_p) if pm is not None else None,
                                              ps.ctypes.data_as(C.c_void_p) if ps is not None else None, C.c_int32(G),
                                              C.c_uint32(gene_id0), C.c_void_p(stream), C.byref(h)))
    b.h = h
    b.gene_id0 = gene_id0
    b.gene_ids = np.arange(G, dtype=np.uint32) + np.uint32(gene_id0)
    if gene_ids is not None:
        ids = np.ascontiguousarray(np.asarray(gene_ids, dtype=np.uint32).reshape(-1))
        if ids.size != G:
            raise CsplotchError("gene_ids must have one entry per gene")
        check(model.lib.csplotch_batch_set_gene_ids(b.h, ids.ctypes.data_as(C.c_void_p)))
        b.gene_ids = ids
    return b


def log_prob_grad(batch: api.Batch, theta, gene_of=None, jacobian: bool = True, want_grad: bool = True):
    """theta: CUDA float64 (B, dim); gene_of: CUDA int32 (B,) or None (b % G).  Returns CUDA tensors lp (B,) and
    grad (B, dim) (None without want_grad), computed on torch's current stream."""
    torch = _torch()
    m = batch.model
    th = _check_cuda(theta, torch.float64, "theta").reshape(-1, m.dim)
    Bn = int(th.shape[0])
    go = _check_cuda(gene_of, torch.int32, "gene_of").reshape(-1) if gene_of is not None else None
    if go is not None and int(go.numel()) != Bn:
        raise CsplotchError("gene_of must have one entry per row of theta")
    lp = torch.empty(Bn, dtype=torch.float64, device=th.device)
    grad = torch.empty((Bn, m.dim), dtype=torch.float64, device=th.device) if want_grad else None
    stream = torch.cuda.current_stream(th.device).cuda_stream
    check(batch.lib.csplotch_log_prob_grad_dev(batch.h, C.c_void_p(th.data_ptr()),
                                               C.c_void_p(go.data_ptr()) if go is not None else None, C.c_int32(Bn),
                                               C.c_void_p(lp.data_ptr()),
                                               C.c_void_p(grad.data_ptr()) if grad is not None else None,
                                               C.c_int32(1 if jacobian else 0), C.c_void_p(stream)))
    return lp, grad
