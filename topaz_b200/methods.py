"""Drop-in mirror of topaz.methods.GE_binomial (reference methods.py:77-165): one training step =
strided forward, fused GE-binomial loss + closed-form score gradient, backward, (multi-GPU: gradient
all-reduce), fused Adam + gradient zeroing.  Returns the same 5-tuple of Python floats.

Multi-GPU (one process per GPU, torch.distributed / NCCL): every rank holds the same parameters and a
shard of the minibatch.  The GE term needs minibatch-global statistics (q_mu, q_var, N, the positives'
mean), so ranks all-gather their logits + labels (B floats), evaluate the loss on the global vector and
back-propagate only their own shard; the summed gradient (one all-reduce of the flat buffer) equals the
single-process gradient."""
from __future__ import absolute_import, print_function, division

import torch
import torch.nn as nn

from topaz_b200 import ops, train_engine


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class _Objective:
    """Shared step skeleton: taped strided forward, fused loss kernel, backward, (DP) gradient all-reduce, fused Adam."""
    n_out = 6          # floats written by the loss kernel (GE_binomial: 5)

    def _setup(self, model, optim, criteria):
        if not isinstance(criteria, nn.BCEWithLogitsLoss):
            raise NotImplementedError('topaz_b200: objectives expect criteria = nn.BCEWithLogitsLoss() (training.py:377)')
        if not isinstance(optim, torch.optim.Adam):
            raise NotImplementedError('topaz_b200: objectives expect torch.optim.Adam (training.py:355-356)')
        self.model, self.optim, self.criteria = model, optim, criteria
        self._out = None           # device floats written by the loss kernel
        self._host = None          # their pinned host mirror (the step's single read-back)

    def _hyper(self):
        g = self.optim.param_groups[0]
        if g.get('weight_decay', 0) != 0 or g.get('amsgrad', False):
            raise NotImplementedError('topaz_b200: Adam weight_decay / amsgrad are not supported')
        return float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps'])

    def _sync_optim_state(self, fp):
        """Keep optim.state inspectable/saveable: exp_avg / exp_avg_sq alias the flat moment buffers."""
        st = self.optim.state
        for p, off in zip(fp.params, fp.offsets):
            s = st[p]
            if 'exp_avg' not in s or s['exp_avg'].data_ptr() != fp.flat_m[off:].data_ptr():
                k = p.numel()
                s['exp_avg'] = fp.flat_m[off:off + k].view_as(p)
                s['exp_avg_sq'] = fp.flat_v[off:off + k].view_as(p)
            s['step'] = torch.tensor(float(fp.step))

    def _loss(self, gs, gy, lo, hi, dscore, out):
        raise NotImplementedError

    # ---- one training step = device-side body (forward, collectives, loss, backward, Adam) + one host read-back ----
    def _step_body(self, X, Y):
        """Everything of a step that runs on the device; writes the loss / metric floats to self._out.  Free of host
        synchronisation and of step-dependent host scalars, so it can be captured into a CUDA graph."""
        model = self.model
        fp = train_engine.flat_params(model)
        score = model(X).view(-1)
        Yd = Y.to(device=score.device, dtype=torch.float64).view(-1)
        b = score.numel()
        dist = _dist()
        if dist is not None:
            world, rank = dist.get_world_size(), dist.get_rank()
            if self.__dict__.get('_checked_b') != b:
                # the gather below (and the global-minibatch BatchNorm count) assume the same shard size on every rank; an
                # uneven last minibatch would otherwise hang or corrupt the gather.  Checked once per shard size.
                sizes = [None] * world
                dist.all_gather_object(sizes, int(b))
                if any(v != b for v in sizes):
                    raise ValueError(f'topaz_b200: data-parallel step needs equal minibatch shards on all ranks, got {sizes}')
                self._checked_b = b
            # ONE all-gather for logits + labels: each rank contributes [b logits | b fp64 labels viewed as 2b floats]
            mine = torch.empty(3 * b, dtype=torch.float32, device=score.device)
            mine[:b] = score
            mine[b:].view(torch.float64).copy_(Yd)
            allb = torch.empty(world * 3 * b, dtype=torch.float32, device=score.device)
            dist.all_gather_into_tensor(allb, mine)
            allb = allb.view(world, 3 * b)
            gs = allb[:, :b].reshape(-1)
            gy = allb[:, b:].contiguous().view(torch.float64).reshape(-1)
            lo, hi = rank * b, (rank + 1) * b
        else:
            gs, gy, lo, hi = score.contiguous(), Yd.contiguous(), 0, b
        if self._out is None or self._out.device != score.device:
            self._out = torch.empty(self.n_out, dtype=torch.float32, device=score.device)
            self._host = torch.empty(self.n_out, dtype=torch.float32)
            if score.is_cuda:
                self._host = self._host.pin_memory()
        dscore = torch.empty(b, dtype=torch.float32, device=score.device)
        self._loss(gs, gy, lo, hi, dscore, self._out)
        if dist is None:
            train_engine.backward(model, dscore)
        else:
            # the flat gradient is laid out in forward order and produced back to front: all-reduce each finished suffix
            # bucket asynchronously while the earlier layers are still back-propagating; Adam waits for all of them
            works, done = [], [fp.n]

            def suffix(off, final=False):
                if off < done[0] and (final or done[0] - off >= self.bucket_elems):
                    works.append(dist.all_reduce(fp.flat_g[off:done[0]], op=dist.ReduceOp.SUM, async_op=True))
                    done[0] = off
            train_engine.backward(model, dscore, on_suffix_done=suffix)
            suffix(0, final=True)
            for w in works:
                w.wait()
        lr, b1, b2, eps = self._hyper()
        train_engine.adam_step(fp, lr, b1, b2, eps, self.l2)
        return fp

    bucket_elems = 1 << 16          # all-reduce buckets of >= 256 KB (the collective is latency-bound below that)

    def _graph_key(self, X, Y, fp):
        return (tuple(X.shape), X.dtype, tuple(Y.shape), Y.dtype, str(X.device), id(fp), self._hyper(), float(self.l2),
                self._loss_key(), _dist() is not None, tuple(m.training for m in self.model.modules()))

    def _loss_key(self):
        return tuple(float(v) if isinstance(v, (int, float)) else v for v in
                     (getattr(self, k, None) for k in ('pi', 'slack', 'momentum', 'beta')))

    def _graph_ok(self, X):
        """CUDA-graph replay of the step: on by default, data-parallel steps included -- the NCCL all-gather and the bucketed
        gradient all-reduces are captured with the kernels (TPZ_TRAIN_GRAPH=0 disables; =nodp keeps data-parallel steps
        eager).  Not with active dropout (its Philox offset advances on the host) nor with an objective whose loss arguments
        change from step to step (GE_KL with momentum < 1)."""
        import os
        mode = os.environ.get('TPZ_TRAIN_GRAPH', '1')
        if mode == '0' or not X.is_cuda:
            return False
        if _dist() is not None and (mode == 'nodp' or _dist().get_backend() != 'nccl'):
            return False
        if getattr(self, 'momentum', 1.0) < 1:
            return False
        return not any(isinstance(m, nn.Dropout) and m.training and m.p > 0 for m in self.model.modules())

    def _run(self, X, Y):
        ops.require_cuda(X, 'training minibatch')
        model = self.model
        if not model.training:
            model.train()
        fp = train_engine.flat_params(model)
        st = self.__dict__.setdefault('_graph_state', dict(key=None, seen=0, graph=None))
        key = self._graph_key(X, Y, fp) if self._graph_ok(X) else None
        if key is None or key != st['key']:
            st.update(key=key, seen=0, graph=None)          # new shapes / parameters / hyper-parameters: start over (eagerly)
        if key is not None and st['graph'] is None and st['seen'] >= 2:
            # third step with this signature: capture.  The two eager steps built the flat buffers, kernel attributes and
            # NCCL communicators; the capture records one more real step on static copies of the minibatch.
            try:
                sx, sy = X.clone(), Y.clone()
                l0 = ops.LAUNCH_COUNT
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body(sx, sy)
                fp.step -= 1                                  # capture does not execute: undo the host mirror's increment
                st.update(graph=g, sx=sx, sy=sy, launches=ops.LAUNCH_COUNT - l0)
            except Exception as e:
                import sys
                print(f'topaz_b200: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); running eagerly',
                      file=sys.stderr)
                torch.cuda.synchronize()
                st.update(key=('failed',), graph=None)
                key = None
        if key is not None and st['graph'] is not None:
            st['sx'].copy_(X, non_blocking=True); st['sy'].copy_(Y, non_blocking=True)
            st['graph'].replay()
            fp.step += 1
            ops._count(st['launches'])
        else:
            self._step_body(X, Y)
            st['seen'] += 1
        model.__dict__['_tpz_epoch'] = model.features.__dict__['_tpz_epoch'] = fp.step
        self._sync_optim_state(fp)
        return [float(v) for v in train_engine.read_back(self._out, self._host)]


class PN(_Objective):
    """reference methods.py:25-74 (autoencoder = 0)."""
    def __init__(self, model, optim, criteria, pi=None, l2=0, autoencoder=0):
        if autoencoder > 0:
            raise NotImplementedError('topaz_b200: autoencoder is outside the B200 hot path')
        self._setup(model, optim, criteria)
        self.pi, self.l2, self.autoencoder = pi, l2, autoencoder
        self.header = ['loss', 'precision', 'adjusted_precision', 'tpr', 'fpr']

    def _loss(self, gs, gy, lo, hi, dscore, out):
        train_engine.pu_objective_loss_grad(gs, gy, 0, self.pi if self.pi is not None else -1.0, 1.0, 1.0, 0.0, lo, hi, dscore, out)

    def step(self, X, Y):
        o = self._run(X, Y)
        return (o[0], o[2], o[3], o[4])


class GE_KL(_Objective):
    """reference methods.py:168-255 (entropy_penalty = 0)."""
    def __init__(self, model, optim, criteria, pi, l2=0, slack=1.0, momentum=1.0, entropy_penalty=0):
        if entropy_penalty > 0:
            raise NotImplementedError('topaz_b200: entropy_penalty is outside the B200 hot path')
        self._setup(model, optim, criteria)
        self.pi, self.l2, self.slack, self.momentum = pi, l2, slack, momentum
        self.running_expectation = pi
        self.entropy_penalty = entropy_penalty
        self.header = ['loss', 'ge_penalty', 'precision', 'adjusted_precision', 'tpr', 'fpr']

    def _loss(self, gs, gy, lo, hi, dscore, out):
        train_engine.pu_objective_loss_grad(gs, gy, 1, self.pi, self.slack, self.momentum, self.running_expectation, lo, hi, dscore, out)

    def step(self, X, Y):
        o = self._run(X, Y)
        if self.momentum < 1:
            self.running_expectation = o[5]
        return o[0], o[1], o[2], o[3], o[4]


class PU(_Objective):
    """reference methods.py:258-322 (non-negative PU risk; autoencoder = 0)."""
    def __init__(self, model, optim, criteria, pi, l2=0, beta=0.0, autoencoder=0):
        if autoencoder > 0:
            raise NotImplementedError('topaz_b200: autoencoder is outside the B200 hot path')
        self._setup(model, optim, criteria)
        self.pi, self.l2, self.beta, self.autoencoder = pi, l2, beta, autoencoder
        self.header = ['loss', 'precision', 'adjusted_precision', 'tpr', 'fpr']

    def _loss(self, gs, gy, lo, hi, dscore, out):
        train_engine.pu_objective_loss_grad(gs, gy, 2, self.pi, 1.0, 1.0, self.beta, lo, hi, dscore, out)

    def step(self, X, Y):
        o = self._run(X, Y)
        return (o[0], o[2], o[3], o[4])


class GE_binomial(_Objective):
    """reference methods.py:77-165 (entropy_penalty = autoencoder = posterior_L1 = 0)."""
    n_out = 5

    def __init__(self, model, optim, criteria, pi, l2=0, slack=1.0, entropy_penalty=0, autoencoder=0, posterior_L1=0):
        if entropy_penalty > 0 or autoencoder > 0 or posterior_L1 > 0:
            raise NotImplementedError('topaz_b200: entropy_penalty / autoencoder / posterior_L1 are outside the B200 hot path')
        self._setup(model, optim, criteria)
        self.slack = slack
        self.pi = pi
        self.entropy_penalty = entropy_penalty
        self.l2 = l2
        self.autoencoder = autoencoder
        self.posterior_L1 = posterior_L1
        self.header = ['loss', 'ge_penalty', 'precision', 'adjusted_precision', 'tpr', 'fpr']

    def _loss(self, gs, gy, lo, hi, dscore, out):
        train_engine.ge_loss_grad(gs, gy, self.pi, self.slack, lo, hi, dscore, out)

    def step(self, X, Y):
        """One training step (methods.py:98-165): taped strided forward (:103), fused loss + closed-form score gradient
        (:105-136), backward (:146), flat-gradient all-reduce under data parallelism, fused Adam + gradient zeroing
        (:153-160).  Returns (classifier_loss, ge_penalty, precision, tpr, fpr) as Python floats."""
        cls_loss, ge, precision, tpr, fpr = self._run(X, Y)[:5]
        return cls_loss, ge, precision, tpr, fpr
