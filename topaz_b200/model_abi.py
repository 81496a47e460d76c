"""Python caller of the model-level C ABI (csrc/tpz_model.cu): a filled LinearClassifier as ONE native handle.

``DenseModel(model)`` describes the network to ``tpz_model_create`` (layer list + DEVICE parameter pointers in the
reference's own OIHW layout); plan building, eval-mode BatchNorm folding and the fp16 repack happen in the library, on the
device.  ``engine.classifier_forward`` routes the default-precision dense forward through it (TPZ_DENSE_ENGINE=py keeps the
Python-built plans, which tests/test_gpu_model_abi.py holds bit-identical to this path)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import TpzLayerDesc, check


class WeightRangeError(RuntimeError):
    """a (BatchNorm-folded) weight row does not fit fp16 without the per-row scale only the Python packer applies"""


E_WEIGHT_RANGE = 4      # include/topaz_b200.h TPZ_E_WEIGHT_RANGE


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t):
    """contiguous fp32 view/copy of a parameter ON ITS DEVICE (no host round trip)"""
    return t.detach().to(torch.float32).contiguous() if t is not None else None


class DenseModel:
    def __init__(self, model: nn.Module):
        from . import engine
        feats = model.features
        if not engine.is_filled(feats):
            raise RuntimeError('topaz_b200: DenseModel needs a filled model (model.fill())')
        self.model = model
        self.handle = C.c_void_p()
        self.pad = feats.width // 2
        self._ws = None
        self._build(create=True)

    def _describe(self):
        """(TpzLayerDesc array, keep-alive tensors, cls_w, cls_b) for the model's CURRENT parameters"""
        from . import engine
        feats, cls = self.model.features, self.model.classifier
        blocks = engine._feature_blocks(feats)
        if feats.training and any(b.get('bn') is not None or b.get('bn0') is not None for b in blocks):
            raise NotImplementedError('topaz_b200: dense forward with BatchNorm requires eval() mode')
        keep = []

        def P(t):
            t = _f32(t)
            if t is not None:
                keep.append(t)
            return t.data_ptr() if t is not None else None

        def BN(bn):
            if bn is None:
                return None, 0.0
            t = torch.stack([bn.weight.detach().float(), bn.bias.detach().float(), bn.running_mean.detach().float(),
                             bn.running_var.detach().float()]).contiguous()
            keep.append(t)
            return t.data_ptr(), float(bn.eps)
        descs = (TpzLayerDesc * len(blocks))()
        for d, b in zip(descs, blocks):
            if b['kind'] == 'conv':
                w = b['w']
                d.kind, d.cin, d.cout, d.k = 0, w.shape[1], w.shape[0], w.shape[-1]
                d.dil0, d.dil1, d.slope0, d.slope1 = b['dil'], 1, b['slope'], 0.0
                d.w0, d.b0 = P(w), P(b['b'])
                d.bn0, d.eps0 = BN(b['bn'])
                if b['stride'] != 1:
                    raise RuntimeError('topaz_b200: dense plan requested on an unfilled (strided) model')
            elif b['kind'] == 'resid':
                w0, w1 = b['w0'], b['w1']
                d.kind, d.cin, d.cout, d.k = 1, w0.shape[0], w1.shape[0], 3
                d.dil0, d.dil1, d.slope0, d.slope1 = b['d0'], b['d1'], b['slope0'], b['slope1']
                d.w0, d.b0, d.w1, d.b1, d.proj = P(w0), P(b['b0']), P(w1), P(b['b1']), P(b['proj'])
                d.bn0, d.eps0 = BN(b['bn0'])
                d.bn1, d.eps1 = BN(b['bn1'])
                if b['stride'] != 1:
                    raise RuntimeError('topaz_b200: dense plan requested on an unfilled (strided) model')
            else:
                raise NotImplementedError(f"topaz_b200: layer kind {b['kind']} in the dense C model")
        cw, cb = _f32(cls.weight).reshape(-1), _f32(cls.bias).reshape(-1)
        keep += [cw, cb]
        # kernels per forward: input range scale + first layer + one per conv step (a ResidA block is two)
        self.n_launch = 2 + sum(2 if b['kind'] == 'resid' else 1 for b in blocks[1:])
        return descs, keep, cw, cb

    def _build(self, create: bool):
        descs, keep, cw, cb = self._describe()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if create:
            rc = _lib.lib().tpz_model_create(descs, len(descs), _ptr(cw), _ptr(cb), self.pad, C.byref(self.handle), s)
        else:
            rc = _lib.lib().tpz_model_update_weights(self.handle, descs, len(descs), _ptr(cw), _ptr(cb), s)
        if rc == E_WEIGHT_RANGE:
            raise WeightRangeError(_lib.lib().tpz_last_error().decode())
        check(rc)
        ops._count(2 * len(descs) + 4)
        del keep          # the library has read (and synchronised on) every parameter

    def update(self):
        """repack after the parameters changed (same architecture)"""
        self._build(create=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: fp32 [B, H, W] contiguous on the device -> logits [B, 1, H, W]"""
        B, H, W = x.shape
        lib = _lib.lib()
        need = lib.tpz_workspace_bytes(self.handle, B, H, W)
        if need < 0:
            raise RuntimeError('topaz_b200: image smaller than the receptive field allows')
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        y = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
        check(lib.tpz_resnet_dense_forward(self.handle, _ptr(x), B, H, W, _ptr(y), _ptr(self._ws), need,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ops._count(self.n_launch)
        return y

    def step_buffers(self, step: int):
        """(fp16 weights [nkb, Co, KC] or None, fp32 bias [Co]) of conv step `step`, copied out of the handle (test hook;
        step -1 = the first layer)."""
        n, co, kc, nkb = C.c_longlong(), C.c_int(), C.c_int(), C.c_int()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        fn = _lib.lib().tpz_model_step_buffers
        check(fn(self.handle, step, None, 0, None, C.byref(n), C.byref(co), C.byref(kc), C.byref(nkb), s))
        dev = next(self.model.parameters()).device
        wt = torch.empty(int(n.value), dtype=torch.float16, device=dev) if n.value else None
        bt = torch.empty(int(co.value), dtype=torch.float32, device=dev)
        check(fn(self.handle, step, _ptr(wt), int(n.value), _ptr(bt), None, None, None, None, s))
        return (wt.view(int(nkb.value), int(co.value), int(kc.value)) if wt is not None else None), bt

    def timing(self, enable: bool):
        """record CUDA events around the last conv step (+ fused classifier) of every forward"""
        check(_lib.lib().tpz_model_timing(self.handle, int(enable)))

    def timing_read(self):
        """durations (ms) of the last conv step of the forwards since the previous read (at most 64), oldest first"""
        buf = (C.c_float * 64)()
        n = C.c_int()
        check(_lib.lib().tpz_model_timing_read(self.handle, buf, 64, C.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def close(self):
        if self.handle:
            _lib.lib().tpz_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
