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


class UnetModel:
    """A U-Net denoiser (UDenoiseNet / UDenoiseNetSmall / UDenoiseNet3D drop-in) as one native handle (csrc/tpz_unet.cu):
    ``tpz_unet_create`` reads the convolutions in the reference's own layout and builds every plan in C++;
    ``forward`` is one ``tpz_unet2d_forward`` / ``tpz_unet3d_forward`` call on a caller-owned workspace.
    ``engine.unet_forward`` routes through it with TPZ_UNET_ENGINE=c (``precision`` = the engine's fast / auto / strict); the
    Python-built plans stay the default path until this one has been run on hardware.  ``host=True`` builds a test handle on HOST tensors that runs only
    under the launch hook (tests/test_unet_abi.py)."""

    PRECISIONS = {'fast': 0, 'auto': 1, 'strict': 2}      # include/topaz_b200.h TPZ_PRECISION_*

    def __init__(self, model: nn.Module, host: bool = False, precision: str = 'fast'):
        self.model = model
        self.host = host
        self.handle = C.c_void_p()
        self._ws = None
        desc, keep = self.describe(model, host)
        desc.precision = self.PRECISIONS[precision]
        self.dims, self.depth = desc.dims, desc.depth
        s = None if host else C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = _lib.lib().tpz_unet_create(C.byref(desc), C.byref(self.handle), s)
        if rc == E_WEIGHT_RANGE:
            raise WeightRangeError(_lib.lib().tpz_last_error().decode())
        check(rc)
        del keep          # the library has copied every parameter

    @staticmethod
    def describe(model: nn.Module, host: bool = False):
        """(TpzUnetDesc, keep-alive tensors) for the model's current parameters (enc{i}.0, dec{l}.{0,2}, dec1.4)"""
        enc = [getattr(model, f'enc{i}') for i in range(1, 10) if hasattr(model, f'enc{i}')]
        depth = len(enc)
        dims = 3 if isinstance(enc[0][0], nn.Conv3d) else 2
        if not 2 <= depth <= _lib.TPZ_UNET_MAX_DEPTH:
            raise NotImplementedError(f'topaz_b200: U-Net depth {depth} in the C model')
        for i, e in enumerate(enc, 1):
            pooled = any(isinstance(c, (nn.MaxPool2d, nn.MaxPool3d)) for c in e)
            if pooled != (i < depth):
                raise NotImplementedError('topaz_b200: the C U-Net expects MaxPool(2) after every encoder stage but the last')
        keep = []
        desc = _lib.TpzUnetDesc()
        desc.dims, desc.depth, desc.slope, desc.host_weights = dims, depth, 0.1, int(host)

        def fill(d, conv):
            w = _f32(conv.weight)
            b = _f32(conv.bias)
            if host:
                w, b = w.cpu(), (b.cpu() if b is not None else None)
            keep.extend([w, b])
            if len(set(w.shape[2:])) != 1:
                raise NotImplementedError('topaz_b200: the C U-Net expects cubic kernels')
            d.w, d.b = w.data_ptr(), (b.data_ptr() if b is not None else None)
            d.cout, d.cin, d.k = w.shape[0], w.shape[1], w.shape[-1]
        for i, e in enumerate(enc):
            fill(desc.enc[i], e[0])
        for l in range(depth - 1, 0, -1):
            dec = getattr(model, f'dec{l}')
            fill(desc.dec_a[l], dec[0])
            fill(desc.dec_b[l], dec[2])
        fill(desc.last, model.dec1[4])
        return desc, keep

    def workspace_bytes(self, shape) -> int:
        N, D, H, W = shape
        need = _lib.lib().tpz_unet_workspace_bytes(self.handle, N, D, H, W)
        if need < 0:
            raise RuntimeError('topaz_b200: patch smaller than the U-Net pooling stages allow')
        return int(need)

    def launch_count(self, shape) -> int:
        N, D, H, W = shape
        return int(_lib.lib().tpz_unet_launch_count(self.handle, N, D, H, W))

    def forward(self, x: torch.Tensor, denorm_stats=None, workspace=None) -> torch.Tensor:
        """x: fp32 [N, 1, (D,) H, W] contiguous (device; host for a test handle) -> same shape"""
        xi = x[:, 0].contiguous().float()
        shape = (xi.shape[0], 1, xi.shape[1], xi.shape[2]) if self.dims == 2 else tuple(xi.shape)
        N, D, H, W = shape
        need = self.workspace_bytes(shape)
        # under CUDA-graph capture the workspace must belong to the graph's own pool (its address is baked into the captured
        # launches; a cached buffer could be re-sized -- freed -- by a later, larger patch shape while the graph still replays)
        capturing = x.is_cuda and torch.cuda.is_current_stream_capturing()
        ws = workspace if workspace is not None else (None if capturing else self._ws)
        if ws is None or ws.numel() < need or ws.device != x.device:
            raw = torch.empty(need + 256, dtype=torch.uint8, device=x.device)
            off = (-raw.data_ptr()) % 256          # the library asks for 256-byte alignment (host allocations give 64)
            ws = raw[off:off + need]
            if workspace is None and not capturing:
                self._ws = ws
        y = torch.empty_like(xi)
        lib = _lib.lib()
        s = None if self.host else C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.dims == 2:
            check(lib.tpz_unet2d_forward(self.handle, _ptr(xi), N, H, W, _ptr(denorm_stats), _ptr(y), _ptr(ws), ws.numel(), s))
        else:
            check(lib.tpz_unet3d_forward(self.handle, _ptr(xi), N, D, H, W, _ptr(denorm_stats), _ptr(y), _ptr(ws), ws.numel(), s))
        ops._count(self.launch_count(shape))
        return y[:, None]

    def plan(self, which: int, index: int = 0, phase: int = 0):
        """(static TpzTcConvArgs, packed weight element count) of one plan (test hook, see tpz_unet_plan)"""
        a = _lib.TpzTcConvArgs()
        n = C.c_longlong()
        check(_lib.lib().tpz_unet_plan(self.handle, which, index, phase, C.byref(a), C.byref(n)))
        return a, int(n.value)

    def close(self):
        if self.handle:
            _lib.lib().tpz_unet_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
