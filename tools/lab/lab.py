"""ctypes binding of the bring-up probe library tools/lab/libtpz_lab.so (tools/lab/tpz_lab.cu).  Not part of the product:
nothing under topaz_b200/ imports this module."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, 'libtpz_lab.so')
_I, _P, _LL = C.c_int, C.c_void_p, C.c_longlong
_PROTOS = {
    'tpz_lab_umma': (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    'tpz_lab_umma_pair': (_I, [_P, _P, _I, _I, _P, _P, _P]),
    'tpz_lab_umma_rate': (_I, [_I, _I, _I, _I, _I, _P, _P]),
    'tpz_lab_tma_stride': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'tpz_lab_umma_raw': (_I, [_P, _I, _P, _I, C.c_ulonglong, C.c_ulonglong, C.c_uint, _I, _I, _P, _P]),
}
_lib = None


def lib():
    global _lib
    if _lib is None:
        from topaz_b200 import _lib as product
        product.lib()                      # the probes use the product library's error / tensor-map helpers
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    from topaz_b200 import _lib as product
    product.check(rc)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def lab_umma(A: torch.Tensor, B: torch.Tensor, shift: int, sbo_rows: int, base_off_mode: int, kc: int = 64) -> torch.Tensor:
    D = torch.zeros((128, B.shape[0]), dtype=torch.float32, device=A.device)
    check(lib().tpz_lab_umma(_ptr(A), A.shape[0], _ptr(B), B.shape[0], shift, sbo_rows, base_off_mode, kc, _ptr(D), _stream()))
    return D


def lab_tma_stride(A: torch.Tensor, start: int, stride: int, nrows: int) -> torch.Tensor:
    """Raw shared-memory image (fp16 [nrows, 64], still 128B-swizzled) of a strided TMA box load."""
    out = torch.zeros((nrows, 64), dtype=torch.float16, device=A.device)
    check(lib().tpz_lab_tma_stride(_ptr(A), A.shape[0], start, stride, nrows, _ptr(out), _stream()))
    return out


def lab_umma_rate(N: int, shift: int, sbo_rows: int, iters: int = 2000, two_acc: bool = False) -> float:
    """SM cycles per (M=128, N, K=16) fp16 MMA for an A operand starting at row `shift` with 8-row groups `sbo_rows` rows apart."""
    cyc = torch.zeros(1, dtype=torch.int64, device='cuda')
    check(lib().tpz_lab_umma_rate(N, shift, sbo_rows, iters, int(two_acc), _ptr(cyc), _stream()))
    torch.cuda.synchronize()
    return float(cyc.item()) / (iters * 4 * (2 if two_acc else 1))


def lab_umma_raw(imgA: torch.Tensor, imgB: torch.Tensor, descA: int, descB: int, idesc: int, N: int, kind: int = 1) -> torch.Tensor:
    """One raw tcgen05.mma on caller-built shared-memory images (uint8/float32 tensors on the device) -> D [128, N] fp32."""
    D = torch.zeros((128, N), dtype=torch.float32, device=imgA.device)
    check(lib().tpz_lab_umma_raw(_ptr(imgA), imgA.numel() * imgA.element_size(), _ptr(imgB), imgB.numel() * imgB.element_size(),
                                 descA, descB, idesc, N, kind, _ptr(D), _stream()))
    torch.cuda.synchronize()
    return D
