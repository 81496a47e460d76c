#!/usr/bin/env python
"""CTA-pair (tcgen05 cta_group::2) probe: D[256][N] = A[256][64] @ B[N][64]^T with the operands split over two CTAs."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lab'))
import lab
from lab import check

L = lab.lib()
torch.manual_seed(0)
for N in (64, 128, 256, 32):
    for use_tma in (0, 1):
        A = (torch.randn(256, 64) / 4).half().cuda()
        B = (torch.randn(N, 64) / 4).half().cuda()
        D = torch.full((256, N), -7.0, device='cuda')
        st = torch.zeros(1, dtype=torch.int32, device='cuda')
        check(L.tpz_lab_umma_pair(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), N, use_tma, C.c_void_p(D.data_ptr()),
                                  C.c_void_p(st.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        err = (D - ref).abs().max().item()
        e0 = (D[:128] - ref[:128]).abs().max().item(); e1 = (D[128:] - ref[128:]).abs().max().item()
        print(f'N={N} use_tma={use_tma} status={int(st.item())} max_err={err:.3e} (cta0 rows {e0:.2e}, cta1 rows {e1:.2e})')
        if err > 1e-2:
            # diagnose: which B rows did each output column use?
            sw = torch.cat([ref[:, N // 2:], ref[:, :N // 2]], dim=1)
            print('   vs column-halves swapped:', (D - sw).abs().max().item())
