"""-m gpu parity of the training path: fp32 fwd/dgrad/wgrad kernels vs torch CPU, the fused GE-binomial loss
vs the oracle (autograd), and three full GE_binomial.step calls vs the reference golden (loss tuple, the
gradient of step 1, the parameters after step 3).  fp32 kernels: tolerance 1e-3 per the north star, observed
~1e-5."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from common import gold, weights_of, rel_err, assert_params_after_adam
from oracle import topaz_oracle as O

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('N,H,Ci,Co,k,stride,dil,org', [
    (5, 71, 1, 32, 7, 2, 1, 0), (4, 33, 32, 32, 3, 1, 1, 0), (3, 31, 32, 64, 3, 2, 2, 0), (3, 27, 32, 64, 1, 2, 1, 3),
    (2, 9, 64, 128, 5, 1, 1, 0), (7, 1, 128, 1, 1, 1, 1, 0),
    # Cin = 1 first layer (row-strip wgrad kernel): 64 channels, k=5 stride 1, and an input one pixel larger than the taps reach
    (3, 71, 1, 64, 7, 2, 1, 0), (2, 40, 1, 32, 5, 1, 1, 0), (2, 72, 1, 32, 7, 2, 1, 0),
])
def test_conv_f32_kernels(N, H, Ci, Co, k, stride, dil, org):
    from topaz_b200 import train_engine as T
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Ci, H, H, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) * 0.1
    b = torch.randn(Co, generator=g) * 0.1
    xin = x[:, :, org:, org:]
    ref = F.conv2d(xin, w, b, stride=stride, dilation=dil)
    Ho = ref.shape[2]
    xin = xin[:, :, :(Ho - 1) * stride + (k - 1) * dil + 1, :(Ho - 1) * stride + (k - 1) * dil + 1]
    xd, wd, bd = _nhwc(x).cuda(), w.cuda(), b.cuda()
    y = T._conv_fwd(xd, wd, bd, stride, dil, org, Ho, Ho, relu=False).cpu()
    assert max(rel_err(y, _nhwc(ref))) < 1e-4
    dy = torch.randn(N, Co, Ho, Ho, generator=g)
    gi = torch.nn.grad.conv2d_input(tuple(xin.shape), w, dy, stride=stride, dilation=dil)
    gref = torch.zeros_like(x); gref[:, :, org:org + gi.shape[2], org:org + gi.shape[3]] = gi
    dx = T._conv_dgrad(_nhwc(dy).cuda(), wd, stride, dil, org, H, H).cpu()
    assert max(rel_err(dx, _nhwc(gref))) < 1e-4
    gw = torch.nn.grad.conv2d_weight(xin.contiguous(), tuple(w.shape), dy, stride=stride, dilation=dil)
    dw = torch.zeros_like(w).cuda(); db = torch.zeros_like(b).cuda()
    T._conv_wgrad(xd, _nhwc(dy).cuda(), dw, db, stride, dil, org)
    assert max(rel_err(dw.cpu(), gw)) < 1e-4 and max(rel_err(db.cpu(), dy.sum((0, 2, 3)))) < 1e-4


def test_ge_loss_kernel_matches_oracle_autograd():
    from topaz_b200 import train_engine as T
    g = torch.Generator().manual_seed(3)
    for B, npos, pi in [(256, 16, 0.035), (64, 4, 0.035), (40, 1, 0.2)]:
        s = (2.0 * torch.randn(B, generator=g) - 2.0)
        Y = torch.tensor([1.0] * npos + [0.0] * (B - npos), dtype=torch.float64)
        sr = s.clone().requires_grad_(True)
        cls, ge, loss = O.ge_binomial_loss(sr, Y, pi, 1.0)
        loss.backward()
        prec, tpr, fpr = O.ge_binomial_metrics(s, Y)
        ds = torch.empty(B, device='cuda'); out5 = torch.empty(5, device='cuda')
        T.ge_loss_grad(s.cuda(), Y.cuda(), pi, 1.0, 0, B, ds, out5)
        np.testing.assert_allclose(out5.cpu().numpy(), [cls.item(), ge.item(), prec, tpr, fpr], rtol=2e-5, atol=1e-6)
        assert max(rel_err(ds.cpu(), sr.grad.float())) < 1e-4
        lo, hi = B // 4, B // 2
        ds2 = torch.empty(hi - lo, device='cuda')
        T.ge_loss_grad(s.cuda(), Y.cuda(), pi, 1.0, lo, hi, ds2, out5)
        assert torch.equal(ds2.cpu(), ds.cpu()[lo:hi])


def test_three_ge_binomial_steps_match_reference_golden():
    from topaz_b200.methods import GE_binomial
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']), l2=0.0, slack=1.0)
    B = int(g['B']); Y = torch.from_numpy(g['Y']).cuda()
    # unfilled forward of crops vs the reference golden
    gc = gold('resnet8_u32_pretrained')
    m.eval()
    with torch.no_grad():
        yc = m(torch.from_numpy(gc['crops']).cuda()).cpu().numpy()
    assert yc.shape == gc['y_crops'].shape and max(rel_err(yc, gc['y_crops'])) < 1e-4
    m.train()
    outs = []
    for step in range(3):
        X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
        if step == 0:     # gradient of step 1
            fp = T.flat_params(m)
            score = m(X).view(-1)
            ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
            T.ge_loss_grad(score.contiguous(), Y, tr.pi, tr.slack, 0, B, ds, o5)
            T.backward(m, ds)
            for k, p in m.named_parameters():
                assert max(rel_err(p.grad.cpu().numpy(), g['g1.' + k])) < 1e-3, k
            fp.flat_g.zero_()
        outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=1e-3, atol=1e-6)
    for k, p in m.named_parameters():
        assert max(rel_err(p.detach().cpu().numpy(), g['p3.' + k])) < 1e-3, k
    # after training, the dense (filled) evaluation forward uses the UPDATED weights (plan cache invalidation)
    m.eval(); m.fill()
    x = gc['x']
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.classifier_forward({k: p.detach().cpu().numpy() for k, p in m.state_dict().items()}, x, 'resnet8', 32, filled=True).numpy()
    assert max(rel_err(y, ref)) < 1e-3


@pytest.mark.parametrize('N,H,Ci,Co,k,stride,dil,org', [
    (4, 33, 32, 32, 3, 1, 1, 0), (3, 31, 32, 64, 3, 2, 2, 0), (3, 27, 32, 64, 1, 2, 1, 3), (2, 11, 64, 64, 3, 1, 2, 0),
    (2, 9, 64, 128, 5, 1, 1, 0), (300, 5, 64, 64, 3, 1, 1, 0),
    (48, 27, 64, 128, 1, 2, 1, 3), (48, 25, 64, 128, 3, 2, 2, 0), (48, 9, 128, 256, 5, 1, 1, 0), (48, 11, 128, 128, 3, 1, 2, 0),   # resnet8_u64 layers
])
def test_conv_mma_kernels_match_fp32(N, H, Ci, Co, k, stride, dil, org):
    """3xTF32 tensor-core fwd / dgrad / wgrad vs torch CPU fp32 (fp32-level accuracy expected)."""
    import ctypes as C
    from topaz_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Ci, H, H, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) * 0.1
    b = torch.randn(Co, generator=g) * 0.1
    xin = x[:, :, org:, org:]
    ref = F.conv2d(xin, w, b, stride=stride, dilation=dil)
    Ho = ref.shape[2]
    ext = (Ho - 1) * stride + (k - 1) * dil + 1
    xin = xin[:, :, :ext, :ext]
    P = lambda t: C.c_void_p(t.data_ptr())
    xd, bd = _nhwc(x).cuda(), b.cuda()
    wf = w.permute(2, 3, 1, 0).contiguous().cuda()      # [tap][ci][co]
    wg = w.permute(2, 3, 0, 1).contiguous().cuda()      # [tap][co][ci]
    y = torch.empty(N, Ho, Ho, Co, device='cuda')
    _lib.check(L.tpz_conv_fwd_mma(P(xd), N, H, H, Ci, P(wf), P(bd), Co, k, k, stride, dil, org, None, 0, 0, 0, 1, 0, P(y), Ho, Ho, None))
    assert max(rel_err(y.cpu(), _nhwc(ref))) < 5e-5
    dy = torch.randn(N, Co, Ho, Ho, generator=g)
    dyd = _nhwc(dy).cuda()
    gi = torch.nn.grad.conv2d_input(tuple(xin.shape), w, dy, stride=stride, dilation=dil)
    gref = torch.zeros_like(x); gref[:, :, org:org + ext, org:org + ext] = gi
    dx = torch.empty(N, H, H, Ci, device='cuda')
    _lib.check(L.tpz_conv_dgrad_mma(P(dyd), N, Ho, Ho, Co, P(wg), Ci, k, k, stride, dil, org, None, 0, P(dx), H, H, None))
    assert max(rel_err(dx.cpu(), _nhwc(gref))) < 5e-5
    gw = torch.nn.grad.conv2d_weight(xin.contiguous(), tuple(w.shape), dy, stride=stride, dilation=dil)
    dw = torch.zeros_like(w).cuda()
    _lib.check(L.tpz_conv_wgrad_mma(P(xd), N, H, H, Ci, P(dyd), Ho, Ho, Co, k, k, stride, dil, org, P(dw), None))
    assert max(rel_err(dw.cpu(), gw)) < 5e-5


@pytest.mark.parametrize('N,H,Ci,Co,k,stride,dil,org', [
    (4, 33, 32, 32, 3, 1, 1, 0), (3, 31, 32, 64, 3, 2, 2, 0), (3, 27, 32, 64, 1, 2, 1, 3), (2, 11, 64, 64, 3, 1, 2, 0),
    (2, 9, 64, 128, 5, 1, 1, 0), (300, 5, 64, 64, 3, 1, 1, 0),
    (48, 27, 64, 128, 1, 2, 1, 3), (48, 25, 64, 128, 3, 2, 2, 0), (48, 9, 128, 256, 5, 1, 1, 0), (48, 11, 128, 128, 3, 1, 2, 0),   # resnet8_u64 layers
    (256, 31, 32, 32, 3, 1, 1, 0),                                                                                                 # resnet8_u32 at the cfg4 minibatch
    # halo-resident kernel (stride 1, 32 -> 32): dilation 2 (largest halo), a cropped origin, 1x1, fewer tiles than SMs
    (40, 31, 32, 32, 3, 1, 2, 0), (3, 20, 32, 32, 3, 1, 1, 2), (2, 17, 32, 32, 1, 1, 1, 1), (1, 9, 32, 32, 3, 1, 1, 0),
    (256, 27, 32, 32, 3, 1, 1, 0),
    # one-pixel outputs (the last 5x5 layer on a training crop): split-K forward + per-tap scatter dgrad; an uncovered variant
    (256, 5, 64, 128, 5, 1, 1, 0), (40, 6, 32, 64, 3, 1, 2, 1), (70, 5, 128, 256, 5, 1, 1, 0),
    # 64-channel halo-resident kernel (streamed weights): r3.conv0 / r3.conv1 at the cfg4 minibatch, a u64 first-stage map, an origin
    (256, 11, 64, 64, 3, 1, 1, 0), (256, 9, 64, 64, 3, 1, 2, 0), (6, 33, 64, 64, 3, 1, 1, 0), (5, 14, 64, 64, 3, 1, 1, 1),
])
def test_conv_tc_kernels_match_fp32(N, H, Ci, Co, k, stride, dil, org):
    """tcgen05 (kind::tf32, 3-pass) fwd / dgrad / wgrad vs torch CPU fp32, with the weights packed by tpz_train_repack_tc
    (fp32-level accuracy expected, as for the mma.sync kernels above)."""
    import ctypes as C
    from topaz_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Ci, H, H, generator=g)
    w = torch.randn(Co, Ci, k, k, generator=g) * 0.1
    b = torch.randn(Co, generator=g) * 0.1
    xin = x[:, :, org:, org:]
    ref = F.conv2d(xin, w, b, stride=stride, dilation=dil)
    Ho = ref.shape[2]
    ext = (Ho - 1) * stride + (k - 1) * dil + 1
    xin = xin[:, :, :ext, :ext]
    P = lambda t: C.c_void_p(t.data_ptr())
    xd, bd = _nhwc(x).cuda(), b.cuda()
    n = w.numel()
    desc = np.zeros(1, dtype=[('src', '<i8'), ('fwd', '<i8'), ('dg', '<i8'), ('co', '<i4'), ('ci', '<i4'), ('taps', '<i4'), ('pad', '<i4')])
    desc[0] = (0, 0, 2 * n, Co, Ci, k * k, 0)
    dd = torch.from_numpy(desc.view(np.uint8).copy()).cuda()
    packed = torch.zeros(4 * n, device='cuda')
    wd = w.contiguous().cuda()
    _lib.check(L.tpz_train_repack_tc(P(wd), P(dd), 1, n, P(packed), None))
    pk = packed.cpu()
    # the packed planes reproduce the weights: hi + lo == w, laid out [tap][ci/32][co][32] / [tap][co/32][ci][32]
    wf = (pk[:n] + pk[n:2 * n]).reshape(k * k, Ci // 32, Co, 32).permute(2, 1, 3, 0).reshape(Co, Ci, k, k)
    wg = (pk[2 * n:3 * n] + pk[3 * n:]).reshape(k * k, Co // 32, Ci, 32).permute(1, 3, 2, 0).reshape(Co, Ci, k, k)
    assert torch.equal(wf, w) and torch.equal(wg, w)
    y = torch.empty(N, Ho, Ho, Co, device='cuda')
    res = torch.randn(N, Ho, Ho, Co, generator=g)
    _lib.check(L.tpz_conv_fwd_tc(P(xd), N, H, H, Ci, P(packed), P(bd), Co, k, k, stride, dil, org, None, 0, 0, 0, 1, 0, P(y), Ho, Ho, None))
    e = max(rel_err(y.cpu(), _nhwc(ref)))
    print('conv_fwd_tc rel err', e)
    assert e < 5e-5
    resd = res.cuda()
    _lib.check(L.tpz_conv_fwd_tc(P(xd), N, H, H, Ci, P(packed), P(bd), Co, k, k, stride, dil, org, P(resd), Ho, Ho, 0, 1, 1, P(y), Ho, Ho, None))
    assert max(rel_err(y.cpu(), torch.relu(_nhwc(ref) + res))) < 5e-5
    dy = torch.randn(N, Co, Ho, Ho, generator=g)
    dyd = _nhwc(dy).cuda()
    gi = torch.nn.grad.conv2d_input(tuple(xin.shape), w, dy, stride=stride, dilation=dil)
    gref = torch.zeros_like(x); gref[:, :, org:org + ext, org:org + ext] = gi
    dx = torch.empty(N, H, H, Ci, device='cuda')
    _lib.check(L.tpz_conv_dgrad_tc(P(dyd), N, Ho, Ho, Co, P(packed[2 * n:]), Ci, k, k, stride, dil, org, None, 0, P(dx), H, H, None))
    e = max(rel_err(dx.cpu(), _nhwc(gref)))
    print('conv_dgrad_tc rel err', e)
    assert e < 5e-5
    mask = torch.randn(N, H, H, Ci, generator=g)
    base = torch.randn(N, H, H, Ci, generator=g)
    dx2 = base.clone().cuda()
    maskd0 = mask.cuda()
    _lib.check(L.tpz_conv_dgrad_tc(P(dyd), N, Ho, Ho, Co, P(packed[2 * n:]), Ci, k, k, stride, dil, org, P(maskd0), 1, P(dx2), H, H, None))
    assert max(rel_err(dx2.cpu(), torch.where(mask > 0, base + _nhwc(gref), torch.zeros(())))) < 5e-5
    if H > 4:       # fused form: + the gradient of a cropped identity skip (embedded at res_org) before the mask
        rs = torch.randn(N, H - 2, H - 2, Ci, generator=g)
        emb = torch.zeros(N, H, H, Ci); emb[:, 1:H - 1, 1:H - 1] = rs
        dx3 = torch.empty(N, H, H, Ci, device='cuda')
        maskd, rsd = mask.cuda(), rs.cuda()                # named: a temporary's block could be reused before the launch reads it
        _lib.check(L.tpz_conv_dgrad_tc_res(P(dyd), N, Ho, Ho, Co, P(packed[2 * n:]), Ci, k, k, stride, dil, org, P(maskd), 0,
                                           P(rsd), H - 2, H - 2, 1, P(dx3), H, H, None))
        assert max(rel_err(dx3.cpu(), torch.where(mask > 0, _nhwc(gref) + emb, torch.zeros(())))) < 5e-5
    gw = torch.nn.grad.conv2d_weight(xin.contiguous(), tuple(w.shape), dy, stride=stride, dilation=dil)
    dw = torch.zeros_like(w).cuda()
    _lib.check(L.tpz_conv_wgrad_tc(P(xd), N, H, H, Ci, P(dyd), Ho, Ho, Co, k, k, stride, dil, org, P(dw), None))
    e = max(rel_err(dw.cpu(), gw))
    print('conv_wgrad_tc rel err', e)
    assert e < 5e-5
    dw2 = torch.zeros_like(w).cuda(); db = torch.full((Co,), 0.5, device='cuda')
    _lib.check(L.tpz_conv_wgrad_tc_bias(P(xd), N, H, H, Ci, P(dyd), Ho, Ho, Co, k, k, stride, dil, org, P(dw2), P(db), None))
    assert max(rel_err(dw2.cpu(), gw)) < 5e-5
    assert max(rel_err(db.cpu(), 0.5 + dy.sum((0, 2, 3)))) < 2e-5


@pytest.mark.parametrize('N,H,stride', [(256, 71, 2), (3, 71, 2), (5, 40, 1), (1, 9, 2)])
def test_first_layer_tc_kernels_match_fp32(N, H, stride):
    """Cin = 1 7x7 first layer on the tensor core (tpz_first_fwd_tc / tpz_first_wgrad_tc: im2col tile built in shared memory,
    3xTF32) vs torch fp32."""
    import ctypes as C
    from topaz_b200 import _lib
    L = _lib.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 1, H, H, generator=g); w = torch.randn(32, 1, 7, 7, generator=g) * 0.2; b = torch.randn(32, generator=g) * 0.1
    ref = F.conv2d(x, w, b, stride=stride)
    Ho = ref.shape[2]
    xd, wd, bd = x.reshape(N, H, H).contiguous().cuda(), w.cuda(), b.cuda()
    y = torch.empty(N, Ho, Ho, 32, device='cuda')
    assert L.tpz_first_fwd_tc(P(xd), N, H, H, P(wd), P(bd), 32, 7, stride, 1, P(y), Ho, Ho, None) == 0
    e = max(rel_err(y.cpu(), _nhwc(torch.relu(ref))))
    print('first_fwd_tc rel err', e)
    assert e < 5e-5
    dy = torch.randn(N, 32, Ho, Ho, generator=g)
    dyd = _nhwc(dy).cuda()
    ext = (Ho - 1) * stride + 7
    gw = torch.nn.grad.conv2d_weight(x[:, :, :ext, :ext].contiguous(), tuple(w.shape), dy, stride=stride)
    dw = torch.full((32, 1, 7, 7), 0.25, device='cuda'); db = torch.full((32,), 0.5, device='cuda')
    assert L.tpz_first_wgrad_tc(P(xd), N, H, H, P(dyd), Ho, Ho, 32, 7, stride, P(dw), P(db), None) == 0
    assert max(rel_err(db.cpu(), 0.5 + dy.sum((0, 2, 3)))) < 2e-5
    e = max(rel_err(dw.cpu() - 0.25, gw))
    print('first_wgrad_tc rel err', e)
    assert e < 5e-5
    # shapes the tensor-core kernels do not cover are declined, not mis-computed
    assert L.tpz_first_fwd_tc(P(xd), N, H, H, P(wd), P(bd), 64, 7, stride, 1, P(y), Ho, Ho, None) == -1


@pytest.mark.parametrize('M,C,masked', [(256, 128, True), (37, 256, False), (1000, 64, True)])
def test_classifier_head_kernels_match_fp32(M, C, masked):
    """tpz_cls_fwd_f32 / tpz_cls_bwd_f32 (1x1 conv C -> 1 and its fused backward) vs torch fp32."""
    import ctypes as C_
    from topaz_b200 import _lib
    L = _lib.lib()
    P = lambda t: C_.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(M, C, generator=g); w = torch.randn(C, generator=g) * 0.1; b = torch.randn(1, generator=g)
    gy = torch.randn(M, generator=g)
    xd, wd, bd, gd = x.cuda(), w.cuda(), b.cuda(), gy.cuda()
    y = torch.empty(M, device='cuda')
    _lib.check(L.tpz_cls_fwd_f32(P(xd), M, C, P(wd), P(bd), P(y), None))
    assert max(rel_err(y.cpu(), x @ w + b)) < 1e-5
    dx = torch.empty(M, C, device='cuda'); dw = torch.full((C,), 0.5, device='cuda'); db = torch.full((1,), 0.25, device='cuda')
    _lib.check(L.tpz_cls_bwd_f32(P(xd), M, C, P(wd), P(gd), int(masked), P(dx), P(dw), P(db), None))
    dx_ref = gy[:, None] * w[None, :]
    if masked:
        dx_ref = torch.where(x > 0, dx_ref, torch.zeros(()))
    assert torch.equal(dx.cpu(), dx_ref)
    assert max(rel_err(dw.cpu(), 0.5 + gy @ x)) < 1e-5
    assert abs(float(db.cpu()) - 0.25 - float(gy.sum())) < 1e-4


@pytest.mark.parametrize('tag', ['PN', 'PNpi', 'GE_KL', 'PU', 'PUclip'])
def test_other_objectives_match_reference_golden(tag):
    """PN / GE_KL / PU on the GPU: loss tuples of 2 steps and updated parameters vs the reference goldens; the fused loss
    kernel's d/dscore vs the oracle's autograd."""
    from topaz_b200 import methods as M, train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold('objectives_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    opt = torch.optim.Adam(m.parameters(), lr=2e-4); crit = nn.BCEWithLogitsLoss()
    cfg = {'PN': (0, -1.0, 1.0, 1.0, 0.0), 'PNpi': (0, 0.1, 1.0, 1.0, 0.0), 'GE_KL': (1, 0.035, 1.0, 0.9, 0.035),
           'PU': (2, 0.035, 1.0, 1.0, 0.0), 'PUclip': (2, 0.6, 1.0, 1.0, 0.0)}[tag]
    # kernel vs oracle autograd on random logits
    gen = torch.Generator().manual_seed(5)
    s = 2.0 * torch.randn(64, generator=gen) - 1.0
    Yt = torch.tensor([1.0] * 6 + [0.0] * 58, dtype=torch.float64)
    sr = s.clone().requires_grad_(True)
    name = ['PN', 'GE_KL', 'PU'][cfg[0]]
    rep, ge, back, _ = O.pu_objective_loss(name, sr, Yt, (cfg[1] if cfg[1] > 0 else None) if cfg[0] == 0 else cfg[1], cfg[2], cfg[3], cfg[4], cfg[4])
    back.backward()
    ds = torch.empty(64, device='cuda'); o6 = torch.empty(6, device='cuda')
    T.pu_objective_loss_grad(s.cuda(), Yt.cuda(), *cfg, 0, 64, ds, o6)
    assert max(rel_err(ds.cpu(), sr.grad.float())) < 1e-4
    assert abs(o6[0].item() - rep.item()) < 1e-5 * max(1, abs(rep.item()))
    tr = {'PN': lambda: M.PN(m, opt, crit, pi=None), 'PNpi': lambda: M.PN(m, opt, crit, pi=0.1),
          'GE_KL': lambda: M.GE_KL(m, opt, crit, 0.035, slack=1.0, momentum=0.9),
          'PU': lambda: M.PU(m, opt, crit, 0.035, beta=0.0), 'PUclip': lambda: M.PU(m, opt, crit, 0.6, beta=0.0)}[tag]()
    B = int(g['B']); Y = torch.from_numpy(g['Y']).cuda()
    outs = []
    for step in range(2):
        X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
        outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g[tag + '.outs'], rtol=1e-3, atol=1e-6)
    sdn = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for k in ['classifier.weight', 'features.features.0.conv.weight', 'features.features.2.proj.weight']:
        assert max(rel_err(sdn[k], g[tag + '.p.' + k])) < 1e-3, k


def test_single_pass_tf32_mode_gradients():
    """Optional single-pass TF32 training mode (precision of the reference's own cuDNN path): gradients of one step
    stay within 5e-2 (max-norm; measured 2e-2 on B200 - TF32 operand rounding through the backward chain, the same class
    of error as the reference's cuDNN/TF32 GPU path) of the fp32 reference golden.  Opt-in only; the default 3xTF32 mode is
    the one held to 1e-3.  The mode switch is restored afterwards."""
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    B = int(g['B']); Y = torch.from_numpy(g['Y']).cuda()
    X = torch.from_numpy(np.random.default_rng(4000).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
    prev = T.set_tf32(True)
    try:
        T.flat_params(m)
        score = m(X).view(-1)
        ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
        T.ge_loss_grad(score.contiguous(), Y, float(g['pi']), 1.0, 0, B, ds, o5)
        T.backward(m, ds)
        worst = 0.0
        for k, p in m.named_parameters():
            worst = max(worst, max(rel_err(p.grad.cpu().numpy(), g['g1.' + k])))
        print('single-pass TF32 worst gradient rel err', worst)
        assert worst < 5e-2
    finally:
        assert T.set_tf32(prev) is True


def test_u64_training_step_gradients_match_oracle_autograd():
    """cfg4 also names resnet8_u64 (64/128/256 channels -> the BN=64 tensor-core training kernels): forward logits and the
    gradient of one GE-binomial step vs autograd through the oracle, pretrained u64 weights, 48 crops."""
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    sd = weights_of(gold('resnet8_u64_pretrained'))
    m = LinearClassifier(get_feature_extractor('resnet8', units=64, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    B, pi = 48, 0.05
    X = torch.from_numpy(np.random.default_rng(77).standard_normal((B, 71, 71)).astype(np.float32))
    Y = torch.tensor([1.0] * 5 + [0.0] * (B - 5), dtype=torch.float64)
    params = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()}
    T.flat_params(m)
    score = m(X.cuda()).view(-1)
    # the oracle runs with the GPU forward's ReLU masks imposed: among the 12 M activations of this step a few lie within
    # rounding noise of zero, and one flipped mask moves upstream gradients by ~3e-3 (see the BatchNorm test below)
    masks = []
    for rec in m.__dict__['_tpz_tape']:
        if rec['kind'] == 'conv':
            masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
        elif rec['kind'] == 'resid':
            masks += [(rec['h'] > 0).permute(0, 3, 1, 2).cpu(), (rec['y'] > 0).permute(0, 3, 1, 2).cpu()]
    score_ref = O.classifier_forward_grad(params, X[:, None], 'resnet8', 64, relu_masks=masks).view(-1)
    _, _, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    T.backward(m, ds)
    errs = {k: max(rel_err(p.grad.cpu().numpy(), params[k].grad.numpy())) for k, p in m.named_parameters()}
    print({k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs


def test_gradients_match_plain_oracle_autograd_without_imposed_masks():
    """The gradient tests above impose the GPU forward's ReLU masks on the oracle, so their backward is never checked against
    an INDEPENDENT forward.  Here nothing is imposed: the oracle's own ReLU decides every mask.  That is only meaningful on a
    minibatch whose pre-activations keep a margin from zero (one flipped mask moves early-layer gradients by ~3e-3): the test
    uses crop seeds whose smallest |pre-activation| on the ORACLE is the largest of a 400-seed scan (asserted here), and keeps
    the first one on which the GPU forward takes the same 247 168 ReLU decisions as the oracle.  (A margin of 1e-4 is out of
    reach: two resnet8_u32 crops have 2.5e5 pre-activations with density ~0.4 per unit around zero, i.e. ~20 of them inside
    +-1e-4 for any seed; the scan's best is 2.2e-5.)"""
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    sd = weights_of(gold('resnet8_u32_pretrained'))
    B, pi = 2, 0.3
    Y = torch.tensor([1.0, 0.0], dtype=torch.float64)
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    T.flat_params(m)
    chosen = None
    # crop seeds with the largest margins among 9000..9399 (found by scanning the oracle on the CPU; re-asserted here)
    for seed in (9063, 9182, 9030, 9022):
        X = torch.from_numpy(np.random.default_rng(seed).standard_normal((B, 71, 71)).astype(np.float32))
        O.PREACT_LOG = []
        try:
            with torch.no_grad():
                O.classifier_forward(sd, X[:, None], 'resnet8', 32, filled=False)
            pre = O.PREACT_LOG
        finally:
            O.PREACT_LOG = None
        margin = min(float(t.abs().min()) for t in pre)
        assert margin >= 1.5e-5, (seed, margin)
        # the minibatch must be one on which both forwards agree on EVERY ReLU decision (nothing is imposed on either side)
        score = m(X.cuda()).view(-1)
        gpu_masks = []
        for rec in m.__dict__['_tpz_tape']:
            if rec['kind'] == 'conv':
                gpu_masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
            elif rec['kind'] == 'resid':
                gpu_masks += [(rec['h'] > 0).permute(0, 3, 1, 2).cpu(), (rec['y'] > 0).permute(0, 3, 1, 2).cpu()]
        flips = sum(int((a != (b > 0)).sum()) for a, b in zip(gpu_masks, pre))
        print(f'un-imposed gradient test: seed {seed}, smallest |pre-activation| on the oracle {margin:.2e}, '
              f'{sum(t.numel() for t in pre)} ReLU decisions, {flips} differ on the GPU')
        if flips == 0:
            chosen = (seed, X, score)
            break
    assert chosen is not None, 'no candidate minibatch with identical ReLU decisions'
    seed, X, score = chosen
    params = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()}
    score_ref = O.classifier_forward_grad(params, X[:, None], 'resnet8', 32).view(-1)          # plain autograd, no masks passed
    _, _, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    T.backward(m, ds)
    errs = {k: max(rel_err(p.grad.cpu().numpy(), params[k].grad.numpy())) for k, p in m.named_parameters()}
    print({k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs


@pytest.mark.parametrize('N,Ci,Co', [(48, 64, 128), (256, 32, 64), (5, 64, 128)])
def test_proj_wgrad_dgrad_engine_geometry(N, Ci, Co):
    """The 1x1 stride-2 projection of ResidA2 as the engine calls it: input 27^2, crop offset 3, output 11^2 (the main
    path's size, one less than the full extent would give); non-negative (post-ReLU) activations like the real ones."""
    import ctypes as C
    from topaz_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    H, Ho, org, stride = 27, 11, 3, 2
    x = torch.relu(torch.randn(N, Ci, H, H, generator=g)) * 3.0
    w = torch.randn(Co, Ci, 1, 1, generator=g) * 0.1
    dy = torch.randn(N, Co, Ho, Ho, generator=g) * 0.01
    xin = x[:, :, org:org + 2 * Ho - 1, org:org + 2 * Ho - 1].contiguous()
    gw = torch.nn.grad.conv2d_weight(xin.double(), tuple(w.shape), dy.double(), stride=stride).float()
    gi = torch.nn.grad.conv2d_input(tuple(xin.shape), w.double(), dy.double(), stride=stride).float()
    P = lambda t: C.c_void_p(t.data_ptr())
    xd, dyd = _nhwc(x).cuda(), _nhwc(dy).cuda()
    dw = torch.zeros_like(w).cuda()
    _lib.check(L.tpz_conv_wgrad_mma(P(xd), N, H, H, Ci, P(dyd), Ho, Ho, Co, 1, 1, stride, 1, org, P(dw), None))
    e = rel_err(dw.cpu(), gw)
    print('proj wgrad rel err', e)
    assert max(e) < 5e-5
    wg = w.permute(2, 3, 0, 1).contiguous().cuda()
    dx = torch.zeros(N, H, H, Ci, device='cuda')
    _lib.check(L.tpz_conv_dgrad_mma(P(dyd), N, Ho, Ho, Co, P(wg), Ci, 1, 1, stride, 1, org, None, 1, P(dx), H, H, None))
    gref = torch.zeros_like(x); gref[:, :, org:org + 2 * Ho - 1, org:org + 2 * Ho - 1] = gi
    e = rel_err(dx.cpu(), _nhwc(gref))
    print('proj dgrad rel err', e)
    assert max(e) < 5e-5


# ------------------------------------------------------------------------------------------------
# training-mode BatchNorm (the default `topaz train` model has --bn on: reference commands/train.py:91)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,H,C', [(64, 33, 32), (16, 11, 64), (256, 1, 128), (3, 5, 6), (2, 7, 48)])
def test_batchnorm_kernels_match_torch_fp64(N, H, C):
    """tpz_bn_stats / tpz_bn_fwd / tpz_bn_bwd_reduce / tpz_bn_bwd vs torch.nn.functional.batch_norm + autograd in fp64
    (training mode: batch statistics, running-buffer update; eval mode: running statistics).  C = 6 takes the scalar
    (C % 4 != 0) path, the others the float4 path; P from 25 to 69 696 rows covers all three reduction grids."""
    from topaz_b200 import train_engine as T
    rng = np.random.default_rng(N * 1000 + C)
    x = torch.from_numpy((1.5 + 2.0 * rng.standard_normal((N, H, H, C))).astype(np.float32)).cuda()       # NHWC, mean != 0
    g = torch.from_numpy(rng.standard_normal((N, H, H, C)).astype(np.float32)).cuda()
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)).cuda()
    beta = torch.from_numpy((0.3 * rng.standard_normal(C)).astype(np.float32)).cuda()
    rm0 = torch.from_numpy((0.1 * rng.standard_normal(C)).astype(np.float32)).cuda()
    rv0 = torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32)).cuda()
    eps, mom = 1e-5, 0.1
    P = N * H * H
    # kernels: statistics + forward
    sums = torch.zeros(2 * C, dtype=torch.float64, device='cuda')
    T._bn_stats(x, sums)
    xd = x.double().reshape(-1, C)
    assert max(rel_err(sums[:C].cpu(), xd.sum(0).cpu())) < 1e-12 and max(rel_err(sums[C:].cpu(), (xd * xd).sum(0).cpu())) < 1e-12
    save = torch.empty(2 * C, device='cuda')
    rm1, rv1 = rm0.clone(), rv0.clone()
    y = T._bn_fwd(x, sums, P, gamma, beta, eps, mom, rm1, rv1, True, save)
    # fp64 reference (NCHW); the ReLU mask of the kernel output is imposed so that outputs within fp32 rounding of zero
    # cannot flip a mask between the two (their forward values are ~1e-7 either way)
    xr = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    yr = F.batch_norm(xr, rm, rv, gr, br, True, mom, eps) * (y > 0).permute(0, 3, 1, 2).double()
    yr.backward(g.double().permute(0, 3, 1, 2))
    assert max(rel_err(y.permute(0, 3, 1, 2).cpu(), yr.detach().cpu())) < 2e-6
    assert max(rel_err(rm1.cpu(), rm.cpu())) < 1e-6 and max(rel_err(rv1.cpu(), rv.cpu())) < 1e-6
    mean = xd.mean(0); invstd = 1.0 / torch.sqrt(xd.var(0, unbiased=False) + eps)
    assert max(rel_err(save[:C].cpu(), mean.cpu())) < 1e-6 and max(rel_err(save[C:].cpu(), invstd.cpu())) < 1e-6
    # backward: g masked by the ReLU of y, as _conv_dgrad(mask=...) / _relu_bwd deliver it
    gm = (g * (y > 0)).contiguous()
    local = torch.zeros(2 * C, dtype=torch.float64, device='cuda')
    T._bn_bwd_reduce(gm, x, save, local)
    dgamma = torch.zeros(C, device='cuda'); dbeta = torch.zeros(C, device='cuda')
    T._bn_bwd(gm, x, save, local, P, gamma, local, dgamma, dbeta)
    assert max(rel_err(dgamma.cpu(), gr.grad.cpu())) < 1e-5, rel_err(dgamma.cpu(), gr.grad.cpu())
    assert max(rel_err(dbeta.cpu(), br.grad.cpu())) < 1e-5
    assert max(rel_err(gm.permute(0, 3, 1, 2).cpu(), xr.grad.cpu())) < 1e-5, rel_err(gm.permute(0, 3, 1, 2).cpu(), xr.grad.cpu())
    # eval mode: statistics read from `save` (running buffers), no ReLU
    save_e = torch.cat([rm0, torch.rsqrt(rv0 + eps)])
    ye = T._bn_fwd(x, None, 0, gamma, beta, eps, 0.0, None, None, False, save_e)
    yer = F.batch_norm(x.double().permute(0, 3, 1, 2), rm0.double(), rv0.double(), gamma.double(), beta.double(), False, mom, eps)
    assert max(rel_err(ye.permute(0, 3, 1, 2).cpu(), yer.cpu())) < 2e-6


def _bn_case():
    from common import seeded_state
    from common_shapes import classifier_shapes
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold('ge_binomial_u32_bn')
    sd = seeded_state(classifier_shapes('resnet8', 32, 1, True), int(g['seed']))
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=True))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return g, sd, m


def test_batchnorm_training_gradients_match_oracle_with_imposed_masks():
    """One GE-binomial step of ResNet8(units=32, bn=True) in train() mode: logits, loss tuple and EVERY gradient vs autograd
    through the oracle.  He-random weights leave a handful of the 8 M BatchNorm outputs within rounding noise of zero, and a
    single flipped ReLU mask moves the early-layer gradients by ~3e-3 (tests/test_host_logic.py), so the oracle is run with
    the ReLU masks of the GPU forward imposed (O._act); the flipped activations themselves are ~1e-7, invisible in the
    forward.  With the masks aligned the north-star tolerance (1e-3) applies to every tensor."""
    from topaz_b200 import train_engine as T
    g, sd, m = _bn_case()
    m.cuda(); m.train()
    B, pi = int(g['B']), float(g['pi'])
    X = torch.from_numpy(np.random.default_rng(4000).standard_normal((B, 71, 71)).astype(np.float32))
    Y = torch.from_numpy(g['Y'])
    T.flat_params(m)
    score = m(X.cuda()).view(-1)
    tape = m.__dict__['_tpz_tape']
    masks = []
    for rec in tape:
        if rec['kind'] == 'conv':
            masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
        elif rec['kind'] == 'resid':
            masks.append((rec['h'] > 0).permute(0, 3, 1, 2).cpu())
            masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
    assert len(masks) == 8
    params = {k: torch.from_numpy(v).clone().requires_grad_('running' not in k and v.dtype == np.float32) for k, v in sd.items()}
    running = {}
    score_ref = O.classifier_forward_grad(params, X, 'resnet8', 32, bn=True, running=running, relu_masks=masks).view(-1)
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    cls, ge, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    np.testing.assert_allclose(o5.cpu().numpy()[:2], [cls.item(), ge.item()], rtol=1e-4)
    T.backward(m, ds)
    errs = {k: max(rel_err(p.grad.cpu().numpy(), params[k].grad.numpy())) for k, p in m.named_parameters()}
    print({k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs
    # running statistics after this one forward (momentum 0.1, unbiased variance) and the batch counter
    sdm = m.state_dict()
    for k, v in running.items():
        assert max(rel_err(sdm[k].cpu().numpy(), v.numpy())) < 1e-5, k
    assert all(int(v) == 1 for k, v in sdm.items() if k.endswith('num_batches_tracked'))


def test_three_ge_binomial_batchnorm_steps_match_reference_golden():
    """Three full GE_binomial.step calls on the BatchNorm model vs the REAL reference's golden: loss tuples, BatchNorm
    running buffers (forward-only quantities: tight), parameters (Adam moves an element by ~lr per step whatever |g| is,
    so elements whose tiny gradient changed sign through a mask flip end up to 2*lr*3 apart: max loose, rel-L2 tight);
    then the eval-mode strided forward (running statistics) and the filled dense forward with the folded BatchNorm."""
    from topaz_b200.methods import GE_binomial
    g, sd, m = _bn_case()
    m.cuda(); m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']), l2=0.0, slack=1.0)
    B = int(g['B']); Y = torch.from_numpy(g['Y']).cuda()
    outs = []
    for step in range(3):
        X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
        outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=2e-3, atol=1e-6)
    for k, v in m.state_dict().items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == 3
        elif 'running' in k:
            assert max(rel_err(v.cpu().numpy(), g['p3.' + k])) < 1e-3, k
        else:
            assert_params_after_adam(v.detach().cpu().numpy(), g['p3.' + k], 3, 1e-3, k)
    m.eval()
    with torch.no_grad():
        yc = m(torch.from_numpy(np.random.default_rng(4100).standard_normal((8, 71, 71)).astype(np.float32)).cuda()).cpu().numpy()
    assert yc.shape == g['y_crops'].shape and max(rel_err(yc, g['y_crops'])) < 2e-3
    m.fill()
    with torch.no_grad():
        yd = m(torch.from_numpy(g['x_dense']).cuda()).cpu().numpy()
    assert yd.shape == g['y_dense'].shape and max(rel_err(yd, g['y_dense'])) < 5e-3


def test_epoch_boundary_round_trip_keeps_the_optimizer_state(tmp_path):
    """What the reference's fit_epochs does between epochs (training.py:576-603): eval() + fill() + a dense forward +
    unfill(), then classifier.cpu(), torch.save(classifier), classifier.cuda(), then train() again.  The parameters change
    storage twice; the flat training buffers are rebuilt around them and must keep the Adam moments and the step count, so
    the interrupted run reproduces the uninterrupted one (up to the fp32 atomics' summation order)."""
    from topaz_b200.methods import GE_binomial

    def run(interrupt):
        g, sd, m = _bn_case()
        m.cuda(); m.train()
        optim = torch.optim.Adam(m.parameters(), lr=2e-4)
        tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']), l2=1e-5, slack=1.0)
        B = 32; Y = torch.from_numpy(g['Y'][:B]).cuda()
        outs = []
        for step in range(4):
            X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
            outs.append(tr.step(X, Y))
            if interrupt and step == 1:
                m.eval(); m.fill()
                with torch.no_grad():
                    yd = m(torch.from_numpy(g['x_dense']).cuda())
                assert torch.isfinite(yd).all()
                m.unfill()
                m.cpu()
                path = str(tmp_path / 'model_epoch1.sav')
                torch.save(m, path)
                assert os.path.getsize(path) < 2 * 1024 * 1024          # parameters + buffers only (1.3 MB), no engine caches
                m.cuda(); m.train()
        st = optim.state[next(iter(m.parameters()))]
        return np.array(outs), {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}, float(st['step'])
    o0, s0, n0 = run(False)
    o1, s1, n1 = run(True)
    assert n0 == n1 == 4.0
    np.testing.assert_allclose(o1, o0, rtol=1e-3, atol=1e-6)
    for k in s0:
        if 'num_batches' in k:
            assert int(s0[k]) == int(s1[k]) == 4
        else:
            assert_params_after_adam(s1[k], s0[k], 4, 1e-3, k)     # a lost Adam state shifts every element by ~lr: rel-L2 ~3e-3


# ------------------------------------------------------------------------------------------------
# PReLU / LeakyReLU extractors in training (`topaz train -m conv31|conv63|conv127`, reference basic.py:16-78)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,learnable', [(1000, True), (64 * 13 * 13 * 16, True), (5 * 7 * 7 * 33, False)])
def test_activation_kernels_match_torch(n, learnable):
    from topaz_b200 import train_engine as T
    rng = np.random.default_rng(n)
    v = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).cuda()
    v[::17] = 0.0                                                      # exact zeros take the "v <= 0" branch, as in torch
    g = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).cuda()
    act = (nn.PReLU(init=0.2) if learnable else nn.LeakyReLU(0.1)).cuda()
    if learnable:
        act.weight.grad = torch.zeros_like(act.weight)
    vr = v.double().clone().requires_grad_(True)
    if learnable:
        ar = act.weight.detach().double().clone().requires_grad_(True)
        yr = F.prelu(vr, ar)
    else:
        yr = F.leaky_relu(vr, 0.1)
    yr.backward(g.double())
    y = T._act_fwd(v, act)
    assert torch.equal(y.cpu(), torch.where(v > 0, v, (act.weight.detach() if learnable else torch.tensor(0.1, device='cuda')) * v).cpu())
    gi = g.clone()
    T._act_bwd(gi, v, act)
    assert max(rel_err(gi.cpu(), vr.grad.cpu())) < 1e-6
    if learnable:
        assert max(rel_err(act.weight.grad.cpu(), ar.grad.cpu())) < 1e-5


@pytest.mark.parametrize('tag,bn', [('ge_binomial_conv31_bn', True), ('ge_binomial_conv31_nobn', False)])
def test_prelu_extractor_training_matches_oracle_and_reference_golden(tag, bn):
    """conv31 (16 units x2; conv -> [BN] -> PReLU with a learnable slope): every gradient of one GE-binomial step vs autograd
    through the oracle with the GPU forward's activation branches imposed (same reasoning as the ReLU masks of the BatchNorm
    ResNet test), then two full steps vs the reference's golden."""
    from common import seeded_state
    from common_shapes import classifier_shapes
    from topaz_b200 import train_engine as T
    from topaz_b200.methods import GE_binomial
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold(tag)
    sd = seeded_state(classifier_shapes('conv31', 16, 2, bn), int(g['seed']))

    def model():
        m = LinearClassifier(get_feature_extractor('conv31', units=16, bn=bn, unit_scaling=2))
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        return m.cuda().train()
    B, W, pi = int(g['B']), int(g['width']), float(g['pi'])
    Y = torch.from_numpy(g['Y'])
    X0 = torch.from_numpy(np.random.default_rng(4200).standard_normal((B, W, W)).astype(np.float32))
    m = model()
    T.flat_params(m)
    score = m(X0.cuda()).view(-1)
    masks = [(rec['v'] > 0).permute(0, 3, 1, 2).cpu() for rec in m.__dict__['_tpz_tape'] if rec['kind'] == 'conv']
    assert len(masks) == 3
    params = {k: torch.from_numpy(v).clone().requires_grad_('running' not in k and v.dtype == np.float32) for k, v in sd.items()}
    score_ref = O.classifier_forward_grad(params, X0, 'conv31', 16, bn=bn, relu_masks=masks, unit_scaling=2).view(-1)
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    _, _, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    T.backward(m, ds)
    errs = {k: max(rel_err(p.grad.cpu().numpy(), params[k].grad.numpy())) for k, p in m.named_parameters()}
    print({k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs
    # two full steps vs the real reference
    m = model()
    tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), pi)
    outs = []
    for step in range(2):
        X = torch.from_numpy(np.random.default_rng(4200 + step).standard_normal((B, W, W)).astype(np.float32)).cuda()
        outs.append(tr.step(X, Y.cuda()))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=2e-3, atol=1e-6)
    for k, v in m.state_dict().items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == 2
        else:
            assert_params_after_adam(v.detach().cpu().numpy(), g['p2.' + k], 2, 1e-3, k)


# ------------------------------------------------------------------------------------------------
# nn.Dropout in training (`topaz train --dropout p`, reference resnet.py:296-303)
# ------------------------------------------------------------------------------------------------
def test_dropout_kernels():
    from topaz_b200 import train_engine as T
    n, p = 1 << 20, 0.25
    x = torch.randn(n + 3, device='cuda')                                  # n % 4 != 0: ragged last Philox group
    torch.manual_seed(77)
    T._DROPOUT['calls'] = 10
    y, keep = T._dropout_fwd(x, p)
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n), frac       # Bernoulli(1-p) within 4 sigma
    assert torch.equal(y, x * (keep.float() / (1.0 - p)))                 # ATen's arithmetic: input * (mask / (1-p))
    T._DROPOUT['calls'] = 10
    y2, keep2 = T._dropout_fwd(x, p)
    assert torch.equal(keep, keep2) and torch.equal(y, y2)                # same (seed, offset) -> same mask
    y3, keep3 = T._dropout_fwd(x, p)                                      # next call: new stream position
    agree = (keep3 == keep).float().mean().item()
    assert abs(agree - (p * p + (1 - p) * (1 - p))) < 0.01                # independent of the previous mask
    torch.manual_seed(78)
    T._DROPOUT['calls'] = 10
    assert not torch.equal(T._dropout_fwd(x, p)[1], keep)                 # another seed -> another mask
    g = torch.randn(n + 3, device='cuda')
    ref = g * (keep.float() / (1.0 - p))
    T._dropout_bwd(g, keep, p)
    assert torch.equal(g, ref)


def test_dropout_training_gradients_match_oracle_with_imposed_masks():
    """ResNet8 (16 units, BatchNorm, dropout 0.25) in train() mode: the GPU draws its own Philox keep-masks; with those masks
    (and the ReLU masks, see the BatchNorm test) imposed on the oracle -- whose dropout handling is pinned to the real
    reference in tests/test_oracle_golden.py -- logits, loss and every gradient agree."""
    from common import seeded_state
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    g = gold('ge_binomial_u16_dropout')
    p, B, pi = float(g['p']), int(g['B']), float(g['pi'])
    m = LinearClassifier(get_feature_extractor('resnet8', units=16, bn=True, dropout=p))
    sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(g['seed']))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.cuda(); m.train()
    X = torch.from_numpy(np.random.default_rng(4300).standard_normal((B, 71, 71)).astype(np.float32))
    Y = torch.from_numpy(g['Y'])
    T.flat_params(m)
    score = m(X.cuda()).view(-1)
    relu_masks, drop_masks = [], []
    for rec in m.__dict__['_tpz_tape']:
        if rec['kind'] == 'conv':
            relu_masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
        elif rec['kind'] == 'resid':
            relu_masks.append((rec['h'] > 0).permute(0, 3, 1, 2).cpu())
            relu_masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
        elif rec['kind'] == 'dropout':
            drop_masks.append(rec['mask'].bool().permute(0, 3, 1, 2).cpu())
    assert len(relu_masks) == 8 and len(drop_masks) == 3
    assert all(abs(dm.float().mean().item() - (1 - p)) < 0.02 for dm in drop_masks[:2])
    params = {k: torch.from_numpy(v).clone().requires_grad_('running' not in k and v.dtype == np.float32) for k, v in sd.items()}
    score_ref = O.classifier_forward_grad(params, X, 'resnet8', 16, bn=True, relu_masks=relu_masks, dropout=p,
                                          dropout_masks=drop_masks).view(-1)
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    _, _, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    T.backward(m, ds)
    errs = {k: max(rel_err(p_.grad.cpu().numpy(), params[k].grad.numpy())) for k, p_ in m.named_parameters()}
    print({k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-3, errs
    # eval(): identity
    m.eval()
    with torch.no_grad():
        a = m(X[:4].cuda()); b = m(X[:4].cuda())
    assert torch.equal(a, b)


def test_resnet16_batchnorm_training_gradients_match_oracle():
    """ResNet16 (the default of `topaz extract`; 32 units, BatchNorm) in train() mode: stride-1 7x7 first layer, nine blocks
    with two strided residual blocks.  Logits and every gradient of a GE-binomial step vs autograd through the oracle with
    the GPU forward's ReLU masks imposed."""
    from common import seeded_state
    from common_shapes import classifier_shapes
    from topaz_b200 import train_engine as T
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet16', units=32, bn=True))
    sd = seeded_state(classifier_shapes('resnet16', 32, 1, True), 405)
    assert list(sd.keys()) == list(m.state_dict().keys())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.cuda(); m.train()
    B, W, pi = 6, m.width, 0.05
    assert W == 91
    X = torch.from_numpy(np.random.default_rng(4400).standard_normal((B, W, W)).astype(np.float32))
    Y = torch.tensor([1.0] + [0.0] * (B - 1), dtype=torch.float64)
    T.flat_params(m)
    score = m(X.cuda()).view(-1)
    masks = []
    for rec in m.__dict__['_tpz_tape']:
        if rec['kind'] == 'conv':
            masks.append((rec['y'] > 0).permute(0, 3, 1, 2).cpu())
        elif rec['kind'] == 'resid':
            masks += [(rec['h'] > 0).permute(0, 3, 1, 2).cpu(), (rec['y'] > 0).permute(0, 3, 1, 2).cpu()]
    assert len(masks) == 16
    params = {k: torch.from_numpy(v).clone().requires_grad_('running' not in k and v.dtype == np.float32) for k, v in sd.items()}
    score_ref = O.classifier_forward_grad(params, X, 'resnet16', 32, bn=True, relu_masks=masks).view(-1)
    assert max(rel_err(score.detach().cpu().numpy(), score_ref.detach().numpy())) < 1e-4
    _, _, loss = O.ge_binomial_loss(score_ref, Y, pi, 1.0)
    loss.backward()
    ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
    T.ge_loss_grad(score.contiguous(), Y.cuda(), pi, 1.0, 0, B, ds, o5)
    T.backward(m, ds)
    errs = {k: max(rel_err(p_.grad.cpu().numpy(), params[k].grad.numpy())) for k, p_ in m.named_parameters()}
    assert max(errs.values()) < 1e-3, errs
