"""Probe: how does tcgen05.mma (kind::tf32) read an MN-major SWIZZLE_128B operand?  One MMA per experiment; the probed operand's
shared-memory image holds its own linear index (i + 1) so that D against an identity operand reveals every address read.
Writes gpurun_out/lab_mn_major.json."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools', 'lab'))
import lab


def desc(start, lbo, sbo, layout=2):
    return (start >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def swz(addr):                      # SWIZZLE_128B: 16-byte piece index ^= row & 7
    return addr ^ (((addr >> 7) & 7) << 4)


def swz32(addr):                    # SWIZZLE_128B_BASE32B (layout type 1): 32-byte piece index ^= row & 3
    return addr ^ (((addr >> 7) & 3) << 5)


def kmajor_identity(rows):
    """K-major SWIZZLE_128B image of X[r][k] = (r == k), rows x 128 B (k < 8 used)"""
    img = np.zeros(rows * 32, dtype=np.float32)
    for r in range(min(rows, 8)):
        k = r
        img[swz(r * 128 + k * 4) // 4] = 1.0
    return img


WINDOW = 2048       # floats: indices 1..2048 are exact in tf32
out = {}
index_img = np.zeros(16384, dtype=np.float32); index_img[:WINDOW] = np.arange(1, WINDOW + 1)
dev = 'cuda'
# ---- A probed (MN-major), B = K-major identity, N = 32
Bid = torch.from_numpy(kmajor_identity(32)).to(dev)
Aimg = torch.from_numpy(index_img).to(dev)
for (layout, start, lbo, sbo) in [(2, 0, 1024, 1024), (1, 0, 1024, 512), (1, 0, 512, 2048), (1, 0, 2048, 512), (1, 0, 128, 512), (1, 384, 1024, 512),
                                  (1, 384, 128, 512), (1, 0, 256, 512), (1, 1024 + 640, 256, 512), (1, 128, 128, 512), (1, 256, 128, 512)]:
    D = lab.lab_umma_raw(Aimg, Bid, desc(start, lbo, sbo, layout), desc(65536, 0, 1024), idesc(128, 32, 1, 0), 32).cpu().numpy()
    got = D[:, :8].astype(np.int64) - 1                      # float index read for (m, k); -1 = outside the window / zero
    hyp = np.zeros((128, 8), dtype=np.int64)
    for m in range(128):
        for k in range(8):
            if layout == 2:
                hyp[m, k] = swz(start + (m >> 5) * lbo + k * 128 + (m & 31) * 4) // 4
            else:
                hyp[m, k] = swz32(start + (m >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 + (m & 31) * 4) // 4
    hyp[hyp >= WINDOW] = -1
    match = float((got == hyp).mean())
    name = f'A_mn layout={layout} start={start} lbo={lbo} sbo={sbo}'
    print(name, 'match with hypothesis', match, 'rest of D zero:', bool((D[:, 8:] == 0).all()))
    if match < 1:
        for m in (0, 1, 4, 5, 8, 9, 32, 33, 64, 96):
            print('   m', m, 'got bytes', (got[m] * 4).tolist(), 'hyp', (hyp[m] * 4).tolist())
    out[name] = dict(match=match, got=got.tolist())
# ---- B probed (MN-major), A = K-major identity (rows 0..7)
Aid = torch.from_numpy(kmajor_identity(128)).to(dev)
Bimg = torch.from_numpy(index_img).to(dev)
for (N, start, lbo, sbo) in [(32, 0, 1024, 512), (64, 0, 2048, 512), (64, 0, 1024, 2048), (64, 1024, 2048, 512), (64, 1024, 4096, 512)]:
    D = lab.lab_umma_raw(Aid, Bimg, desc(0, 0, 1024), desc(65536 + start, lbo, sbo, 1), idesc(128, N, 0, 1), N).cpu().numpy()
    got = D[:8, :].T.astype(np.int64) - 1                    # [n][k]
    hyp = np.zeros((N, 8), dtype=np.int64)
    for n in range(N):
        for k in range(8):
            hyp[n, k] = swz32(start + (n >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 + (n & 31) * 4) // 4
    hyp[hyp >= WINDOW] = -1
    match = float((got == hyp).mean())
    name = f'B_mn N={N} start={start} lbo={lbo} sbo={sbo}'
    print(name, 'match with hypothesis', match)
    if match < 1:
        for n in (0, 1, 4, 5, 32, 33):
            if n < N:
                print('   n', n, 'got bytes', (got[n] * 4).tolist(), 'hyp', (hyp[n] * 4).tolist())
    out[name] = dict(match=match, got=got.tolist())
# ---- control: A K-major probed the same way (known behaviour)
D = lab.lab_umma_raw(Aimg, Bid, desc(0, 0, 1024), desc(65536, 0, 1024), idesc(128, 32, 0, 0), 32).cpu().numpy()
got = D[:, :8].astype(np.int64) - 1
hyp = np.array([[swz(m * 128 + k * 4) // 4 for k in range(8)] for m in range(128)]); hyp[hyp >= WINDOW] = -1
print('control A K-major match', float((got == hyp).mean()))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'lab_mn_major.json'), 'w'))
