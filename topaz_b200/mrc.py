"""Minimal MRC2014 reader/writer used by score_images and the benchmarks (the reference's full codec,
topaz/mrc.py, is host-side I/O outside the B200 hot path and is reused unchanged when topaz is installed)."""
import struct
import numpy as np

_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}


def parse(content: bytes):
    nx, ny, nz, mode = struct.unpack('<4i', content[:16])
    nsymbt = struct.unpack('<i', content[92:96])[0]
    if mode not in _MODES:
        raise ValueError(f'unsupported MRC mode {mode}')
    dtype = np.dtype(_MODES[mode])
    start = 1024 + nsymbt
    arr = np.frombuffer(content, dtype=dtype, count=nx * ny * nz, offset=start)
    arr = arr.reshape((nz, ny, nx)) if nz > 1 else arr.reshape((ny, nx))
    return arr, dict(nx=nx, ny=ny, nz=nz, mode=mode)


def read(path):
    with open(path, 'rb') as f:
        return parse(f.read())


def write(f, array: np.ndarray):
    a = np.ascontiguousarray(array, dtype=np.float32)
    if a.ndim == 2:
        nz, (ny, nx) = 1, a.shape
    else:
        nz, ny, nx = a.shape
    hdr = bytearray(1024)
    struct.pack_into('<4i', hdr, 0, nx, ny, nz, 2)
    struct.pack_into('<3i', hdr, 28, nx, ny, nz)          # mx, my, mz
    struct.pack_into('<3f', hdr, 40, float(nx), float(ny), float(nz))   # cell
    struct.pack_into('<3f', hdr, 52, 90.0, 90.0, 90.0)
    struct.pack_into('<3i', hdr, 64, 1, 2, 3)            # mapc, mapr, maps
    struct.pack_into('<3f', hdr, 76, float(a.min()), float(a.max()), float(a.mean()))
    hdr[208:212] = b'MAP '
    hdr[212:216] = bytes([0x44, 0x44, 0x00, 0x00])       # little-endian machine stamp
    struct.pack_into('<f', hdr, 216, float(a.std()))
    own = isinstance(f, str)
    fh = open(f, 'wb') if own else f
    fh.write(bytes(hdr)); fh.write(a.tobytes())
    if own:
        fh.close()


def load_image(path, make_image=False, return_header=False):
    """Mirror of topaz.utils.data.loader.load_image for the formats used on the hot path."""
    ext = path.rsplit('.', 1)[-1].lower()
    if ext == 'mrc':
        arr, hdr = read(path)
        arr = np.array(arr)
        return (arr, hdr) if return_header else arr
    if ext == 'npy':
        return np.load(path)
    from PIL import Image
    im = np.array(Image.open(path), copy=False)
    return im
