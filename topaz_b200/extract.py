"""Drop-in mirror of topaz.extract.score_images (reference extract.py:224-256): stream micrographs through
the dense classifier.  H2D copy of the image, the whole network, and the D2H copy of the score map are the
only per-image work; there is no CPU model path (device must be >= 0)."""
from __future__ import division, print_function

from typing import Iterable, Iterator, List, Tuple, Union

import numpy as np
import torch

from topaz_b200.model.factory import load_model
from topaz_b200.model.utils import predict_in_patches
from topaz_b200.mrc import load_image


def _ensure_filled(model):
    """Dense mode for scoring (reference extract.py:231-232 calls fill() unconditionally on a freshly loaded model; a
    second fill() would multiply the ResidA dilations again -- in the reference too -- so an already filled model is left
    as it is)."""
    from topaz_b200.engine import is_filled
    model.eval()
    if not is_filled(model.features):
        model.fill()
    return model


def score_images(model: Union[torch.nn.Module, str], paths: Union[List[str], Iterable[str]], device: int = 0,
                 patch_size: int = 0, batch_size: int = 1) -> Iterator[Tuple[str, np.ndarray]]:
    if model is not None and model != 'none':
        if device is None or device < 0 or not torch.cuda.is_available():
            raise RuntimeError('topaz_b200.score_images needs a CUDA device (device >= 0); there is no CPU fallback')
        torch.cuda.set_device(device)
        if isinstance(model, str):
            model = load_model(model)
        _ensure_filled(model)
        model.cuda()
        if patch_size:
            for path, image in _prefetch(paths):
                is_3d = len(image.shape) == 3
                image = torch.from_numpy(np.ascontiguousarray(image)).float()
                image = image.unsqueeze(0).unsqueeze(0)
                patch_overlap = model.width // 2
                scores = predict_in_patches(model, image, patch_size + 2 * patch_overlap, is_3d=is_3d, use_cuda=True)
                yield path, scores[0, 0]
        else:
            # file reads (background thread), pinned H2D / D2H (side streams) and the network overlap across images
            names = []

            def images():
                for path, image in _prefetch(paths):
                    names.append(path)
                    yield image
            for i, scores in enumerate(_score_stream(model, images())):
                yield names[i], scores
    else:
        for path in paths:
            yield path, load_image(path, make_image=False, return_header=False)


def _prefetch(paths, depth: int = 2):
    """Read micrographs one ahead of the consumer on a background thread (file I/O and numpy decoding release the GIL)."""
    import queue
    import threading
    q = queue.Queue(maxsize=depth)
    end = object()

    def reader():
        try:
            for path in paths:
                q.put((path, load_image(path, make_image=False, return_header=False)))
            q.put(end)
        except BaseException as e:           # surface I/O errors in the consumer
            q.put(e)
    threading.Thread(target=reader, daemon=True).start()
    while True:
        item = q.get()
        if item is end:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


def score_arrays(model: torch.nn.Module, images: Iterable[np.ndarray], device: int = 0, pinned: bool = True):
    """Array-in / array-out variant used by bench.py's end-to-end leg: double-buffered pinned H2D / D2H on
    side streams so copies of image i+1 / i-1 overlap the network of image i."""
    torch.cuda.set_device(device)
    _ensure_filled(model); model.cuda()
    return _score_stream(model, images, pinned)


def _score_stream(model: torch.nn.Module, images: Iterable[np.ndarray], pinned: bool = True):
    copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    pending = None
    for img in images:
        h = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32))
        if pinned and not h.is_pinned():
            h = h.pin_memory()
        with torch.cuda.stream(copy_in):
            d = h.cuda(non_blocking=True)
        main.wait_stream(copy_in)
        with torch.no_grad():
            s = model(d[None, None])[0, 0]
        d.record_stream(main)
        copy_out.wait_stream(main)
        with torch.cuda.stream(copy_out):
            o = torch.empty(s.shape, dtype=torch.float32, pin_memory=True)
            o.copy_(s, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_out)
        s.record_stream(copy_out)
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].numpy()
        pending = (o, ev)
    if pending is not None:
        pending[1].synchronize()
        yield pending[0].numpy()


def nms_iterator(paths_scores, radius: int, threshold: float, pool=None, dims: int = 2, patch_size: int = 0,
                 patch_overlap: int = 0, verbose: bool = False):
    """Drop-in for topaz.extract.nms_iterator (reference extract.py:91-103) on the GPU NMS: yields (name, scores, coords).
    `pool` is accepted and ignored (the reference parallelises its Python loop over CPU processes; one GPU call replaces it).
    The patched variant of the reference (`patch_size > 0`, extract.py:44-75) unpacks three values from the two that
    non_maximum_suppression returns and cannot run; it is not reproduced."""
    from topaz_b200.algorithms import non_maximum_suppression, non_maximum_suppression_3d
    if patch_size:
        raise NotImplementedError('topaz_b200.extract.nms_iterator: patched NMS is not supported (see docstring)')
    nms = non_maximum_suppression if dims == 2 else non_maximum_suppression_3d
    for name, score in paths_scores:
        s, c = nms(score, radius, threshold=threshold)
        yield name, s, c


def pick_arrays(model: torch.nn.Module, images: Iterable[np.ndarray], radius: int, threshold: float = -6.0, device: int = 0):
    """Score micrographs and run the greedy NMS without the score map leaving the GPU: yields (scores float32 [j],
    coords int32 [j,2] as (x, y)) per image -- `topaz extract` (extract.py:224-256 followed by extract.py:91-103) minus
    the 2 x 64 MB of PCIe traffic per 4096^2 micrograph and the reference's multi-second Python NMS loop."""
    from topaz_b200.algorithms import non_maximum_suppression
    torch.cuda.set_device(device)
    _ensure_filled(model); model.cuda()
    copy_in = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    nxt = None
    it = iter(images)

    def stage(img):
        h = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32))
        if not h.is_pinned():
            h = h.pin_memory()
        with torch.cuda.stream(copy_in):
            d = h.cuda(non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_in)
        return d, ev, h
    try:
        nxt = stage(next(it))
    except StopIteration:
        return
    while nxt is not None:
        d, ev, keep = nxt
        try:
            nxt = stage(next(it))            # H2D of image i+1 overlaps the network + NMS of image i
        except StopIteration:
            nxt = None
        main.wait_event(ev)
        with torch.no_grad():
            s = model(d[None, None])[0, 0]
        d.record_stream(main)
        yield non_maximum_suppression(s, radius, threshold)
