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


def score_images(model: Union[torch.nn.Module, str], paths: Union[List[str], Iterable[str]], device: int = 0,
                 patch_size: int = 0, batch_size: int = 1) -> Iterator[Tuple[str, np.ndarray]]:
    if model is not None and model != 'none':
        if device is None or device < 0 or not torch.cuda.is_available():
            raise RuntimeError('topaz_b200.score_images needs a CUDA device (device >= 0); there is no CPU fallback')
        torch.cuda.set_device(device)
        if isinstance(model, str):
            model = load_model(model)
        model.eval()
        model.fill()
        model.cuda()
        for path in paths:
            image = load_image(path, make_image=False, return_header=False)
            is_3d = len(image.shape) == 3
            image = torch.from_numpy(np.ascontiguousarray(image)).float()
            image = image.unsqueeze(0).unsqueeze(0)
            if patch_size:
                patch_overlap = model.width // 2
                scores = predict_in_patches(model, image, patch_size + 2 * patch_overlap, is_3d=is_3d, use_cuda=True)
                scores = scores[0, 0]
            else:
                with torch.no_grad():
                    scores = model(image.cuda(non_blocking=True)).data[0, 0].cpu().numpy()
            yield path, scores
    else:
        for path in paths:
            yield path, load_image(path, make_image=False, return_header=False)


def score_arrays(model: torch.nn.Module, images: Iterable[np.ndarray], device: int = 0, pinned: bool = True):
    """Array-in / array-out variant used by bench.py's end-to-end leg: double-buffered pinned H2D / D2H on
    side streams so copies of image i+1 / i-1 overlap the network of image i."""
    torch.cuda.set_device(device)
    model.eval(); model.fill(); model.cuda()
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
