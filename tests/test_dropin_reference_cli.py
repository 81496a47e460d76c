"""The drop-in claim, end to end, in the build container: the UNMODIFIED reference pipeline code
(`topaz.extract.score_images`, `topaz.denoise.Denoise`) runs on the topaz_b200 modules after
`topaz_b200.compat.install()`, with the kernels simulated on the CPU (tests/sim_backend.py).  Skipped where the
reference checkout is absent (e.g. on the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from common import ROOT, gold, rel_err
import sim_backend

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'topaz')), reason='reference checkout not present')


@pytest.fixture()
def aliased():
    from topaz_b200 import compat
    for k in [k for k in sys.modules if k == 'topaz' or k.startswith('topaz.')]:
        del sys.modules[k]
    sys.path.insert(0, REF); sys.path.insert(0, os.path.join(ROOT, 'tools', 'stubs'))
    sys.dont_write_bytecode = True
    compat.install()
    yield
    compat.uninstall()
    for k in [k for k in sys.modules if k == 'topaz' or k.startswith('topaz.')]:
        del sys.modules[k]
    sys.path.remove(REF); sys.path.remove(os.path.join(ROOT, 'tools', 'stubs'))


def test_reference_score_images_runs_on_dropin_modules(aliased, tmp_path, monkeypatch):
    import topaz.extract as ref_extract                      # the reference's own pipeline code
    import topaz.cuda
    from topaz_b200.model.classifier import LinearClassifier as Ours
    from topaz_b200 import mrc
    g = gold('resnet8_u32_pretrained')
    p = str(tmp_path / 'mic.mrc'); mrc.write(p, g['x'][0, 0])
    monkeypatch.setattr(topaz.cuda, 'set_device', lambda device, **kw: True)       # pretend a GPU is selected
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)       # ... and keep tensors where they are
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    with sim_backend.patched():
        out = list(ref_extract.score_images('resnet8_u32', [p], device=0))
    assert isinstance(ref_extract.load_model('resnet8_u32'), Ours)
    assert out[0][0] == p and out[0][1].dtype == np.float32
    mx, l2 = rel_err(out[0][1], g['y_dense'][0, 0])
    assert mx < 1e-3 and l2 < 1e-3, (mx, l2)


def test_reference_whole_module_pickle_loads_into_dropin_classes(aliased, tmp_path):
    """training.py:601 saves whole modules; those pickles must resolve to the drop-in classes and still work."""
    import importlib
    from topaz_b200 import compat
    compat.uninstall()
    for k in [k for k in sys.modules if k.startswith('topaz.model')]:
        del sys.modules[k]
    ref_factory = importlib.import_module('topaz.model.factory')           # the REAL reference classes
    m = ref_factory.load_model('resnet8_u32')
    path = str(tmp_path / 'model_epoch10.sav')
    torch.save(m, path)
    x = torch.from_numpy(gold('resnet8_u32_pretrained')['x'])
    for k in [k for k in sys.modules if k.startswith('topaz.model')]:
        del sys.modules[k]
    compat.install()
    from topaz_b200.model.classifier import LinearClassifier as Ours
    loaded = torch.load(path, weights_only=False)
    assert type(loaded) is Ours and type(loaded.features).__module__ == 'topaz_b200.model.features.resnet'
    loaded.eval(); assert loaded.fill() == 4
    with sim_backend.patched(), torch.no_grad():
        y = loaded(x).numpy()
    mx, l2 = rel_err(y, gold('resnet8_u32_pretrained')['y_dense'])
    assert mx < 1e-3 and l2 < 1e-3, (mx, l2)


def test_reference_denoise_pipeline_runs_on_dropin_unet(aliased, monkeypatch):
    """reference topaz.denoise.Denoise (its own _denoise / denoise_patches code) driving the drop-in UDenoiseNet."""
    import topaz.denoise as ref_denoise
    from topaz_b200.denoising.models import UDenoiseNet as Ours
    g = gold('unet_pretrained')
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    dn = ref_denoise.Denoise('unet', use_cuda=False)
    assert type(dn.model) is Ours
    with sim_backend.patched():
        y = dn.denoise(g['img'].copy(), patch_size=64, padding=24)
    mx, l2 = rel_err(y, g['y_pat'])
    assert mx < 1e-3 and l2 < 1e-3, (mx, l2)
