"""The drop-in claim, end to end, in the build container: the UNMODIFIED reference pipeline code
(`topaz.extract.score_images`, `topaz.denoise.Denoise`) runs on the topaz_b200 modules after
`topaz_b200.compat.install()`, with the kernels simulated on the CPU (tests/sim_backend.py).  Skipped where the
reference checkout is absent (e.g. on the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from common import ROOT, gold, weights_of, rel_err, assert_params_after_adam
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


def test_reference_fit_epochs_runs_on_dropin_modules(aliased, tmp_path):
    """The reference's own training loop (training.py: make_training_step_method -> fit_epochs -> fit_epoch /
    evaluate_model -> torch.save of the whole module) on the default `topaz train` model (ResNet8, 32 units, BatchNorm ON,
    l2 > 0): once in a subprocess on the unmodified reference, once in-process on the drop-in modules (simulated kernels).
    Same log lines (train: loss, GE penalty, precision, tpr, fpr; test: loss, ..., auprc), same final state, and the saved
    .sav holds no engine caches and still loads / evaluates."""
    import subprocess
    import ref_fit_script
    ref_dir, our_dir = tmp_path / 'ref', tmp_path / 'ours'
    ref_dir.mkdir(); our_dir.mkdir()
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=os.pathsep.join([REF, os.path.join(ROOT, 'tools', 'stubs')]))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_fit_script.py'), str(ref_dir)], env=env, cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == 'topaz.model.classifier'
    with sim_backend.patched_training():
        owner = ref_fit_script.run(str(our_dir))
    assert owner == 'topaz_b200.model.classifier'

    def table(p):
        rows = [l.rstrip('\n').split('\t') for l in open(p)]
        return rows[0], rows[1:]
    (h0, ref_rows), (h1, our_rows) = table(ref_dir / 'log.tsv'), table(our_dir / 'log.tsv')
    assert h0 == h1 and len(ref_rows) == len(our_rows) == 6
    for a, b in zip(ref_rows, our_rows):
        assert a[:3] == b[:3]
        for u, v in zip(a[3:], b[3:]):
            if u == '-':
                assert v == '-'
            else:
                # train lines: fp32-level; test lines: dense fp16-operand forward of He-random weights (TOL_SEEDED-like)
                assert abs(float(u) - float(v)) <= (2e-3 if a[2] == 'train' else 1e-2) * max(abs(float(u)), 1e-3), (a, b)
    fr, fo = np.load(ref_dir / 'final.npz'), np.load(our_dir / 'final.npz')
    assert fr.files == fo.files
    for k in fr.files:
        if k.endswith('num_batches_tracked'):
            assert int(fr[k]) == int(fo[k]) == 4
            continue
        if 'running' in k:
            assert max(rel_err(fo[k], fr[k])) < 1e-3, k
        else:
            assert_params_after_adam(fo[k], fr[k], 4, 1e-3, k)
    # the whole-module pickle written by training.py:600-601 carries parameters and buffers only
    for ep in (1, 2):
        saved = torch.load(str(our_dir / f'model_epoch{ep}.sav'), weights_only=False)
        assert not [k for mod in saved.modules() for k in mod.__dict__ if k.startswith('_tpz_')]
        assert abs(os.path.getsize(our_dir / f'model_epoch{ep}.sav') - os.path.getsize(ref_dir / f'model_epoch{ep}.sav')) < 65536
    saved.eval(); saved.fill()
    with sim_backend.patched(), torch.no_grad():
        y = saved(torch.zeros(1, 1, 80, 80))
    assert y.shape == (1, 1, 80, 80) and torch.isfinite(y).all()


def test_adam_state_survives_a_device_round_trip():
    """training.py:600-603 moves the classifier to the CPU for torch.save and back after every epoch: the parameters get
    new storage, the flat buffers must be rebuilt around them WITHOUT resetting the Adam moments or the step count."""
    import torch.nn as nn
    from topaz_b200.methods import GE_binomial
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier

    def run(round_trip):
        torch.manual_seed(0)
        m = LinearClassifier(get_feature_extractor('resnet8', units=16, bn=True)); m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        tr = GE_binomial(m, opt, nn.BCEWithLogitsLoss(), 0.05)
        Y = torch.tensor([1.0] * 2 + [0.0] * 6, dtype=torch.float64)
        outs = []
        with sim_backend.patched_training():
            for step in range(4):
                X = torch.from_numpy(np.random.default_rng(step).standard_normal((8, 71, 71)).astype(np.float32))
                outs.append(tr.step(X, Y))
                if round_trip and step == 1:
                    for p in m.parameters():          # what nn.Module.cpu().cuda() does to every parameter
                        p.data = p.data.clone()
                        if p.grad is not None:
                            p.grad.data = p.grad.data.clone()
        st = opt.state[next(iter(m.parameters()))]
        return np.array(outs), {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}, float(st['step'])
    o0, s0, n0 = run(False)
    o1, s1, n1 = run(True)
    assert n0 == n1 == 4.0
    np.testing.assert_array_equal(o0, o1)
    for k in s0:
        np.testing.assert_array_equal(s0[k], s1[k], err_msg=k)


def test_reference_train_command_runs_on_dropin_modules(aliased, tmp_path, monkeypatch):
    """`topaz train` end to end: the reference's command function (commands/train.py main: make_model -> train_model ->
    its own data loaders, crop sampler and augmentation -> fit_epochs with a test split and --save-prefix) drives the drop-in
    modules (simulated kernels) on a tiny data set with a pretended GPU, so that the reference takes its use_cuda=True paths.
    The reference's samplers draw from unseeded generators (two runs of the reference itself see different minibatches), so
    this test checks the run structurally; the numeric comparison of the same loop on fixed minibatches is
    test_reference_fit_epochs_runs_on_dropin_modules.  With --minibatch-balance 0.0625 x 8 crops most minibatches hold no
    positive: the classifier loss is then nan in the log -- in the reference too (mean over an empty selection) -- while
    the GE term still trains, and the parameters must stay finite."""
    import ref_train_cli_script
    monkeypatch.setattr(torch.nn.Module, 'cuda', torch.nn.Module.cuda)             # restored after pretend_gpu() below
    monkeypatch.setattr(torch.Tensor, 'cuda', torch.Tensor.cuda)
    import topaz.cuda
    monkeypatch.setattr(topaz.cuda, 'set_device', topaz.cuda.set_device)
    ref_train_cli_script.pretend_gpu()
    with sim_backend.patched_training():
        log = ref_train_cli_script.run(str(tmp_path), 0)
    rows = [l.rstrip('\n').split('\t') for l in open(log)]
    assert rows[0] == ['epoch', 'iter', 'split', 'loss', 'ge_penalty', 'precision', 'adjusted_precision', 'tpr', 'fpr', 'auprc']
    body = rows[1:]
    assert [r[2] for r in body] == ['train', 'train', 'test', 'train', 'train', 'test']
    for r in body:
        if r[2] == 'train':
            ge, prec, fpr = float(r[4]), float(r[5]), float(r[8])
            assert np.isfinite(ge) and ge > 0 and 0.0 <= prec <= 1.0 and 0.0 < fpr < 1.0 and r[9] == '-', r
        else:
            assert np.isfinite(float(r[3])) and r[4] == '-' and 0.0 <= float(r[9]) <= 1.0, r
    from topaz_b200.model.classifier import LinearClassifier as Ours
    for ep in (1, 2):
        saved = torch.load(str(tmp_path / f'model_epoch{ep}.sav'), weights_only=False)
        assert type(saved) is Ours and not [k for mod in saved.modules() for k in mod.__dict__ if k.startswith('_tpz_')]
        assert all(torch.isfinite(v).all() for v in saved.state_dict().values())
    assert int(saved.state_dict()['features.features.0.bn.num_batches_tracked']) == 4


def test_reference_extract_command_runs_on_dropin_modules(aliased, tmp_path, monkeypatch):
    """`topaz extract` end to end: the reference's command function (commands/extract.py main -> extract_particles ->
    score_images -> nms_iterator -> coordinate table) on the drop-in classifier (simulated conv kernels).  The build container
    has no GPU for the NMS kernel, so the reference module's NMS name is bound to the oracle's (the GPU NMS itself is held
    bit-exact to that oracle in tests/test_gpu_parity.py); what is checked here is that the reference pipeline scores through
    the drop-in modules and writes the picks of THAT score map."""
    import topaz.cuda
    import topaz.extract as ref_extract
    import topaz.commands.extract as extract_cmd
    from oracle import topaz_oracle as O
    from topaz_b200 import mrc
    g = gold('resnet8_u32_pretrained')
    p = str(tmp_path / 'mic.mrc'); mrc.write(p, g['x'][0, 0])
    monkeypatch.setattr(topaz.cuda, 'set_device', lambda device, **kw: True)
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(ref_extract, 'non_maximum_suppression', lambda x, r, threshold=-np.inf: O.nms(x, r, threshold))
    out = str(tmp_path / 'picks.txt')
    args = extract_cmd.add_arguments().parse_args([p, '-m', 'resnet8_u32', '-r', '8', '-t', '-6', '-o', out, '-d', '0'])
    with sim_backend.patched():
        extract_cmd.main(args)
        scores = dict(ref_extract.score_images('resnet8_u32', [p], device=0))[p]
    rows = [l.rstrip('\n').split('\t') for l in open(out)]
    assert rows[0] == ['image_name', 'x_coord', 'y_coord', 'score']
    sc, coords = O.nms(scores, 8, -6)
    assert len(rows) - 1 == len(sc) > 0
    for r, s, (x, y) in zip(rows[1:], sc, coords):
        assert r[0] == 'mic' and int(r[1]) == x and int(r[2]) == y and abs(float(r[3]) - s) < 1e-5
    # the best pick is also the best pick of the reference's own score map
    ref_sc, ref_coords = O.nms(g['y_dense'][0, 0], 8, -6)
    assert tuple(ref_coords[0]) == tuple(coords[0]) and abs(ref_sc[0] - sc[0]) < 1e-3 * np.abs(g['y_dense']).max()


def test_reference_denoise_command_runs_on_dropin_modules(aliased, tmp_path, monkeypatch):
    """`topaz denoise` end to end: the reference's command function (commands/denoise.py main -> Denoise('unet') ->
    denoise_stream -> denoise_image -> Denoise.denoise in patches -> MRC written) with the drop-in UDenoiseNet behind the
    reference's own Denoise class (simulated kernels, pretended GPU); the written micrograph equals the oracle's result of
    the same pipeline (normalise, patched denoise with the packaged v0.2.2 weights, de-normalise)."""
    import topaz.cuda
    import topaz.commands.denoise as denoise_cmd
    from oracle import topaz_oracle as O
    from topaz_b200 import mrc
    from topaz_b200.denoising.models import UDenoiseNet as Ours
    g = gold('unet_pretrained')
    img = g['img']
    p = str(tmp_path / 'mic.mrc'); mrc.write(p, img)
    monkeypatch.setattr(denoise_cmd, 'set_device', lambda device, **kw: True)
    monkeypatch.setattr(topaz.cuda, 'set_device', lambda device, **kw: True)
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    outdir = str(tmp_path / 'out')
    args = denoise_cmd.add_arguments().parse_args([p, '-o', outdir, '-s', '64', '-p', '24', '-d', '0'])
    with sim_backend.patched():
        denoised = denoise_cmd.main(args)
    assert len(denoised) == 1 and denoised[0].shape == img.shape
    written = mrc.parse(open(os.path.join(outdir, 'mic.mrc'), 'rb').read())[0]
    written = written[0] if written.ndim == 3 else written
    mu, std = img.mean(), img.std()
    ref = std * O.denoise(weights_of(g), (img - mu) / std, 64, 24) + mu
    for y in (denoised[0], written):
        mx, l2 = rel_err(y, ref)
        assert mx < 2e-3 and l2 < 2e-3, (mx, l2)


def test_reference_denoise3d_command_runs_on_dropin_modules(aliased, tmp_path):
    """`topaz denoise3d` end to end: the reference's command function (commands/denoise3d.py main -> Denoise3D('unet-3d') ->
    denoise_tomogram_stream -> PatchDataset crops -> model -> MRC written) with the drop-in UDenoiseNet3D (packaged
    unet-3d-10a weights, simulated kernels); the written tomogram equals the oracle's result of the same pipeline."""
    import topaz.commands.denoise3d as denoise3d_cmd
    import topaz.mrc as ref_mrc
    from oracle import topaz_oracle as O
    from topaz_b200.denoising.models import UDenoiseNet3D as Ours, load_model
    tomo = gold('unet3d_seeded')['tomo']
    p = str(tmp_path / 'tomo.mrc')
    with open(p, 'wb') as f:
        ref_mrc.write(f, tomo)
    outdir = str(tmp_path / 'out'); os.makedirs(outdir)
    args = denoise3d_cmd.add_arguments().parse_args([p, '-o', outdir, '-s', '16', '-p', '8', '-d', '-1'])
    with sim_backend.patched():
        denoise3d_cmd.main(args)
    with open(os.path.join(outdir, 'tomo.mrc'), 'rb') as f:
        written = ref_mrc.parse(f.read())[0]
    m = load_model('unet-3d')
    assert type(m) is Ours
    ref = O.denoise3d({k: v.numpy() for k, v in m.state_dict().items()}, tomo, 16, 8)
    mx, l2 = rel_err(written, ref)
    assert written.shape == tomo.shape and mx < 2e-3 and l2 < 2e-3, (mx, l2)


def test_reference_preprocess_command_runs_on_dropin_kernels(aliased, tmp_path, monkeypatch):
    """`topaz preprocess -s 2` end to end: the reference's command function (commands/normalize.py main -> normalize_images ->
    Normalize.__call__: load, downsample, GMM-normalise, write MRC + metadata) with compat's function patches
    (`topaz.utils.image.downsample`, `topaz.stats.normalize` -> the Fourier-crop GEMMs and the one-pass-per-iteration EM of
    tpz_preproc.cu, simulated here) vs the unmodified reference command run in a subprocess on the CPU."""
    import json
    import subprocess
    import topaz.cuda
    import topaz.stats
    import topaz.commands.normalize as normalize_cmd
    from topaz_b200 import mrc, preprocess
    assert topaz.stats.downsample is preprocess.downsample          # the name `Normalize.__call__` resolves (stats.py:304)
    x = gold('preprocess')['x']
    p = str(tmp_path / 'mic.mrc'); mrc.write(p, x)
    ref_dir, our_dir = str(tmp_path / 'ref'), str(tmp_path / 'ours')
    common = [p, '-s', '2', '--sample', '1', '--metadata']
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=os.pathsep.join([REF, os.path.join(ROOT, 'tools', 'stubs')]))
    r = subprocess.run([sys.executable, os.path.join(REF, 'topaz', 'commands', 'normalize.py')] + common + ['-o', ref_dir, '-d', '-1'],
                       env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    monkeypatch.setattr(normalize_cmd, 'set_device', lambda device, **kw: True)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    with sim_backend.patched():
        normalize_cmd.main(normalize_cmd.add_arguments().parse_args(common + ['-o', our_dir, '-d', '0']))
    ours, ref = mrc.read(os.path.join(our_dir, 'mic.mrc'))[0], mrc.read(os.path.join(ref_dir, 'mic.mrc'))[0]
    ours, ref = np.squeeze(ours), np.squeeze(ref)
    assert ours.shape == ref.shape == (75, 66)
    assert np.abs(ours - ref).max() < 2e-3, np.abs(ours - ref).max()          # normalised micrograph: unit variance
    mo, mr = json.load(open(os.path.join(our_dir, 'mic.metadata.json'))), json.load(open(os.path.join(ref_dir, 'mic.metadata.json')))
    assert set(mo) == set(mr)
    for k in ('mu', 'std', 'pi', 'logp'):
        assert abs(mo[k] - mr[k]) <= 2e-3 * max(1.0, abs(mr[k])), (k, mo[k], mr[k])


def test_reference_segment_command_runs_on_dropin_modules(aliased, tmp_path, monkeypatch):
    """`topaz segment` end to end: the reference's command function (commands/segment.py main -> load_model -> eval/fill ->
    the reference's own segment_images -> TIFF score map) on the drop-in classifier (simulated kernels)."""
    import topaz.cuda
    import topaz.commands.segment as segment_cmd
    from PIL import Image
    from topaz_b200 import mrc
    g = gold('resnet8_u32_pretrained')
    p = str(tmp_path / 'mic.mrc'); mrc.write(p, g['x'][0, 0])
    monkeypatch.setattr(topaz.cuda, 'set_device', lambda device, **kw: True)
    monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    outdir = str(tmp_path / 'seg')
    with sim_backend.patched():
        segment_cmd.main(segment_cmd.add_arguments().parse_args([p, '-m', 'resnet8_u32', '-o', outdir, '-d', '0']))
    y = np.array(Image.open(os.path.join(outdir, 'mic.tiff')))
    mx, l2 = rel_err(y, g['y_dense'][0, 0])
    assert y.dtype == np.float32 and mx < 1e-3 and l2 < 1e-3, (mx, l2)


def test_python_m_topaz_b200_runs_the_reference_command_line(tmp_path):
    """`python -m topaz_b200 train --describe --no-pretrained`: the reference's own dispatcher (topaz/main.py) with the
    drop-in modules installed prints the default BatchNorm ResNet8 built from the drop-in classes."""
    import subprocess
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=os.pathsep.join([ROOT, REF, os.path.join(ROOT, 'tools', 'stubs')]))
    r = subprocess.run([sys.executable, '-m', 'topaz_b200', 'train', '--describe', '--no-pretrained'], env=env, cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.startswith('LinearClassifier(') and '(bn1): BatchNorm2d(64' in r.stdout and 'ResidA(' in r.stdout
    r = subprocess.run([sys.executable, '-c', 'import topaz_b200.compat as c; c.install(); import topaz.model.classifier as m; '
                        'print(m.LinearClassifier.__module__)'], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.stdout.strip() == 'topaz_b200.model.classifier', (r.stdout, r.stderr[-500:])
