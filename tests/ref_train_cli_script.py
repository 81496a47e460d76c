"""Runs the reference's `topaz train` command function (topaz/commands/train.py: add_arguments + main -> make_model ->
train_model -> make_data_iterators -> fit_epochs) on a tiny synthetic data set written with the reference's own MRC writer.
Only `topaz.*` names are imported, so the same code exercises whichever implementation those names resolve to: used by
tests/test_dropin_reference_cli.py in a subprocess (unmodified reference, CPU) and in-process after
topaz_b200.compat.install() (drop-in modules, simulated kernels)."""
import os
import random
import sys

import numpy as np
import torch


def pretend_gpu():
    """Take the reference's use_cuda=True code paths (GPU-side data loading, X.cuda()) on a machine without a GPU: device
    selection succeeds and .cuda() keeps tensors where they are."""
    import topaz.cuda
    topaz.cuda.set_device = lambda device, **kw: True
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self


def run(out_dir, device, extra=()):
    import topaz.mrc as mrc
    import topaz.commands.train as train_cmd
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    data = os.path.join(out_dir, 'data')
    os.makedirs(data, exist_ok=True)
    rng = np.random.default_rng(11)
    rows = ['image_name\tx_coord\ty_coord']
    for i in range(2):
        img = rng.standard_normal((160, 176)).astype(np.float32)
        pts = rng.integers(40, 120, size=(6, 2))
        for (x, y) in pts:
            img[y - 3:y + 4, x - 3:x + 4] += 2.0
            rows.append(f'mic{i}\t{x}\t{y}')
        with open(os.path.join(data, f'mic{i}.mrc'), 'wb') as f:
            mrc.write(f, img[np.newaxis])          # the writer expects [nz, ny, nx]
    targets = os.path.join(out_dir, 'targets.txt')
    with open(targets, 'w') as f:
        f.write('\n'.join(rows) + '\n')
    argv = ['--train-images', data, '--train-targets', targets, '--test-images', data, '--test-targets', targets,
            '-n', '20', '--num-epochs', '2', '--epoch-size', '2', '--minibatch-size', '8', '--no-pretrained', '--units', '16',
            '--save-prefix', os.path.join(out_dir, 'model'), '-o', os.path.join(out_dir, 'log.tsv'), '-d', str(device)] + list(extra)
    args = train_cmd.add_arguments().parse_args(argv)
    train_cmd.main(args)
    return os.path.join(out_dir, 'log.tsv')


if __name__ == '__main__':
    pretend_gpu()
    print(run(sys.argv[1], 0, sys.argv[2:]))
