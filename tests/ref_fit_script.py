"""Drives the reference's own training loop (topaz.training.make_training_step_method + fit_epochs: training.py:352-400,
506-603) on a seeded ResNet8(units=32, bn=True) -- the default `topaz train` model.  Used twice by
tests/test_dropin_reference_cli.py: in a subprocess against the unmodified reference, and in-process after
topaz_b200.compat.install() (drop-in modules, simulated kernels).  Only `topaz.*` names are imported here, so the same
code exercises whichever implementation those names resolve to."""
import io
import os
import sys

import numpy as np
import torch


def run(out_dir, seed=401, B=16, epochs=2, steps=2):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import seeded_state
    import topaz.training as training
    from topaz.model.factory import get_feature_extractor
    from topaz.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, dropout=0.0, bn=True, unit_scaling=2, pooling=None, dims=2))
    sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.train()
    trainer, criteria, split = training.make_training_step_method(m, 100, 0.01, lr=2e-4, l2=1e-5, method='GE-binomial',
                                                                  pi=0.05, slack=-1)
    rng = np.random.default_rng(seed)
    Y = torch.tensor([1.0] * 2 + [0.0] * (B - 2), dtype=torch.float64)
    train = [(torch.from_numpy(rng.standard_normal((B, 71, 71)).astype(np.float32)), Y.clone()) for _ in range(steps)]
    xt = torch.from_numpy(rng.standard_normal((1, 96, 80)).astype(np.float32))
    yt = torch.from_numpy((rng.random((1, 96, 80)) < 0.05).astype(np.float32))
    buf = io.StringIO()
    prefix = os.path.join(out_dir, 'model')
    training.fit_epochs(m, criteria, trainer, train, [(xt, yt)], epochs, 0.9, save_prefix=prefix, use_cuda=False, output=buf)
    with open(os.path.join(out_dir, 'log.tsv'), 'w') as f:
        f.write(buf.getvalue())
    np.savez(os.path.join(out_dir, 'final.npz'), **{k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    return type(m).__module__


if __name__ == '__main__':
    print(run(sys.argv[1]))
