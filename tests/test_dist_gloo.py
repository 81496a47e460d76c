"""world_size-2 (gloo, CPU) test of the data-parallel GE-binomial step: each rank runs the product's host logic
(GE_binomial.step: all-gather of logits/labels, sharded loss gradient, flat-gradient all-reduce, fused Adam)
on half of the minibatch with simulated kernels; the result must equal the single-process reference golden.
Also checks the image / patch sharding helpers used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from common import gold, weights_of, rel_err, assert_params_after_adam


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, bn=False):
    import torch.distributed as dist
    import torch.nn as nn
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import sim_backend
    from topaz_b200.methods import GE_binomial
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    if bn:
        from common import seeded_state
        from common_shapes import classifier_shapes
        g = gold('ge_binomial_u32_bn'); sd = seeded_state(classifier_shapes('resnet8', 32, 1, True), int(g['seed']))
    else:
        g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=bn))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']))
    B = int(g['B']); b = B // world
    Y = torch.from_numpy(g['Y'])
    outs = []
    with sim_backend.patched_training():
        for step in range(1 if bn else 2):
            X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32))
            outs.append(tr.step(X[rank * b:(rank + 1) * b], Y[rank * b:(rank + 1) * b]))
    q.put((rank, outs, {k: p.detach().numpy().copy() for k, p in m.state_dict().items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_ge_binomial_matches_single_process():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    g = gold('ge_binomial_u32')
    # loss / ge / metrics are minibatch-global: identical on both ranks and equal to the single-process golden
    np.testing.assert_allclose(np.array(res[0][1]), g['outs'][:2], rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(np.array(res[0][1]), np.array(res[1][1]), rtol=1e-6)
    # replicas stay bit-identical after the all-reduced update
    for k in res[0][2]:
        assert np.array_equal(res[0][2][k], res[1][2][k]), k


def test_data_parallel_batchnorm_uses_global_minibatch_statistics():
    """BatchNorm model, two ranks x 32 crops: the fp64 per-channel sums are all-reduced in the forward (statistics) and in
    the backward (sum g, sum g*xhat), so one step equals the single-process step on the 64-crop minibatch: loss tuple and
    running buffers (forward quantities) tightly, parameters within the Adam sign-flip bound (see test_host_logic)."""
    from oracle import topaz_oracle as O
    from common import seeded_state
    from common_shapes import classifier_shapes
    g = gold('ge_binomial_u32_bn')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    np.testing.assert_allclose(np.array(res[0][1]), g['outs'][:1], rtol=1e-3, atol=1e-6)
    for k in res[0][2]:
        assert np.array_equal(res[0][2][k], res[1][2][k]), k          # replicas (buffers included) stay bit-identical
    # the golden stores the state after 3 steps; the single-process state after ONE step comes from the oracle (pinned to
    # that golden in tests/test_oracle_golden.py)
    sd = seeded_state(classifier_shapes('resnet8', 32, 1, True), int(g['seed']))
    X = np.random.default_rng(4000).standard_normal((int(g['B']), 71, 71)).astype(np.float32)
    _, _, final = O.ge_binomial_steps(sd, [X], [g['Y']], 'resnet8', 32, float(g['pi']), bn=True)
    for k, v in res[0][2].items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == 1
        elif 'running' in k:
            assert max(rel_err(v, final[k])) < 1e-5, k
        else:
            assert_params_after_adam(v, final[k], 1, 5e-4, k)


def test_sharding_helpers():
    from topaz_b200.parallel import shard_range, shard_list
    assert [shard_range(216, r, 8) for r in range(8)] == [(27 * r, 27 * (r + 1)) for r in range(8)]
    cover = [shard_range(10, r, 4) for r in range(4)]
    assert cover[0][0] == 0 and cover[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert shard_list(list(range(10)), 1, 4) == [1, 5, 9]
