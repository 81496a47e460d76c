"""state_dict key -> shape tables of the reference architectures (written from the reference's
constructors: resnet.py:50-339, basic.py:12-78, classifier.py:14-30, denoising/models.py:74-128,452-506)."""
from collections import OrderedDict


def _bn(d, p, c):
    d[p + '.weight'] = (c,); d[p + '.bias'] = (c,)
    d[p + '.running_mean'] = (c,); d[p + '.running_var'] = (c,); d[p + '.num_batches_tracked'] = ()


def classifier_shapes(arch, units, scaling=1, bn=False):
    d = OrderedDict()
    if arch in ('resnet8', 'resnet16'):
        u0, u1, u2 = units, 2 * units, 4 * units
        if arch == 'resnet8':
            blocks = [('c', 1, u0, 7), ('r', u0, u0), ('r', u0, u1), ('r', u1, u1), ('c', u1, u2, 5)]
        else:
            blocks = [('c', 1, u0, 7)] + [('r', u0, u0)] * 4 + [('r', u0, u1), ('r', u1, u1), ('r', u1, u1), ('c', u1, u2, 5)]
        for i, b in enumerate(blocks):
            p = f'features.features.{i}.'
            if b[0] == 'c':
                d[p + 'conv.weight'] = (b[2], b[1], b[3], b[3])
                if not bn:
                    d[p + 'conv.bias'] = (b[2],)
                else:
                    _bn(d, p + 'bn', b[2])
            else:
                nin, nout = b[1], b[2]
                if nin != nout:
                    d[p + 'proj.weight'] = (nout, nin, 1, 1)
                d[p + 'conv0.weight'] = (nin, nin, 3, 3)
                if not bn:
                    d[p + 'conv0.bias'] = (nin,)
                else:
                    _bn(d, p + 'bn0', nin)
                d[p + 'conv1.weight'] = (nout, nin, 3, 3)
                if not bn:
                    d[p + 'conv1.bias'] = (nout,)
                else:
                    _bn(d, p + 'bn1', nout)
        latent = u2
    else:
        sizes = {'conv31': [7, 5, 5], 'conv63': [7, 5, 5, 5], 'conv127': [7, 5, 5, 5, 5]}[arch]
        idx, nin, u = 0, 1, units
        for j, k in enumerate(sizes):
            p = f'features.features.{idx}'
            d[p + '.weight'] = (u, nin, k, k)
            if not bn:
                d[p + '.bias'] = (u,)
            idx += 1
            if bn:
                _bn(d, f'features.features.{idx}', u); idx += 1
            d[f'features.features.{idx}.weight'] = (1,); idx += 1
            nin = u
            if j < len(sizes) - 1:
                u = u * scaling
        latent = nin
    d['classifier.weight'] = (1, latent) + (1, 1)
    d['classifier.bias'] = (1,)
    return d


def unet_shapes(nf, base, top, dims):
    d = OrderedDict()
    def cv(name, co, ci, k):
        d[name + '.weight'] = (co, ci) + (k,) * dims
        d[name + '.bias'] = (co,)
    cv('enc1.0', nf, 1, base)
    for i in range(2, 7):
        cv(f'enc{i}.0', nf, nf, 3)
    cv('dec5.0', 2 * nf, 2 * nf, 3); cv('dec5.2', 2 * nf, 2 * nf, 3)
    for l in (4, 3, 2):
        cv(f'dec{l}.0', 2 * nf, 3 * nf, 3); cv(f'dec{l}.2', 2 * nf, 2 * nf, 3)
    cv('dec1.0', 64, 2 * nf + 1, top); cv('dec1.2', 32, 64, top); cv('dec1.4', 1, 32, top)
    return d
