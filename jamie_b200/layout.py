"""Parameter / buffer naming of ``edModelVar`` (reference: jamie/model.py:147-220).

``param_spec`` is the ``named_parameters()`` registration order, which is also the packed order of the C ABI
(``jb_set_params`` / ``jb_get_params``); ``bn_spec`` is the module order of the 8 BatchNorm1d layers.
"""


def param_spec(dims, L):
    spec = [('sigma', (2,))]
    for i, D in enumerate(dims):
        spec += [(f'encoders.{i}.0.weight', (2 * D, D)), (f'encoders.{i}.0.bias', (2 * D,)),
                 (f'encoders.{i}.1.weight', (2 * D,)), (f'encoders.{i}.1.bias', (2 * D,)),
                 (f'encoders.{i}.4.weight', (D, 2 * D)), (f'encoders.{i}.4.bias', (D,)),
                 (f'encoders.{i}.5.weight', (D,)), (f'encoders.{i}.5.bias', (D,))]
    for i, D in enumerate(dims):
        spec += [(f'fc_mus.{i}.weight', (L, D)), (f'fc_mus.{i}.bias', (L,))]
    for i, D in enumerate(dims):
        spec += [(f'fc_vars.{i}.weight', (L, D)), (f'fc_vars.{i}.bias', (L,))]
    for i, D in enumerate(dims):
        spec += [(f'decoders.{i}.0.weight', (D, L)), (f'decoders.{i}.0.bias', (D,)),
                 (f'decoders.{i}.1.weight', (D,)), (f'decoders.{i}.1.bias', (D,)),
                 (f'decoders.{i}.4.weight', (2 * D, D)), (f'decoders.{i}.4.bias', (2 * D,)),
                 (f'decoders.{i}.5.weight', (2 * D,)), (f'decoders.{i}.5.bias', (2 * D,)),
                 (f'decoders.{i}.8.weight', (D, 2 * D)), (f'decoders.{i}.8.bias', (D,))]
    return spec


def bn_spec(dims):
    out = []
    for i, D in enumerate(dims):
        out += [(f'encoders.{i}.1', 2 * D), (f'encoders.{i}.5', D)]
    for i, D in enumerate(dims):
        out += [(f'decoders.{i}.1', D), (f'decoders.{i}.5', 2 * D)]
    return out
