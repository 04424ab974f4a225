"""get_network(name) as lib/networks/factory.py:23-33."""
from .MV3D_test import MV3D_test


def get_network(name, **kw):
    split = name.split('_')[1]
    if split == 'test':
        return MV3D_test(**kw)
    if split == 'train':
        from .MV3D_train import MV3D_train  # noqa: WPS433
        return MV3D_train(**kw)
    raise KeyError('Unknown dataset: {}'.format(name))
