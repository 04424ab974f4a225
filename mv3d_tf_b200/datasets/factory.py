"""get_imdb(name) / list_imdbs() as lib/datasets/factory.py:52-58,95-105 (only the KITTI-MV3D sets exist here)."""
from .kitti_mv3d import kitti_mv3d

__sets = {}
for _split in ('train', 'val', 'trainval', 'test'):
    __sets['kitti_{}'.format(_split)] = (lambda split=_split, **kw: kitti_mv3d(split, **kw))


def get_imdb(name, **kw):
    """Get an imdb (image database) by name.  `kitti_path=` overrides <ROOT_DIR>/data/KITTI."""
    if name not in __sets:
        raise KeyError('Unknown dataset: {}'.format(name))
    return __sets[name](**kw)


def list_imdbs():
    return list(__sets.keys())
