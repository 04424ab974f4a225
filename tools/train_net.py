#!/usr/bin/env python
"""Train an MV3D network -- the argument list of the reference's tools/train_net.py:23-62 (mv3d.sh passes
--device gpu --device_id 0 --weights <npy> --imdb kitti_train --iters N --cfg experiments/cfgs/faster_rcnn_end2end.yml
--network MV3D_train).  `--solver` is accepted and ignored as in the reference; `--kitti` points at the KITTI root
(<ROOT>/data/KITTI by default)."""
import argparse
import os
import pprint
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Train a Fast R-CNN network')
    parser.add_argument('--device', dest='device', help='device to use', default='gpu', type=str)
    parser.add_argument('--device_id', dest='device_id', help='device id to use', default=0, type=int)
    parser.add_argument('--solver', dest='solver', help='solver prototxt', default=None, type=str)
    parser.add_argument('--iters', dest='max_iters', help='number of iterations to train', default=10000, type=int)
    parser.add_argument('--weights', dest='pretrained_model', help='initialize with pretrained model weights',
                        default=None, type=str)
    parser.add_argument('--cfg', dest='cfg_file', help='optional config file', default=None, type=str)
    parser.add_argument('--imdb', dest='imdb_name', help='dataset to train on', default='kitti_train', type=str)
    parser.add_argument('--rand', dest='randomize', help='randomize (do not use a fixed seed)', action='store_true')
    parser.add_argument('--network', dest='network_name', help='name of the network', default=None, type=str)
    parser.add_argument('--kitti', dest='kitti_path', help='KITTI root (contains object/ and ImageSets/)', default=None)
    parser.add_argument('--set', dest='set_cfgs', help='set config keys', default=None, nargs=argparse.REMAINDER)
    if argv is None and len(sys.argv) == 1:
        parser.print_help()
        sys.exit(1)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    print('Called with args:')
    print(args)
    from mv3d_tf_b200.datasets.factory import get_imdb
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_file, cfg_from_list, get_output_dir
    from mv3d_tf_b200.fast_rcnn.train_mv import get_training_roidb, train_net
    from mv3d_tf_b200.networks.factory import get_network

    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)
    if args.set_cfgs is not None:
        cfg_from_list(args.set_cfgs)
    print('Using config:')
    pprint.pprint(cfg)
    if not args.randomize:
        np.random.seed(cfg.RNG_SEED)      # tools/train_net.py:78-80
    if args.device == 'gpu':
        import torch
        torch.cuda.set_device(args.device_id)
    imdb = get_imdb(args.imdb_name, **({'kitti_path': args.kitti_path} if args.kitti_path else {}))
    print('Loaded dataset `{:s}` for training'.format(imdb.name))
    roidb = get_training_roidb(imdb)
    output_dir = get_output_dir(imdb, None)
    print('Output will be saved to `{:s}`'.format(output_dir))
    network = get_network(args.network_name)
    print('Use network `{:s}` in training'.format(args.network_name))
    return train_net(network, imdb, roidb, output_dir, pretrained_model=args.pretrained_model, max_iters=args.max_iters)


if __name__ == '__main__':
    main()
