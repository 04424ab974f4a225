#!/usr/bin/env python
"""Test an MV3D network -- the argument list of the reference's tools/test_net.py:23-56.  `--weights` is a `.npy` layer
dict (network.py:45-64 format, also what SolverWrapper.snapshot writes); detections are written in KITTI format by
imdb.evaluate_detections (kitti_mv3d.py:321-352)."""
import argparse
import os
import pprint
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Test a Fast R-CNN network')
    parser.add_argument('--device', dest='device', help='device to use', default='cpu', type=str)
    parser.add_argument('--device_id', dest='device_id', help='device id to use', default=0, type=int)
    parser.add_argument('--def', dest='prototxt', help='prototxt file defining the network', default=None, type=str)
    parser.add_argument('--weights', dest='model', help='model to test', default=None, type=str)
    parser.add_argument('--cfg', dest='cfg_file', help='optional config file', default=None, type=str)
    parser.add_argument('--wait', dest='wait', help='wait until net file exists', default=True, type=bool)
    parser.add_argument('--imdb', dest='imdb_name', help='dataset to test', default='voc_2007_test', type=str)
    parser.add_argument('--comp', dest='comp_mode', help='competition mode', action='store_true')
    parser.add_argument('--network', dest='network_name', help='name of the network', default=None, type=str)
    parser.add_argument('--kitti', dest='kitti_path', help='KITTI root (contains object/ and ImageSets/)', default=None)
    if argv is None and len(sys.argv) == 1:
        parser.print_help()
        sys.exit(1)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    print('Called with args:')
    print(args)
    from mv3d_tf_b200.datasets.factory import get_imdb
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_file, get_output_dir
    from mv3d_tf_b200.fast_rcnn.test_mv import test_net
    from mv3d_tf_b200.networks.factory import get_network

    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)
    print('Using config:')
    pprint.pprint(cfg)
    while not os.path.exists(args.model) and args.wait:
        print('Waiting for {} to exist...'.format(args.model))
        time.sleep(10)
    weights_filename = os.path.splitext(os.path.basename(args.model))[0]
    imdb = get_imdb(args.imdb_name, **({'kitti_path': args.kitti_path} if args.kitti_path else {}))
    imdb.competition_mode(args.comp_mode)
    # test_net.py:84-88: the device flag selects the NMS comparison rule (gpu: IoU > thresh, cpu: >=)
    cfg.USE_GPU_NMS = args.device == 'gpu'
    cfg.GPU_ID = args.device_id
    import torch
    torch.cuda.set_device(args.device_id)
    network = get_network(args.network_name)
    print('Use network `{:s}` in training'.format(args.network_name))
    network.load(args.model, None, None, True)
    print('Loading model weights from {:s}'.format(args.model))
    all_boxes, all_boxes_cnr = test_net(None, network, imdb, weights_filename)
    out = imdb.evaluate_detections(all_boxes, all_boxes_cnr, get_output_dir(imdb, weights_filename))
    print('Wrote KITTI results to {:s}'.format(out))
    return out


if __name__ == '__main__':
    main()
