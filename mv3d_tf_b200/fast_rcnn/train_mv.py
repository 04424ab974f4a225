"""Training entry points with the reference's names (lib/fast_rcnn/train_mv.py): SolverWrapper, train_net.

The reference builds the loss graph (train_mv.py:94-139) and lets TensorFlow differentiate it and apply
`tf.train.AdamOptimizer(1e-5)` (:144-146).  Here `SolverWrapper.train_step` plays the same step explicitly on the
current CUDA stream through the C ABI: the forward program of `MV3D_train`, the two loss kernels, a hand-written
backward pass (tcgen05 backward-data / backward-filter GEMMs, ROI-pool / max-pool / bias kernels), one optional NCCL
all-reduce of the flat gradient buffer (data-parallel training) and one fused Adam kernel over the flat
parameter buffer.  PyTorch tensors are containers; no torch math touches activations, gradients or weights.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import kernels as K
from .._lib import check, current_stream, lib, ptr
from ..networks.network import Node, Val

RPN_SIGMA = 3.0    # train_mv.py:113,131


def _node(net, name) -> Node:
    n = net.layers[name]
    return n[0] if isinstance(n, tuple) else n


class SolverWrapper(object):
    """train_mv.py:27-219.  `network` is an MV3D_train; `imdb` / `roidb` are accepted for signature compatibility
    (`roidb` may be any iterable of blob dicts, the output of RoIDataLayer.forward() in the reference)."""

    def __init__(self, sess=None, saver=None, network=None, imdb=None, roidb=None, output_dir=None,
                 pretrained_model=None, lr=0.00001, beta1=0.9, beta2=0.999, epsilon=1e-8, keep_prob=0.5,
                 process_group=None):
        self.net = network
        self.imdb = imdb
        self.roidb = roidb
        self.output_dir = output_dir
        self.pretrained_model = pretrained_model
        self.lr, self.beta1, self.beta2, self.epsilon = lr, beta1, beta2, epsilon
        self.keep_prob = keep_prob
        self.pg = process_group
        from ..sharding import FlatGradExchange
        self.exchange = FlatGradExchange(process_group) if process_group is not None else None
        self.step = 0
        net = self.net
        if process_group is not None:   # data-parallel ranks must not draw identical dropout masks
            import torch.distributed as dist
            net.dropout_seed = (dist.get_rank(process_group) + 1) << 32
        net.training = True
        if not net.params:
            net.init_weights()
        if pretrained_model is not None:
            net.load(pretrained_model, None, None, True)
        self._flatten_parameters()
        self._dpacked: Dict[str, K.PackedWeight] = {}
        net._on_weights_changed.append(self._dpacked.clear)   # Network.load on a live solver (resume)
        self.loss = torch.zeros(4, dtype=torch.float32, device=net.device)  # rpn_cls, rpn_box, cls, box
        self.last_grad_events = None
        self._side = None
        self._bias_streams = {}          # producing stream handle -> helper stream for that trunk's bias gradients
        if self.exchange is not None:
            self._set_exchange_regions()

    # ------------------------------------------------------------------ parameters
    def _flatten_parameters(self):
        """One flat fp32 buffer each for parameters, gradients and the two Adam moments (Adam and the gradient
        all-reduce are then single launches).  fc layers that follow roi_pool keep their rows in the kernel-native
        (H,W,C) order inside the buffer; `export_params` converts back to the reference's (C,H,W) row order."""
        net = self.net
        self.names = list(net.param_specs)
        total = 0
        self.slices = {}
        for name in self.names:
            shape = net.param_specs[name]['shape']
            nw, nb = int(np.prod(shape)), int(shape[-1])
            self.slices[name] = (total, nw, nb, shape)
            total += (nw + nb + 3) // 4 * 4
        dev = net.device
        self.theta = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.perm = {}
        for n in net._program:
            if n.kind == 'fc' and 'flatten_chw' in n.attrs:
                self.perm[n.name] = n.attrs['flatten_chw']
        self.gviews = {}
        for name in self.names:
            off, nw, nb, shape = self.slices[name]
            w = net.params[name]['weights']
            if name in self.perm and not net.native_fc_layout:
                cc, ph, pw = self.perm[name]
                w = w.view(cc, ph * pw, shape[-1]).permute(1, 0, 2).reshape(shape)
            self.theta[off:off + nw].view(shape).copy_(w)
            self.theta[off + nw:off + nw + nb].copy_(net.params[name]['biases'])
            net.params[name] = dict(weights=self.theta[off:off + nw].view(shape),
                                    biases=self.theta[off + nw:off + nw + nb])
            self.gviews[name] = dict(weights=self.grad[off:off + nw].view(shape),
                                     biases=self.grad[off + nw:off + nw + nb])
        net.native_fc_layout = True
        net._packed.clear()
        # parameters are declared trunks -> RPN -> fusion head, so the head (fc6_1 ... bbox_pred) is the buffer's tail
        self._head_off = self.slices['fc6_1'][0] if 'fc6_1' in self.slices else total
        assert all(self.slices[n][0] >= self._head_off for n in self.names if n.startswith(('fc', 'cls_score', 'bbox_pred')))
        assert all(self.slices[n][0] < self._head_off for n in self.names if n.startswith(('conv', 'rpn')))

    def _set_exchange_regions(self):
        """Gradient-buffer regions by producing stream: the side-stream trunk's parameters, and everything else."""
        net = self.net
        side_names = [n.name for n in net._program if n.kind == 'conv' and n.attrs.get('side') and n.name in self.slices]
        if not side_names or not net.use_side_stream:
            self.exchange.set_regions({})
            return
        lo = min(self.slices[n][0] for n in side_names)
        hi = max(self.slices[n][0] + (self.slices[n][1] + self.slices[n][2] + 3) // 4 * 4 for n in side_names)
        total = self.theta.numel()
        # [0, lo) main-stream trunk, [lo, hi) side-stream trunk, [hi, total) RPN + head (main stream, finished first)
        self.exchange.set_regions({'low': (0, lo), 'side': (lo, hi), 0: (hi, total)})

    def export_params(self):
        """{layer: {'weights', 'biases'}} numpy dict in the reference's variable layouts (the `.npy` format of
        network.py:45-64)."""
        out = {}
        for name in self.names:
            off, nw, nb, shape = self.slices[name]
            w = self.theta[off:off + nw].view(shape)
            if name in self.perm:
                cc, ph, pw = self.perm[name]
                w = w.view(ph * pw, cc, shape[-1]).permute(1, 0, 2).reshape(shape)
            out[name] = dict(weights=w.cpu().numpy().copy(), biases=self.theta[off + nw:off + nw + nb].cpu().numpy().copy())
        return out

    def export_grads(self):
        out = {}
        for name in self.names:
            off, nw, nb, shape = self.slices[name]
            w = self.grad[off:off + nw].view(shape)
            if name in self.perm:
                cc, ph, pw = self.perm[name]
                w = w.view(ph * pw, cc, shape[-1]).permute(1, 0, 2).reshape(shape)
            out[name] = dict(weights=w.cpu().numpy().copy(), biases=self.grad[off + nw:off + nw + nb].cpu().numpy().copy())
        return out

    def snapshot(self, sess=None, iter=0):
        """train_mv.py:49-65: write the weights (as the reference-format .npy dict instead of a TF checkpoint)."""
        if self.output_dir is None:
            return None
        os.makedirs(self.output_dir, exist_ok=True)
        from .config import cfg
        infix = ('_' + cfg.TRAIN.SNAPSHOT_INFIX if cfg.TRAIN.get('SNAPSHOT_INFIX', '') != '' else '')
        filename = os.path.join(self.output_dir, cfg.TRAIN.SNAPSHOT_PREFIX + infix + '_iter_{:d}'.format(iter + 1) + '.npy')
        np.save(filename, self.export_params(), allow_pickle=True)
        print('Wrote snapshot to: {:s}'.format(filename))
        return filename

    def _dweight(self, key, names) -> K.PackedWeight:
        """Backward-data operand of one layer (or of sibling heads concatenated along the output axis)."""
        pw = self._dpacked.get(key)
        if pw is None:
            ws = [self.net.params[n]['weights'] for n in names]
            w = ws[0] if len(ws) == 1 else torch.cat(ws, dim=-1).contiguous()
            pw = self._dpacked[key] = K.pack_weights_dgrad(w)
        return pw

    # ------------------------------------------------------------------ one training step
    def train_step(self, blobs, keep_prob=None, apply_update=True):
        """blobs: the reference's feed (train_mv.py:168-176): 'image_data', 'lidar_bv_data' (dense (B,H,W,C), or a
        kernels.PadAct from BevRasterizer.to_pad), 'im_info', 'gt_boxes_bv', 'gt_boxes_3d', 'gt_boxes_corners', 'calib'
        (GT entries: one array, or a list with one array per frame).  Returns the four loss values as a device
        tensor [rpn_loss_cls, rpn_loss_box, loss_cls, loss_box] (no host sync)."""
        net = self.net
        kp = self.keep_prob if keep_prob is None else keep_prob
        feed = {net.image_data: blobs['image_data'], net.lidar_bv_data: blobs['lidar_bv_data'],
                net.im_info: blobs['im_info'], net.keep_prob: kp, net.gt_boxes_bv: blobs['gt_boxes_bv'],
                net.gt_boxes_3d: blobs['gt_boxes_3d'], net.gt_boxes_corners: blobs['gt_boxes_corners'],
                net.calib: blobs['calib']}
        fetch = ['cls_score', 'bbox_pred', 'rpn_cls_score', 'rpn_bbox_pred', 'rpn_data', 'roi_data_3d']
        net.training = True
        bv = blobs['lidar_bv_data']
        if hasattr(bv, 'B'):
            B, Hb, Wb = bv.B, bv.H, bv.W
        else:
            B, Hb, Wb = int(bv.shape[0]), int(bv.shape[1]), int(bv.shape[2])
        # the RPN targets do not depend on the network: compute them (and take their host sync) off the critical path
        net.precompute_anchor_targets(B, Hb // 2 // 2 // 2, Wb // 2 // 2 // 2, 8, blobs['gt_boxes_bv'], blobs['gt_boxes_3d'],
                                      blobs['im_info'])
        net.run([_node(net, f) for f in fetch], feed)
        vals = net.last_vals
        self.grad.zero_()
        self.loss.zero_()
        self._backward(vals, kp)
        # the exchange step: one sum over the flat gradient buffer, issued as two NCCL calls so that the
        # head's 78 % of the bytes (fc6/fc7, finished first) travel while the trunks are still in backward
        grad_scale = self.exchange.finish(self.grad, self._head_off) if self.exchange is not None else 1.0
        if apply_update:
            self.step += 1
            check(lib().mv3d_adam(ptr(self.theta), ptr(self.grad), ptr(self.m), ptr(self.v), self.theta.numel(),
                                  self.lr, self.beta1, self.beta2, self.epsilon, self.step, grad_scale,
                                  current_stream()), 'mv3d_adam')
            net._packed.clear()
            self._dpacked.clear()
        return self.loss

    # ------------------------------------------------------------------ backward pass
    def _backward(self, vals, kp):
        net, precise = self.net, self.net.precise
        g = self.gviews
        stream = current_stream()
        L = lib()

        # ---------------- R-CNN head losses (train_mv.py:121-133) ----------------
        cls_v, box_v = vals[_node(net, 'cls_score')].dense, vals[_node(net, 'bbox_pred')].dense
        rd = vals[_node(net, 'roi_data_3d')].extra
        R, nb, B = cls_v.shape[0], box_v.shape[1], rd['B']
        gh = torch.empty((R, 64), dtype=K.BF16, device=net.device)
        gl = torch.empty_like(gh) if precise else None
        check(L.mv3d_rcnn_loss(ptr(cls_v), cls_v.stride(0), ptr(box_v), box_v.stride(0), ptr(rd['labels']),
                               ptr(rd['targets']), nb, ptr(rd['bv']), ptr(rd['frame_counts']), B, R, 64, RPN_SIGMA,
                               ptr(gh), ptr(gl), ptr(self.loss[2:]), stream), 'mv3d_rcnn_loss')
        # fused [cls_score | bbox_pred] head on the dropped-out concat (MV3D_train.py:175-182)
        xin = vals[_node(net, 'drop7')]                      # (R,4096) after both dropouts
        n_cls = net.param_specs['cls_score']['shape'][1]
        dwc = torch.empty((xin.hi.shape[1], n_cls + nb), dtype=torch.float32, device=net.device)
        K.linear_wgrad(xin.hi, xin.lo, gh, gl, dwc, precise=precise, accumulate=False)
        g['cls_score']['weights'].copy_(dwc[:, :n_cls])
        g['bbox_pred']['weights'].copy_(dwc[:, n_cls:])
        dbc = torch.zeros(n_cls + nb, dtype=torch.float32, device=net.device)
        K.bias_grad(gh, gl, n_cls + nb, dbc)
        g['cls_score']['biases'].copy_(dbc[:n_cls])
        g['bbox_pred']['biases'].copy_(dbc[n_cls:])
        scale2 = 1.0 / (kp * kp) if kp < 1.0 else 1.0       # drop7 (per branch) and drop7 (fused) between fc7 and here
        d7h, d7l, _ = K.linear(gh, gl, self._dweight('head', ['cls_score', 'bbox_pred']), relu=False, precise=precise,
                               out_bf16=True, out_f32=False, mask_hi=xin.hi, mask_scale=scale2, use_bias=False)
        half = xin.hi.shape[1] // 2
        scale1 = 1.0 / kp if kp < 1.0 else 1.0
        for bi, (suffix, drop6, pool) in enumerate((('_1', 'drop6', 'pool_5'), ('_2', 'drop6_2', 'pool_5_2'))):
            g7h = d7h[:, bi * half:(bi + 1) * half].contiguous()
            g7l = d7l[:, bi * half:(bi + 1) * half].contiguous() if d7l is not None else None
            x6 = vals[_node(net, drop6)]                     # fc6 output after dropout: input of fc7
            K.linear_wgrad(x6.hi, x6.lo, g7h, g7l, g['fc7' + suffix]['weights'], precise=precise, accumulate=False)
            K.bias_grad(g7h, g7l, half, g['fc7' + suffix]['biases'])
            g6h, g6l, _ = K.linear(g7h, g7l, self._dweight('fc7' + suffix, ['fc7' + suffix]), relu=False,
                                   precise=precise, out_bf16=True, out_f32=False, mask_hi=x6.hi, mask_scale=scale1,
                                   use_bias=False)
            xp = vals[_node(net, pool)]                      # pooled features (R, 7*7*512), rows in (H,W,C) order
            K.linear_wgrad(xp.hi, xp.lo, g6h, g6l, g['fc6' + suffix]['weights'], precise=precise, accumulate=False)
            K.bias_grad(g6h, g6l, g6h.shape[1], g['fc6' + suffix]['biases'])
            _, _, dpool = K.linear(g6h, g6l, self._dweight('fc6' + suffix, ['fc6' + suffix]), relu=False,
                                   precise=precise, out_bf16=False, out_f32=True, use_bias=False)
            e = xp.extra
            feat = e['feat']
            dfeat = torch.empty_like(feat)
            ph, pw, _sc = _node(net, pool).attrs['cfg']
            check(L.mv3d_roi_pool_backward(ptr(dpool), e['scale'], feat.shape[0], R, feat.shape[1], feat.shape[2],
                                           feat.shape[3], ph, pw, ptr(e['rois']), ptr(dfeat), ptr(e['argmax']),
                                           stream), 'mv3d_roi_pool_backward')
            vals[_node(net, pool)].extra['dfeat'] = dfeat

        if self.exchange is not None:   # head gradients are final: start their all-reduce under the trunk backward
            self.exchange.start_tail(self.grad, self._head_off)

        # ---------------- RPN losses (train_mv.py:94-119) ----------------
        rcls, rbox = vals[_node(net, 'rpn_cls_score')].dense, vals[_node(net, 'rpn_bbox_pred')].dense
        ad = vals[_node(net, 'rpn_data')].extra
        Bq, Hf, Wf = rcls.shape[0], rcls.shape[1], rcls.shape[2]
        A = ad['A']
        ghd_hi, ghd_lo = K._new_pad(Bq, Hf, Wf, 64, precise, net.device)
        check(L.mv3d_rpn_loss(ptr(rcls), ptr(rbox), ptr(ad['labels']), ptr(ad['targets']), ptr(ad['counts']), Bq, Hf,
                              Wf, A, 64, RPN_SIGMA, ptr(ghd_hi), ptr(ghd_lo), ptr(self.loss), stream), 'mv3d_rpn_loss')
        ghd = K.PadAct(ghd_hi, ghd_lo, Bq, Hf, Wf, 8 * A)
        rc = vals[_node(net, 'rpn_conv/3x3')].pad
        dwh = torch.zeros((1, rc.C, 8 * A), dtype=torch.float32, device=net.device)
        K.conv_wgrad(rc, ghd, dwh, precise=precise, accumulate=True)
        g['rpn_cls_score']['weights'].view(rc.C, 2 * A).copy_(dwh[0, :, :2 * A])
        g['rpn_bbox_pred']['weights'].view(rc.C, 6 * A).copy_(dwh[0, :, 2 * A:])
        dbh = torch.zeros(8 * A, dtype=torch.float32, device=net.device)
        K.bias_grad(ghd.hi, ghd.lo, 8 * A, dbh)
        g['rpn_cls_score']['biases'].copy_(dbh[:2 * A])
        g['rpn_bbox_pred']['biases'].copy_(dbh[2 * A:])
        grc, _ = K.conv(ghd, self._dweight('rpn_heads', ['rpn_cls_score', 'rpn_bbox_pred']), relu=False,
                        precise=precise, out_pad=True, mask=rc, use_bias=False)

        # ---------------- trunks, last layer first ----------------
        # The two trunks are independent below conv5: the RGB trunk's backward runs on a side stream next to the BEV
        # trunk's (forward does the same), and each hands its finished gradient slices to the exchange on its own stream.
        grads: Dict[Node, K.PadAct] = {_node(net, 'rpn_conv/3x3'): grc}
        dense_in: Dict[Node, torch.Tensor] = {_node(net, 'conv5_3'): vals[_node(net, 'pool_5')].extra['dfeat'],
                                              _node(net, 'conv5_3_2'): vals[_node(net, 'pool_5_2')].extra['dfeat']}
        main = torch.cuda.current_stream()
        side = None
        if net.use_side_stream:
            # the stream the forward pass ran this trunk on (its activations live in that stream's allocator pool)
            side = (net._side_stream or {}).get(1)
            if side is None:
                if self._side is None:
                    self._side = torch.cuda.Stream()
                side = self._side
            side.wait_stream(main)      # the ROI-path gradient of conv5_3_2 is ready
        for node in reversed(net._program):
            on_side = side is not None and bool(node.attrs.get('side'))
            with torch.cuda.stream(side if on_side else main):
                self._backward_node(node, vals, grads, dense_in, g, precise)
        if side is not None:
            main.wait_stream(side)
        for bs in self._bias_streams.values():   # every bias gradient landed before the exchange finishes / Adam reads
            main.wait_stream(bs)

    def _region_of(self, off):
        for key, (lo, hi) in self.exchange._regions.items():
            if lo <= off < hi:
                return key
        return 0

    def _backward_node(self, node, vals, grads, dense_in, g, precise):
        net = self.net
        if node.kind == 'max_pool':
            gp = grads.pop(node, None)
            if gp is None:
                return
            src = node.inputs[0]
            grads[src] = K.maxpool2x2_bwd(vals[src].pad, gp)   # routed + gated by the producing conv's ReLU
            return
        if node.kind != 'conv' or node.name in ('rpn_cls_score', 'rpn_bbox_pred'):
            return
        gn = grads.pop(node, None)
        if gn is None:
            d = dense_in.pop(node, None)
            if d is None:
                return
            gn = K.pad_nhwc_masked(d, vals[node].pad, precise=precise)   # only the ROI path feeds this layer
        # The bias gradient re-reads the whole gradient at the HBM rate while the filter-gradient GEMM of the same layer is
        # tensor-bound and leaves most of every SM's shared memory / registers free: it goes to a helper stream and runs
        # UNDER that GEMM instead of in front of it.
        cur = torch.cuda.current_stream()
        bs = self._bias_streams.get(cur.cuda_stream)
        if bs is None:   # MV3D_BIAS_STREAM=0: A/B switch (same stream); measured -0.4 ms per step with the helper stream
            bs = self._bias_streams[cur.cuda_stream] = cur if os.environ.get('MV3D_BIAS_STREAM', '1') == '0' \
                else torch.cuda.Stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        bs.wait_event(ready)
        with torch.cuda.stream(bs):
            K.bias_grad(gn.hi, gn.lo, node.channels, g[node.name]['biases'])
            bias_done = torch.cuda.Event()
            bias_done.record(bs)
        gn.hi.record_stream(bs)
        if gn.lo is not None:
            gn.lo.record_stream(bs)
        src = node.inputs[0]
        x = vals[src].pad
        if x is None and isinstance(vals[src].extra, K.PadAct):
            # first image layer: the forward kept the im2col rows of the input (Network.conv) -- one taps = 1 GEMM, K = 27
            dw = g[node.name]['weights']
            K.conv_wgrad(vals[src].extra, gn, dw.view(1, dw.shape[0] * dw.shape[1] * dw.shape[2], dw.shape[3]),
                         precise=precise, accumulate=True, tap_window=False)
        else:
            K.conv_wgrad(x, gn, g[node.name]['weights'], precise=precise, accumulate=True)
        if self.exchange is not None:   # this layer's slice (and everything above it in its region) is final
            cur.wait_event(bias_done)
            off = self.slices[node.name][0]
            self.exchange.ready(self.grad, off, key=self._region_of(off))
        if src.kind == 'placeholder':
            return
        gated = src.kind == 'conv'   # the dgrad epilogue applies the ReLU gate of the producing conv
        gsrc, _ = K.conv(gn, self._dweight(node.name, [node.name]), relu=False, precise=precise, out_pad=True,
                         mask=x if gated else None, addend=dense_in.pop(src, None), use_bias=False)
        grads[src] = gsrc

    # ------------------------------------------------------------------ reference-shaped loop
    def train_model(self, sess=None, max_iters=10000, data=None, display=None):
        """train_mv.py:87-219: loop over blobs, one Adam step each, periodic loss print + snapshot."""
        from .config import cfg
        data = self.roidb if data is None else data
        display = int(cfg.TRAIN.DISPLAY) if display is None else display    # train_mv.py:197
        if isinstance(data, list) and data and isinstance(data[0], dict) and 'lidar_bv_path' in data[0]:
            data = get_data_layer(data, self.imdb.num_classes if self.imdb is not None else 2)   # train_mv.py:154
        it = iter(data)
        last = None
        for i in range(max_iters):
            try:
                blobs = next(it)
            except StopIteration:
                it = iter(data)
                blobs = next(it)
            loss = self.train_step(blobs)
            if (i + 1) % display == 0:
                v = loss.tolist()
                print('iter: %d / %d, total loss: %.4f, rpn_loss_cls: %.4f, rpn_loss_box: %.4f, loss_cls: %.4f, '
                      'loss_box: %.4f, lr: %f' % (i + 1, max_iters, sum(v), v[0], v[1], v[2], v[3], self.lr))
            if (i + 1) % cfg.TRAIN.SNAPSHOT_ITERS == 0:
                last = i
                self.snapshot(None, i)
        if last != max_iters - 1 and max_iters > 0:
            self.snapshot(None, max_iters - 1)


def get_training_roidb(imdb):
    """train_mv.py:315-332: (optionally flipped) roidb enriched by prepare_roidb."""
    from .config import cfg
    from ..roi_data_layer import roidb as rdl_roidb
    if cfg.TRAIN.USE_FLIPPED:
        raise NotImplementedError('USE_FLIPPED: the MV3D minibatch never flips its blobs (minibatch_mv3d.py:32-40); '
                                  'the reference default is False (config.py:84)')
    print('Preparing training data...')
    rdl_roidb.prepare_roidb(imdb)
    print('done')
    return imdb.roidb


def get_data_layer(roidb, num_classes):
    """train_mv.py:335-346."""
    from ..roi_data_layer.layer import RoIDataLayer
    return RoIDataLayer(roidb, num_classes)


def filter_roidb(roidb):
    """train_mv.py:348-371: drop entries with neither a foreground nor a background RoI."""
    from .config import cfg

    def is_valid(entry):
        overlaps = entry['max_overlaps']
        fg_inds = np.where(overlaps >= cfg.TRAIN.FG_THRESH)[0]
        bg_inds = np.where((overlaps < cfg.TRAIN.BG_THRESH_HI) & (overlaps >= cfg.TRAIN.BG_THRESH_LO))[0]
        return len(fg_inds) > 0 or len(bg_inds) > 0
    num = len(roidb)
    filtered_roidb = [entry for entry in roidb if is_valid(entry)]
    print('Filtered {} roidb entries: {} -> {}'.format(num - len(filtered_roidb), num, len(filtered_roidb)))
    return filtered_roidb


def train_net(network, imdb, roidb, output_dir, pretrained_model=None, max_iters=10000):
    """train_mv.py:373-381."""
    if isinstance(roidb, list) and roidb and isinstance(roidb[0], dict) and 'max_overlaps' in roidb[0]:
        roidb = filter_roidb(roidb)
    sw = SolverWrapper(None, None, network, imdb, roidb, output_dir, pretrained_model=pretrained_model)
    print('Solving...')
    sw.train_model(None, max_iters)
    print('done solving')
    return sw
