"""Graph builder with the reference's layer-method signatures (lib/networks/network.py:14-409).

The reference builds a TF1 graph with `self.feed(...).conv(...).max_pool(...)` chains and runs it with
`sess.run(fetches, feed_dict)`.  This class keeps exactly that surface -- `feed`, `get_output`, the
`@layer` methods with the same argument lists, `load` of the `.npy` weight dict -- but records a small
program of kernel launches; `run(fetches, feed_dict)` plays it on the current CUDA stream through the
C ABI (mv3d_tf_b200.kernels).  Nothing here computes on the host or with torch math.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import ctypes as C
import os

import numpy as np
import torch

from .. import kernels as K
from .._lib import (ROI_BEV, ROI_FV, ROI_GIVEN, ROI_IMG, RoiProjection, RoiView, check, current_stream, lib, ptr)
from ..fast_rcnn.config import cfg
from ..rpn_msr.anchor_target_layer_tf import AnchorTargetLayer
from ..rpn_msr.proposal_layer_tf import ProposalLayer3D
from ..rpn_msr.proposal_target_layer_tf import ProposalTargetLayer3D
from ..utils.transform import REF_GEOMETRY, BevGeometry

DEFAULT_PADDING = 'SAME'


@dataclass
class Node:
    """A symbolic layer output (the analogue of a tf.Tensor handle)."""
    name: str
    kind: str
    inputs: List["Node"] = field(default_factory=list)
    fn: Optional[Callable] = None
    channels: int = 0            # feature channels (maps) or vector width (fc)
    pooled: Optional[tuple] = None  # (PH, PW) after roi_pool
    consumers: List[str] = field(default_factory=list)
    consumer_nodes: List["Node"] = field(default_factory=list)
    attrs: Dict[str, Any] = field(default_factory=dict)

    def __hash__(self):
        return id(self)


@dataclass
class Val:
    """Runtime value of a node: any subset of the representations below."""
    pad: Optional[K.PadAct] = None          # conv activations (PAD layout, bf16 hi/lo)
    dense: Optional[torch.Tensor] = None    # float32 NHWC / generic float32 tensor
    hi: Optional[torch.Tensor] = None       # (rows, width) bf16 pair feeding fc
    lo: Optional[torch.Tensor] = None
    extra: Any = None


def layer(op):
    def layer_decorated(self, *args, **kwargs):
        name = kwargs.setdefault('name', self.get_unique_name(op.__name__))
        if len(self.inputs) == 0:
            raise RuntimeError('No input variables found for layer %s.' % name)
        elif len(self.inputs) == 1:
            layer_input = self.inputs[0]
        else:
            layer_input = list(self.inputs)
        layer_output = op(self, layer_input, *args, **kwargs)
        self.layers[name] = layer_output
        self.feed(layer_output)
        return self
    return layer_decorated


class Network(object):
    def __init__(self, inputs, trainable=True, precise=True, geometry: BevGeometry = REF_GEOMETRY,
                 img_size=(375, 1242), device='cuda', fv_geometry=None, mixed=False):
        self.inputs = []
        self.layers = dict(inputs)
        self.trainable = trainable
        self.precise = precise          # True: 3-pass bf16 hi/lo GEMMs (parity mode); False: single pass
        # mixed (inference, with precise): activations between 3x3 convs travel as fp16 + e5m2 pairs (FMT_F16E5) and those
        # convs run one fp16 pass + one e5m2 correction pass (2/3 of the 3-pass tensor time, ~1e-4 at conv5_3 instead
        # of ~2e-5; the contract is 1e-3).  1x1 convs, fc layers and training keep the bf16 hi/lo 3-pass form.
        self.mixed = bool(mixed) and bool(precise)
        self.geometry = geometry
        self.img_size = img_size
        self.fv_geometry = fv_geometry   # None: two-view network exactly as the reference (network.py:313-315)
        self.device = torch.device(device)
        self.params: Dict[str, Dict[str, torch.Tensor]] = {}
        self.param_specs: Dict[str, dict] = {}
        self._packed: Dict[str, K.PackedWeight] = {}
        self._on_weights_changed: List[Callable] = []   # e.g. the solver's backward-data operand cache
        self._program: List[Node] = []
        self._proposal_layers: Dict[tuple, ProposalLayer3D] = {}
        self._anchor_target_layers: Dict[tuple, AnchorTargetLayer] = {}
        self._proposal_target_layer: Optional[ProposalTargetLayer3D] = None
        self._roi_nodes: List[Node] = []
        self._needed_now = set()
        self._anchor_pre = None
        self._aux_stream = None
        self.roi_from_pad = os.environ.get('MV3D_ROI_FROM_PAD', '1') != '0'   # fused ROI pool reads conv5's PAD planes
        self.fc_f16e5 = os.environ.get('MV3D_FC_F16E5', '1') != '0'   # mixed mode: ROI pool -> fc6 in the 2-pass f16e5 format
        self.node_events = None         # list -> Network.run appends (name, kind, start event, end event) per node
        self.last_num_rois = None
        self.training = False           # True: roi_pool keeps argmax, dropout draws masks (set by the solver)
        self.native_fc_layout = False   # True: fc-after-roi_pool weights are stored with rows already in (H,W,C) order
        self.dropout_seed = 0
        self.last_vals = None
        self.use_side_stream = True     # independent branches marked 'side' (the RGB trunk) run on a second stream
        self._side_stream = None
        self.setup()

    def setup(self):
        raise NotImplementedError('Must be subclassed.')

    # ------------------------------------------------------------------ graph plumbing
    def feed(self, *args):
        assert len(args) != 0
        self.inputs = []
        for lyr in args:
            if isinstance(lyr, str):
                try:
                    lyr = self.layers[lyr]
                except KeyError:
                    raise KeyError('Unknown layer name fed: %s' % lyr)
            self.inputs.append(lyr)
        return self

    def get_output(self, layer):
        try:
            return self.layers[layer]
        except KeyError:
            raise KeyError('Unknown layer name fed: %s' % layer)

    def get_unique_name(self, prefix):
        idx = sum(t.startswith(prefix) for t, _ in self.layers.items()) + 1
        return '%s_%d' % (prefix, idx)

    def validate_padding(self, padding):
        assert padding in ('SAME', 'VALID')

    def _node(self, name, kind, inputs, fn, **kw):
        ins = [i[0] if isinstance(i, tuple) else i for i in inputs]
        n = Node(name=name, kind=kind, inputs=ins, fn=fn, **kw)
        for i in ins:
            if isinstance(i, Node):
                i.consumers.append(kind)
                i.consumer_nodes.append(n)
        self._program.append(n)
        return n

    @staticmethod
    def placeholder(name, channels=0):
        return Node(name=name, kind='placeholder', channels=channels)

    # ------------------------------------------------------------------ parameters
    def _declare(self, name, shape, stddev):
        self.param_specs[name] = dict(shape=tuple(int(s) for s in shape), stddev=stddev)

    def init_weights(self, seed=7, mode='reference'):
        """Random initialisation.  'reference': truncated_normal(0, 0.01) (0.001 for bbox_pred), zero biases
        (network.py:117-118,385-390).  'he': fan-in scaled so activations survive 13 layers (synthetic benchmarks)."""
        g = torch.Generator(device='cpu').manual_seed(seed)
        for name, spec in self.param_specs.items():
            shape = spec['shape']
            w = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, 0.0, 1.0, -2.0, 2.0, generator=g)
            if mode == 'he':
                fan_in = int(np.prod(shape[:-1]))
                w *= (2.0 / fan_in) ** 0.5
                b = torch.empty(shape[-1]).uniform_(-0.05, 0.05, generator=g)
            else:
                w *= spec['stddev']
                b = torch.zeros(shape[-1])
            self.params[name] = dict(weights=w.to(self.device), biases=b.to(self.device))
        self._packed.clear()

    def load(self, data_path, session=None, saver=None, ignore_missing=False):
        """`.npy` dict {layer: {'weights': HWIO / (in,out), 'biases'}} (network.py:45-64).  session/saver are
        accepted for signature compatibility and unused.
        A parameter that already exists is overwritten IN PLACE: under a live SolverWrapper the parameters are views
        into its flat buffer (Adam updates that buffer), and fc-after-roi_pool weights live there with their rows in the
        kernel-native (H,W,C) order (`native_fc_layout`), so the file's (C,H,W) rows are permuted on the way in."""
        data_dict = np.load(data_path, allow_pickle=True, encoding='latin1').item()
        chw = {n.name: n.attrs['flatten_chw'] for n in self._program if n.kind == 'fc' and 'flatten_chw' in n.attrs}
        for key in data_dict:
            if key not in self.param_specs:
                if not ignore_missing:
                    raise ValueError('no layer named %s' % key)
                continue
            tgt = self.params.setdefault(key, {})
            for subkey, arr in data_dict[key].items():
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)
                shape = self.param_specs[key]['shape']
                want = shape if subkey == 'weights' else (shape[-1],)
                if tuple(t.shape) != tuple(want):
                    if not ignore_missing:
                        raise ValueError('shape mismatch for %s/%s' % (key, subkey))
                    continue
                if subkey == 'weights' and self.native_fc_layout and key in chw:
                    cc, ph, pw = chw[key]
                    t = t.view(cc, ph * pw, shape[-1]).permute(1, 0, 2).reshape(shape)
                if subkey in tgt and tgt[subkey].shape == t.shape:
                    tgt[subkey].copy_(t)
                else:
                    tgt[subkey] = t.contiguous()
        self._packed.clear()
        for hook in self._on_weights_changed:
            hook()

    def _weight(self, name, transform=None, fmt=K.FMT_BF16X2) -> K.PackedWeight:
        key = name if fmt == K.FMT_BF16X2 else name + '/f16e5'
        pw = self._packed.get(key)
        if pw is None:
            p = self.params[name]
            w = p['weights']
            if transform is not None and not self.native_fc_layout:
                w = transform(w)
            pw = self._packed[key] = K.pack_weights(w, p['biases'], fmt=fmt)
        return pw

    def _pad_out_fmt(self, node) -> int:
        """Operand format of a conv node's PAD output: FMT_F16E5 when every reader of that PAD tensor is a 3x3 conv
        (directly or through 2x2 pools), else bf16 hi/lo."""
        if not self.mixed or self.training or node.channels % 64 != 0:
            return K.FMT_BF16X2
        fmt = node.attrs.get('pad_out_fmt')
        if fmt is None:
            def accepts(c):
                if c.kind == 'conv':
                    return c.attrs.get('k') == (3, 3)
                if c.kind == 'max_pool':
                    return all(accepts(cc) for cc in c.consumer_nodes if cc.kind in ('conv', 'max_pool'))
                return True
            # the ROI pool reads either rendering, so it does not vote: a map only it reads stays bf16 hi/lo (2^-17)
            readers = [c for c in node.consumer_nodes if c.kind in ('conv', 'max_pool')]
            fmt = K.FMT_F16E5 if readers and all(accepts(c) for c in readers) else K.FMT_BF16X2
            node.attrs['pad_out_fmt'] = fmt
        return fmt

    # ------------------------------------------------------------------ layers (reference signatures)
    @layer
    def conv(self, input, k_h, k_w, c_o, s_h, s_w, name, relu=True, padding=DEFAULT_PADDING, group=1, trainable=True):
        self.validate_padding(padding)
        assert group == 1 and s_h == 1 and s_w == 1, 'the MV3D nets only use stride-1 ungrouped convs'
        assert (k_h, k_w) in ((3, 3), (1, 1))
        assert (k_h, k_w) == (1, 1) or padding == 'SAME'
        c_i = input.channels
        self._declare(name, (k_h, k_w, c_i, c_o), 0.01)

        def run(vals, node):
            v = vals[node.inputs[0]]
            # the fused ROI pool reads the PAD operand planes directly (inference): no dense float32 copy of conv5_3
            pad_readers = ('conv', 'max_pool') if (self.training or not self.roi_from_pad) else ('conv', 'max_pool', 'roi_pool')
            want_pad = any(c in pad_readers for c in node.consumers) or self.training
            want_dense = (not want_pad) or any(c not in pad_readers for c in node.consumers) \
                or node.attrs.get('fetched', False)
            unpad_for_fetch = False
            if want_dense and 'roi_pool' in pad_readers and 'roi_pool' in node.consumers and (k_h, k_w) == (3, 3) \
                    and all(c in pad_readers for c in node.consumers):
                # fetched only: return the operand rendering the pool reads (unpad), not a second float32 epilogue output
                want_dense, unpad_for_fetch = False, True
            if (k_h, k_w) == (3, 3) and c_i <= 4 and c_o % 8 == 0 and v.dense is not None and v.pad is None \
                    and not self.training and want_pad and not want_dense:
                # tiny-channel first layer (RGB / front view, K = 27): one direct kernel straight into the consumer's
                # operand format instead of im2col + a K = 32 GEMM whose time is all epilogue.  (Training keeps the
                # nine-tap GEMM form: its backward-filter expects the PAD input.)
                p = self.params[name]
                return Val(pad=K.conv3x3_small_cin(v.dense, p['weights'].contiguous(), p['biases'], relu=relu,
                                                   precise=self.precise, out_fmt=self._pad_out_fmt(node)))
            if (k_h, k_w) == (3, 3) and c_i <= 3 and c_o == 64 and v.dense is not None and v.pad is None \
                    and self.training and not want_dense:
                # training: the same tcgen05 first-layer kernel forward (bf16 hi/lo out); the backward-filter pass of this
                # layer is ONE taps = 1 GEMM over the im2col rows kept here (K = 27 of 32) instead of nine N = 16 taps
                p = self.params[name]
                if not isinstance(v.extra, K.PadAct):
                    v.extra = K.im2col3x3(v.dense, precise=self.precise)
                return Val(pad=K.conv3x3_small_cin(v.dense, p['weights'].contiguous(), p['biases'], relu=relu,
                                                   precise=self.precise, out_fmt=K.FMT_BF16X2))
            if (k_h, k_w) == (3, 3) and 9 * c_i <= 32 and v.dense is not None and v.pad is None and not self.training:
                # same layer when its dense output is wanted too: im2col once, then ONE K=32 GEMM instead of nine taps
                # of 13/16 zero padding
                col = v.extra if isinstance(v.extra, K.PadAct) else K.im2col3x3(v.dense, precise=self.precise)
                v.extra = col
                pw = self._packed.get(name + '/im2col')
                if pw is None:
                    p = self.params[name]
                    pw = self._packed[name + '/im2col'] = K.pack_weights(p['weights'].reshape(1, 1, 9 * c_i, c_o),
                                                                         p['biases'], cin_pad=32)
                out, dense = K.conv(col, pw, relu=relu, precise=self.precise, out_pad=want_pad, out_f32_dense=want_dense,
                                    out_fmt=self._pad_out_fmt(node))
                return Val(pad=out, dense=dense)
            if v.pad is None:
                v.pad = K.pad_nhwc(v.dense, precise=self.precise)
            cached = node.attrs.pop('result', None)
            if cached is not None:
                return cached
            followers = [] if self.training else node.attrs.get('fused_followers', [])
            if followers and not want_pad:
                # sibling 1x1 heads on one input (rpn_cls_score | rpn_bbox_pred, MV3D_test.py:72-75): ONE GEMM over the
                # concatenated output channels; when the leader only feeds reshape -> softmax -> reshape (:76-81) the
                # pair softmax runs in that GEMM's epilogue and the three glue nodes pass the result through
                key = name + '+' + '+'.join(f.name for f in followers)
                pw = self._packed.get(key)
                if pw is None:
                    ws = [self.params[name]] + [self.params[f.name] for f in followers]
                    pw = self._packed[key] = K.pack_weights(torch.cat([q['weights'] for q in ws], dim=3).contiguous(),
                                                            torch.cat([q['biases'] for q in ws], dim=0).contiguous())
                fold = self._softmax_chain(node) is not None
                _, dense = K.conv(v.pad, pw, relu=False, precise=self.precise, out_pad=False, out_f32_dense=True,
                                  softmax_cols=c_o if fold else 0)
                off = c_o
                for f in followers:
                    f.attrs['result'] = Val(dense=dense[..., off:off + f.channels])
                    off += f.channels
                return Val(dense=dense[..., :c_o], extra=dict(softmax_folded=True) if fold else None)
            pw = self._weight(name, fmt=v.pad.fmt)
            if (k_h, k_w) == (3, 3) and self.precise and not self.training and want_pad and not want_dense \
                    and [c.kind for c in node.consumer_nodes] == ['max_pool'] and K.pool_fusable(v.pad, pw):
                # conv + Network.max_pool as one kernel: the un-pooled activation (4x the bytes) is never written
                out, _ = K.conv(v.pad, pw, relu=relu, precise=self.precise, out_fmt=self._pad_out_fmt(node), pool=True)
                return Val(pad=out, extra=dict(pooled=True))
            out, dense = K.conv(v.pad, pw, relu=relu, precise=self.precise,
                                out_pad=want_pad, out_f32_dense=want_dense, out_fmt=self._pad_out_fmt(node))
            if unpad_for_fetch:
                dense = K.unpad_nhwc(out)
            return Val(pad=out, dense=dense)
        n = self._node(name, 'conv', [input], run, channels=c_o)
        n.attrs['k'] = (k_h, k_w)
        n.attrs['relu'] = relu
        if (k_h, k_w) == (1, 1) and not relu:   # fusion planning: an earlier linear 1x1 head on the same input
            for m in self._program:
                if m is not n and m.kind == 'conv' and m.attrs.get('k') == (1, 1) and not m.attrs.get('relu', True) \
                        and m.inputs[0] is input and 'fused_into' not in m.attrs:
                    m.attrs.setdefault('fused_followers', []).append(n)
                    n.attrs['fused_into'] = m
                    break
        return n

    @staticmethod
    def _softmax_chain(node):
        """[reshape(2), softmax, reshape(C)] when `node` (a linear 1x1 conv) feeds exactly that chain and none of the
        intermediate tensors is fetched; else None."""
        chain, cur = [], node
        for kind in ('reshape', 'softmax', 'reshape'):
            if cur.attrs.get('fetched', False) or len(cur.consumer_nodes) != 1 or cur.consumer_nodes[0].kind != kind:
                return None
            cur = cur.consumer_nodes[0]
            chain.append(cur)
        if chain[0].channels != 2 or chain[2].channels != node.channels:
            return None
        return chain

    @layer
    def max_pool(self, input, k_h, k_w, s_h, s_w, name, padding=DEFAULT_PADDING):
        self.validate_padding(padding)
        assert (k_h, k_w, s_h, s_w) == (2, 2, 2, 2) and padding == 'VALID'

        def run(vals, node):
            v = vals[node.inputs[0]]
            if isinstance(v.extra, dict) and v.extra.get('pooled'):
                return Val(pad=v.pad)   # taken in the producing conv's epilogue
            return Val(pad=K.maxpool2x2(v.pad))
        return self._node(name, 'max_pool', [input], run, channels=input.channels)

    @layer
    def reshape_layer(self, input, d, name):
        def run(vals, node):
            v = vals[node.inputs[0]]
            if isinstance(v.extra, dict) and v.extra.get('softmax_folded'):
                return v   # the fused head already holds the probabilities in the final (B,H,W,2A) arrangement
            x = v.dense
            return Val(dense=x.reshape(x.shape[0], x.shape[1], -1, int(d)))
        return self._node(name, 'reshape', [input], run, channels=int(d))

    @layer
    def softmax(self, input, name):
        def run(vals, node):
            v = vals[node.inputs[0]]
            if isinstance(v.extra, dict) and v.extra.get('softmax_folded'):
                return v   # computed in the producing GEMM's epilogue
            x = v.dense
            assert x.shape[-1] == 2, 'MV3D only ever takes 2-way softmaxes'
            return Val(dense=K.softmax_pairs(x, 1))
        return self._node(name, 'softmax', [input], run, channels=input.channels)

    @layer
    def proposal_layer_3d(self, input, _feat_stride, cfg_key, name):
        def run(vals, node):
            prob = vals[node.inputs[0]].dense
            deltas = vals[node.inputs[1]].dense
            im_info = np.asarray(vals[node.inputs[2]].extra, dtype=np.float32).reshape(-1, 3)
            calib = vals[node.inputs[3]].extra
            if not isinstance(calib, torch.Tensor):
                calib = np.asarray(calib, dtype=np.float32)
            B, H, W = prob.shape[0], prob.shape[1], prob.shape[2]
            info = tuple(float(x) for x in im_info[0])
            key = (H, W, cfg_key, int(_feat_stride), info, cfg[cfg_key].RPN_PRE_NMS_TOP_N,
                   cfg[cfg_key].RPN_POST_NMS_TOP_N, bool(cfg.USE_GPU_NMS))
            pl = self._proposal_layers.get(key)
            if pl is None:
                pl = self._proposal_layers[key] = ProposalLayer3D(H, W, cfg_key, int(_feat_stride), info,
                                                                  geom=self.geometry, img_size=self.img_size,
                                                                  device=self.device)
            outs = []
            for b in range(B):  # the reference asserts B == 1; frames of a batch are independent
                if isinstance(calib, torch.Tensor):
                    cb = calib.view(-1, 12)[b if calib.numel() > 12 else 0]
                else:
                    cb = calib.reshape(-1, 4, 12)[b if calib.size > 48 else 0]
                outs.append(pl(prob[b], deltas[b], cb, batch_index=float(b)))
            if B == 1:
                o = outs[0]
                bv, img, p3d, num = o['bv'], o['img'], o['p3d'], o['num']
            else:
                bv = torch.cat([o['bv'] for o in outs]); img = torch.cat([o['img'] for o in outs])
                p3d = torch.cat([o['p3d'] for o in outs]); num = torch.cat([o['num'] for o in outs])
            self.last_num_rois = num
            return Val(extra=dict(bv=bv, img=img, p3d=p3d, num=num, per_frame=pl.capacity, outs=outs,
                                  layer=pl, calib=cb if B == 1 else None))
        n = self._node(name, 'proposal', list(input), run)
        # the reference returns the 4-tuple (rois_bv, rois_img, rois_3d, rois_3d)  (network.py:234)
        return (n, n, n, n)

    @layer
    def proposal_transform(self, input, name, target='bv'):
        assert target in ('bv', 'img', 'fv')
        if target == 'fv' and self.fv_geometry is None:
            return None  # as the reference (network.py:313-315)
        src = input[0] if isinstance(input, (tuple, list)) else input

        def run(vals, node):
            e = vals[node.inputs[0]].extra
            if target == 'fv' and 'fv' not in e:   # the 3-D proposals projected into the front-view map
                p3d = e['p3d'].contiguous()
                fv = torch.empty((p3d.shape[0], 5), dtype=torch.float32, device=p3d.device)
                e['fv'] = fv
                pools = [c for c in node.consumer_nodes if c.kind == 'roi_pool' and c in self._needed_now]
                if pools and self._fusable(e):
                    e['fv_pending'] = True         # written by the fused ROI-pool launch (d_rois_out), no extra launch
                else:
                    num = e['num'] if (e.get('num') is not None and e['num'].numel() == 1) else None
                    H, W, t0, dt, p1, dp = self.fv_geometry.c_args()
                    check(lib().mv3d_rois_to_fv(ptr(p3d), p3d.shape[0], ptr(num), H, W, t0, dt, p1, dp, ptr(fv),
                                                current_stream()), 'mv3d_rois_to_fv')
            return Val(dense=e[target], extra=e)
        n = self._node(name, 'rois', [src], run)
        n.attrs['target'] = target
        return n

    def _fusable(self, e) -> bool:
        """In-kernel projection of the 3-D proposals (mv3d_roi_pool_fused) needs the proposal layer's own constants and
        ONE projection matrix: single-frame inference."""
        return (not self.training) and isinstance(e, dict) and e.get('layer') is not None and e.get('calib') is not None \
            and e.get('p3d') is not None and e.get('num') is not None and e['num'].numel() == 1

    @layer
    def roi_pool(self, input, pooled_height, pooled_width, spatial_scale, name):
        data, rois = input[0], input[1]
        if isinstance(data, tuple):
            data = data[0]
        if isinstance(rois, tuple):
            rois = rois[0]

        def run(vals, node):
            cached = node.attrs.pop('result', None)
            if cached is not None:
                return cached
            # fuse every roi_pool node whose inputs are ready into one multi-view launch
            group = [m for m in self._roi_nodes if all(i in vals for i in m.inputs) and 'result' not in m.attrs]
            views = (RoiView * len(group))()
            results = []
            R = None
            for k, m in enumerate(group):
                fv_ = vals[m.inputs[0]]
                from_pad = (not self.training) and self.roi_from_pad and fv_.pad is not None and fv_.pad.lo is not None
                feat = None if from_pad else fv_.dense
                r = vals[m.inputs[1]].dense.contiguous()
                R = r.shape[0]
                ph, pw, sc = m.attrs['cfg']
                Cc = fv_.pad.C if from_pad else feat.shape[-1]
                fdev = fv_.pad.hi.device if from_pad else feat.device
                fH, fW = (fv_.pad.H, fv_.pad.W) if from_pad else (feat.shape[1], feat.shape[2])
                # mixed-mode inference: the pooled rows leave as f16e5 operands when every reader is a wide fc over <= 512
                # ROIs (the swapped split-K fc kernel runs them in 2 pass-equivalents instead of 3); same 16-bit containers
                top_fmt = K.FMT_F16E5 if (self.mixed and self.fc_f16e5 and not self.training and Cc % 64 == 0
                                          and 64 <= R <= 512 and m.consumer_nodes
                                          and all(c.kind == 'fc' and c.channels % 256 == 0 for c in m.consumer_nodes)) \
                    else K.FMT_BF16X2
                hi = torch.empty((R, ph * pw * Cc), dtype=torch.bfloat16, device=fdev)
                lo = torch.empty_like(hi) if (self.precise or top_fmt == K.FMT_F16E5) else None
                top = None
                if m.attrs.get('fetched', False):
                    top = torch.empty((R, ph, pw, Cc), dtype=torch.float32, device=fdev)
                arg = torch.empty((R, ph, pw, Cc), dtype=torch.int32, device=fdev) if self.training else None
                v = views[k]
                v.d_data, v.d_rois, v.height, v.width = ptr(feat), ptr(r), fH, fW
                if from_pad:
                    v.d_pad_hi, v.d_pad_lo, v.pad_fmt, v.pad_c = ptr(fv_.pad.hi), ptr(fv_.pad.lo), fv_.pad.fmt, fv_.pad.c_pad
                v.spatial_scale, v.d_top, v.d_argmax, v.d_top_hi, v.d_top_lo = sc, ptr(top), ptr(arg), ptr(hi), ptr(lo)
                v.top_fmt = top_fmt
                results.append(Val(dense=top, hi=hi, lo=lo, extra=dict(feat=feat, rois=r, argmax=arg, scale=sc, fmt=top_fmt)))
            ph, pw, _ = node.attrs['cfg']
            e = vals[node.inputs[1]].extra
            num = e['num'] if (isinstance(e, dict) and e.get('num') is not None and e['num'].numel() == 1) else None
            targets = [m.inputs[1].attrs.get('target') if isinstance(m.inputs[1], Node) else None for m in group]
            if self._fusable(e) and all(t in ('bv', 'img', 'fv') for t in targets) and group[0].channels % 8 == 0 \
                    and all(vals[m.inputs[1]].extra is e for m in group):
                # north-star (iv): one launch projects every 3-D proposal into its view's plane and pools all views
                pl, calib = e['layer'], e['calib']
                pp = pl.params
                proj = RoiProjection()
                proj.xn, proj.yn, proj.x_min, proj.y_min, proj.res = pp.xn, pp.yn, pp.x_min, pp.y_min, pp.res
                proj.im_h, proj.im_w = pp.im_h, pp.im_w
                if isinstance(calib, torch.Tensor):
                    proj.d_proj = calib.data_ptr()
                else:
                    from ..utils.transform import projection_matrix
                    proj.h_proj = (C.c_float * 12)(*[float(x) for x in projection_matrix(calib).reshape(-1)])
                    proj.d_proj = None
                if self.fv_geometry is not None:
                    (proj.fv_h, proj.fv_w, proj.fv_theta_min, proj.fv_dtheta, proj.fv_phi_max,
                     proj.fv_dphi) = self.fv_geometry.c_args()
                # MV3D_ROI_PROJECT=0: BEV / image rectangles are taken from the proposal layer's blobs (bit-identical to the
                # in-kernel projection; A/B switch for the cost of recomputing them), the FV box is always made here
                given = os.environ.get('MV3D_ROI_PROJECT', '1') == '0'
                for k, t in enumerate(targets):
                    views[k].source = ROI_GIVEN if (given and t != 'fv') else {'bv': ROI_BEV, 'img': ROI_IMG, 'fv': ROI_FV}[t]
                    views[k].d_rois_out = ptr(e['fv']) if (t == 'fv' and e.pop('fv_pending', False)) else None
                p3d = e['p3d'].contiguous()
                check(lib().mv3d_roi_pool_fused(views, len(group), ptr(p3d), C.byref(proj), R, ptr(num), group[0].channels,
                                                ph, pw, current_stream()), 'mv3d_roi_pool_fused')
            else:
                check(lib().mv3d_roi_pool_multiview(views, len(group), R, ptr(num), group[0].channels, ph, pw,
                                                    current_stream()), 'mv3d_roi_pool_multiview')
            mine = None
            for m, res in zip(group, results):
                if m is node:
                    mine = res
                else:
                    m.attrs['result'] = res
            return mine
        n = self._node(name, 'roi_pool', [data, rois], run, channels=data.channels,
                       pooled=(pooled_height, pooled_width))
        n.attrs['cfg'] = (pooled_height, pooled_width, float(spatial_scale))
        self._roi_nodes.append(n)
        return n

    @layer
    def fc(self, input, num_out, name, relu=True, trainable=True):
        if isinstance(input, tuple):
            input = input[0]
        if input.pooled is not None:
            ph, pw = input.pooled
            dim = ph * pw * input.channels
            cc = input.channels
            # the reference flattens NHWC pooled maps in (C,H,W) order (network.py:381); our pooled rows are
            # (H,W,C), so the weight rows are permuted once instead of transposing activations every frame
            transform = lambda w: w.view(cc, ph * pw, num_out).permute(1, 0, 2).reshape(dim, num_out).contiguous()
        else:
            dim, transform = input.channels, None
        self._declare(name, (dim, num_out), 0.001 if name == 'bbox_pred' else 0.01)

        def split_for(M, n_out, k):
            # few output tiles + long K (weight streaming): split K across CTAs to fill the 148 SMs
            if self.precise:
                bn = 128 if n_out > 64 else (64 if n_out > 32 else 32)   # mirrors dispatch_bn in conv_gemm_tcgen05.cu
            else:
                bn = 256 if n_out > 128 else (128 if n_out > 64 else (64 if n_out > 32 else 32))
            ctas = ((M + 127) // 128) * ((n_out + bn - 1) // bn)
            if k < 2048:
                return 1
            return max(1, min(8, 148 // max(1, ctas), k // 512))

        def run(vals, node):
            cached = node.attrs.pop('result', None)
            if cached is not None:
                return cached
            v = vals[node.inputs[0]]
            want_vec = any(c in ('fc', 'concat', 'dropout') for c in node.consumers)
            want_f32 = (not want_vec) or node.attrs.get('fetched', False) or \
                any(c not in ('fc', 'concat', 'dropout') for c in node.consumers)
            M = v.hi.shape[0]
            followers = node.attrs.get('fused_followers', [])
            if followers:  # sibling heads on the same input (cls_score + bbox_pred): ONE GEMM over concatenated weights
                key = name + '+' + '+'.join(f.name for f in followers)
                pw = self._packed.get(key)
                if pw is None:
                    ws = [self.params[name]] + [self.params[f.name] for f in followers]
                    wcat = torch.cat([p['weights'] for p in ws], dim=1).contiguous()
                    bcat = torch.cat([p['biases'] for p in ws], dim=0).contiguous()
                    pw = self._packed[key] = K.pack_weights(wcat, bcat)
                _, _, f32 = K.linear(v.hi, v.lo, pw, relu=False, precise=self.precise, out_bf16=False, out_f32=True,
                                     split_k=split_for(M, pw.cout, dim))
                off = num_out
                for f in followers:
                    f.attrs['result'] = Val(dense=f32[:, off:off + f.channels])
                    off += f.channels
                return Val(dense=f32[:, :num_out])
            a_fmt = v.extra.get('fmt', K.FMT_BF16X2) if isinstance(v.extra, dict) else K.FMT_BF16X2
            hi, lo, f32 = K.linear(v.hi, v.lo, self._weight(name, transform, fmt=a_fmt), relu=relu, precise=self.precise,
                                   out_bf16=want_vec, out_f32=want_f32, split_k=max(2, split_for(M, num_out, dim))
                                   if a_fmt == K.FMT_F16E5 else split_for(M, num_out, dim))
            return Val(hi=hi, lo=lo, dense=f32)
        n = self._node(name, 'fc', [input], run, channels=num_out)
        n.attrs['relu'] = relu
        if input.pooled is not None:   # rows of the reference weight are in (C,H,W) order (network.py:381)
            n.attrs['flatten_chw'] = (input.channels, input.pooled[0], input.pooled[1])
        # fusion planning: an earlier fc without ReLU fed by the same tensors (directly or through identical concats)
        if not relu:
            def src(x):
                return tuple(id(i) for i in x.inputs) if x.kind == 'concat' else (id(x),)
            for m in self._program:
                if m is not n and m.kind == 'fc' and not m.attrs.get('relu', True) and 'fused_into' not in m.attrs \
                        and src(m.inputs[0]) == src(input):
                    m.attrs.setdefault('fused_followers', []).append(n)
                    n.attrs['fused_into'] = m
                    break
        return n

    @layer
    def concat(self, inputs, axis, name):
        assert axis == 1

        def run(vals, node):
            vs = [vals[i] for i in node.inputs]
            hi = torch.cat([v.hi for v in vs], dim=1)
            lo = torch.cat([v.lo for v in vs], dim=1) if vs[0].lo is not None else None
            return Val(hi=hi, lo=lo)
        return self._node(name, 'concat', list(inputs), run, channels=sum(i.channels for i in inputs))

    @layer
    def dropout(self, input, keep_prob, name):
        if isinstance(input, tuple):
            input = input[0]

        def run(vals, node):
            v = vals[node.inputs[0]]
            kp = keep_prob
            if isinstance(kp, Node):   # the keep_prob placeholder (MV3D_train.py:19)
                kv = vals.get(kp)
                kp = 1.0 if kv is None else float(np.asarray(kv.extra).reshape(-1)[0])
            node.attrs['keep_prob_value'] = float(kp)
            if not self.training or kp >= 1.0:
                return v   # inference / parity runs: identity
            hi = v.hi.clone()
            lo = v.lo.clone() if v.lo is not None else None
            self.dropout_seed += 1
            check(lib().mv3d_dropout(ptr(hi), ptr(lo), hi.shape[0], node.channels, hi.shape[1], float(kp),
                                     int(self.dropout_seed) * 0x9E3779B97F4A7C15 % (1 << 64), current_stream()),
                  'mv3d_dropout')
            return Val(hi=hi, lo=lo)
        return self._node(name, 'dropout', [input], run, channels=input.channels, pooled=input.pooled)

    # ------------------------------------------------------------------ training-only layers (MV3D_train.py)
    @staticmethod
    def _per_frame(x, B):
        """Ground-truth feeds: one array for the reference's single-frame batch, or a list with one array per frame."""
        if isinstance(x, (list, tuple)):
            assert len(x) == B
            return list(x)
        assert B == 1, 'feed a list with one array per frame for multi-frame batches'
        return [x]

    @layer
    def anchor_target_layer(self, input, _feat_stride, anchor_scales, name):
        """network.py:237-253 -> (rpn_labels, rpn_bbox_targets, rois_bv, rois_3d)."""
        def run(vals, node):
            score = vals[node.inputs[0]].dense
            B, H, W = score.shape[0], score.shape[1], score.shape[2]
            gt_bv = self._per_frame(vals[node.inputs[1]].extra, B)
            gt_3d = self._per_frame(vals[node.inputs[2]].extra, B)
            im_info = np.asarray(vals[node.inputs[3]].extra, dtype=np.float32).reshape(-1, 3)
            key = (H, W, int(_feat_stride))
            al = self._anchor_target_layers.get(key)
            if al is None:
                al = self._anchor_target_layers[key] = AnchorTargetLayer(H, W, int(_feat_stride), geom=self.geometry,
                                                                         device=self.device)
            pre = self._anchor_pre
            self._anchor_pre = None
            if pre is not None and pre[0] == (B, H, W, int(_feat_stride)):
                outs = pre[1]   # computed ahead of the trunks by precompute_anchor_targets (same RNG draws, same order)
            else:
                outs = [al(self._dev(gt_bv[b]), self._dev(gt_3d[b]), im_info[min(b, im_info.shape[0] - 1)],
                           want_rois=node.attrs.get('fetched', False)) for b in range(B)]
            return Val(extra=dict(labels=torch.stack([o['labels'] for o in outs]),
                                  targets=torch.stack([o['targets'] for o in outs]),
                                  counts=torch.stack([o['counts'] for o in outs]), outs=outs, A=al.N // (H * W)))
        n = self._node(name, 'anchor_target', list(input), run)
        return (n, n, n, n)

    @layer
    def proposal_target_layer_3d(self, input, classes, name):
        """network.py:256-273 -> (rois_bv, rois_img, labels, bbox_targets, rois_3d)."""
        def run(vals, node):
            e = vals[node.inputs[0]].extra
            outs = e['outs']
            B = len(outs)
            gt_bv = self._per_frame(vals[node.inputs[1]].extra, B)
            gt_3d = self._per_frame(vals[node.inputs[2]].extra, B)
            gt_cnr = self._per_frame(vals[node.inputs[3]].extra, B)
            calib = np.asarray(vals[node.inputs[4]].extra, dtype=np.float32)
            if self._proposal_target_layer is None:
                self._proposal_target_layer = ProposalTargetLayer3D(device=self.device)
            res = []
            for b in range(B):
                cb = calib.reshape(-1, 4, 12)[b if calib.size > 48 else 0]
                res.append(self._proposal_target_layer(outs[b]['bv'], outs[b]['p3d'], outs[b]['num'],
                                                       self._dev(gt_bv[b]), self._dev(gt_3d[b]), self._dev(gt_cnr[b]),
                                                       cb, int(classes), batch_index=float(b)))
            cat = (lambda k: res[0][k]) if B == 1 else (lambda k: torch.cat([r[k] for r in res]))
            counts = torch.tensor([r['bv'].shape[0] for r in res], dtype=torch.int32).to(self.device)
            return Val(extra=dict(bv=cat('bv'), img=cat('img'), p3d=cat('p3d'), labels=cat('labels'),
                                  targets=cat('targets'), num=None, frame_counts=counts, B=B))
        n = self._node(name, 'proposal_target', list(i[0] if isinstance(i, tuple) else i for i in input), run)
        return (n, n, n, n, n)

    def precompute_anchor_targets(self, B, Hf, Wf, feat_stride, gt_bv, gt_3d, im_info):
        """anchor_target_layer reads only the ground truth and the feature-map SIZE (anchor_target_layer_tf.py:59-98: the
        score tensor is used for its shape), so a training step can run it -- kernel, 35 kB D2H, host `npr.choice` -- on a
        side stream BEFORE the trunks are enqueued instead of stalling the pipeline behind them.  The draws from numpy's
        global stream happen in the same order as when the node runs in place (before the proposal-target layer)."""
        key = (Hf, Wf, int(feat_stride))
        al = self._anchor_target_layers.get(key)
        if al is None:
            al = self._anchor_target_layers[key] = AnchorTargetLayer(Hf, Wf, int(feat_stride), geom=self.geometry,
                                                                     device=self.device)
        gt_bv, gt_3d = self._per_frame(gt_bv, B), self._per_frame(gt_3d, B)
        im_info = np.asarray(im_info, dtype=np.float32).reshape(-1, 3)
        if self._aux_stream is None:
            self._aux_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._aux_stream):
            outs = [al(self._dev(gt_bv[b]), self._dev(gt_3d[b]), im_info[min(b, im_info.shape[0] - 1)]) for b in range(B)]
        main.wait_stream(self._aux_stream)
        self._anchor_pre = ((B, Hf, Wf, int(feat_stride)), outs)

    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.device, dtype=torch.float32)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    # ------------------------------------------------------------------ execution (the sess.run analogue)
    def run(self, fetches, feed_dict):
        """Execute the recorded program.  `fetches`: list of Nodes / layer names; feed_dict: {placeholder Node or
        name: numpy array or CUDA tensor}.  Returns a list of CUDA tensors (float32) -- no host sync."""
        fetch_nodes = []
        for f in fetches:
            n = self.layers[f] if isinstance(f, str) else f
            n = n[0] if isinstance(n, tuple) else n
            fetch_nodes.append(n)
        for n in self._program:
            n.attrs['fetched'] = n in fetch_nodes
            n.attrs.pop('result', None)
        vals: Dict[Node, Val] = {}
        for k, v in feed_dict.items():
            node = self.layers[k] if isinstance(k, str) else k
            if isinstance(v, K.PadAct):  # already in the trunk's input layout (BevRasterizer.to_pad)
                vals[node] = Val(pad=v)
            elif node.name.startswith('gt_'):
                vals[node] = Val(extra=v)   # numpy / tensor, or a list with one array per frame
            elif node.name == 'calib' and isinstance(v, torch.Tensor) and v.is_cuda:
                vals[node] = Val(extra=v)   # (12,) / (B,12) projection floats already on the device (graph replay)
            elif node.name in ('im_info', 'calib', 'keep_prob'):
                vals[node] = Val(extra=v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            else:
                t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
                vals[node] = Val(dense=t.to(self.device, dtype=torch.float32).contiguous())
        needed = self._needed(fetch_nodes)
        self._needed_now = needed
        main = torch.cuda.current_stream()
        # independent branches (attrs['side'] = k > 0: the RGB and FV trunks) run on their own streams, forked when first
        # reached and joined before the first node that may consume any of them
        use_side = self.use_side_stream
        forked, joined = {}, set()

        def join_all():
            for k, st in forked.items():
                if k not in joined:
                    main.wait_stream(st)
                    joined.add(k)
        for n in self._program:
            if n not in needed:
                continue
            k = n.attrs.get('side') if use_side else None
            if k:
                st = forked.get(k)
                if st is None:
                    if self._side_stream is None:
                        self._side_stream = {}
                    st = self._side_stream.get(k)
                    if st is None:
                        st = self._side_stream[k] = torch.cuda.Stream()
                    st.wait_stream(main)   # the branch input (fed before the loop) is ready
                    forked[k] = st
                with torch.cuda.stream(st):
                    vals[n] = n.fn(vals, n)
                continue
            # (roi_pool launches are fused across views, so any roi_pool node may read a side branch's output)
            if forked and (n.kind == 'roi_pool' or any(isinstance(i, Node) and i.attrs.get('side') for i in n.inputs)):
                join_all()
            if self.node_events is not None:   # measurement hook (tools/node_times.py): CUDA events around every node
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
                vals[n] = n.fn(vals, n)
                e1.record(main)
                self.node_events.append((n.name, n.kind, e0, e1))
                continue
            vals[n] = n.fn(vals, n)
        join_all()
        self.last_vals = vals if self.training else None
        out = []
        for n in fetch_nodes:
            v = vals[n]
            if v.dense is None and v.pad is None and v.hi is None:
                out.append(v.extra)   # host-layer outputs (target layers): the dict of device tensors
            elif v.dense is not None:
                out.append(v.dense)
            elif v.pad is not None:
                out.append(K.unpad_nhwc(v.pad))
            else:
                out.append(v.hi.float() + (v.lo.float() if v.lo is not None else 0))
        return out

    def _needed(self, fetch_nodes):
        need, stack = set(), list(fetch_nodes)
        while stack:
            n = stack.pop()
            if n in need or not isinstance(n, Node):
                continue
            need.add(n)
            if 'fused_into' in n.attrs:
                stack.append(n.attrs['fused_into'])
            else:
                stack.extend(n.inputs)
        return need
