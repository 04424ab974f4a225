"""MV3D inference network: same layer names, wiring and placeholders as lib/networks/MV3D_test.py:8-123.
`bv_channels` generalises the BEV depth (9 in the reference, 36 for the BASELINE 700x800x36 grid)."""
from .network import Network

n_classes = 2
_feat_stride = [8, 8]
anchor_scales = [1, 1]


class MV3D_test(Network):
    def __init__(self, trainable=True, bv_channels=9, fv=False, **kw):
        """fv=True adds the paper's third view (front view): its own VGG16 trunk ('*_3' layers) on `lidar_fv_data`,
        an FV ROI per proposal, fc6_3/fc7_3 and a 6144-wide fusion.  The reference has no such branch
        (network.py:313-315); with fv=False the graph is exactly MV3D_test.py."""
        self.with_fv = bool(fv)
        if self.with_fv and kw.get('fv_geometry') is None:
            from ..utils.transform import FV_GEOMETRY
            kw['fv_geometry'] = FV_GEOMETRY
        self.lidar_fv_data = Network.placeholder('lidar_fv_data', 3)
        self.lidar_bv_data = Network.placeholder('lidar_bv_data', bv_channels)
        self.image_data = Network.placeholder('image_data', 3)
        self.im_info = Network.placeholder('im_info')
        self.gt_boxes = Network.placeholder('gt_boxes')
        self.gt_boxes_bv = Network.placeholder('gt_boxes_bv')
        self.gt_boxes_3d = Network.placeholder('gt_boxes_3d')
        self.gt_boxes_corners = Network.placeholder('gt_boxes_corners')
        self.calib = Network.placeholder('calib')
        self.keep_prob = Network.placeholder('keep_prob')
        inputs = {'lidar_bv_data': self.lidar_bv_data, 'lidar_fv_data': self.lidar_fv_data,
                  'image_data': self.image_data, 'calib': self.calib,
                  'im_info': self.im_info, 'gt_boxes': self.gt_boxes, 'gt_boxes_bv': self.gt_boxes_bv,
                  'gt_boxes_3d': self.gt_boxes_3d, 'gt_boxes_corners': self.gt_boxes_corners}
        super().__init__(inputs, trainable=trainable, **kw)

    def _vgg_trunk(self, source, suffix, side=False):
        s = suffix
        first = len(self._program)
        (self.feed(source)
             .conv(3, 3, 64, 1, 1, name='conv1_1' + s)
             .conv(3, 3, 64, 1, 1, name='conv1_2' + s)
             .max_pool(2, 2, 2, 2, padding='VALID', name='pool1' + s)
             .conv(3, 3, 128, 1, 1, name='conv2_1' + s)
             .conv(3, 3, 128, 1, 1, name='conv2_2' + s)
             .max_pool(2, 2, 2, 2, padding='VALID', name='pool2' + s)
             .conv(3, 3, 256, 1, 1, name='conv3_1' + s)
             .conv(3, 3, 256, 1, 1, name='conv3_2' + s)
             .conv(3, 3, 256, 1, 1, name='conv3_3' + s)
             .max_pool(2, 2, 2, 2, padding='VALID', name='pool3' + s)
             .conv(3, 3, 512, 1, 1, name='conv4_1' + s)
             .conv(3, 3, 512, 1, 1, name='conv4_2' + s)
             .conv(3, 3, 512, 1, 1, name='conv4_3' + s)
             .conv(3, 3, 512, 1, 1, name='conv5_1' + s)
             .conv(3, 3, 512, 1, 1, name='conv5_2' + s)
             .conv(3, 3, 512, 1, 1, name='conv5_3' + s))
        if side:  # independent of the other trunks until roi_pool: eligible for its own stream
            for n in self._program[first:]:
                n.attrs['side'] = side

    def setup(self):
        self._vgg_trunk('lidar_bv_data', '')     # MV3D_test.py:33-49
        self._vgg_trunk('image_data', '_2', side=1)      # :51-67
        if self.with_fv:
            self._vgg_trunk('lidar_fv_data', '_3', side=2)
        # ========= RPN ============  (:70-86)
        (self.feed('conv5_3')
             .conv(3, 3, 512, 1, 1, name='rpn_conv/3x3')
             .conv(1, 1, len(anchor_scales) * 2 * 2, 1, 1, padding='VALID', relu=False, name='rpn_cls_score'))
        (self.feed('rpn_conv/3x3')
             .conv(1, 1, len(anchor_scales) * 2 * 6, 1, 1, padding='VALID', relu=False, name='rpn_bbox_pred'))
        (self.feed('rpn_cls_score')
             .reshape_layer(2, name='rpn_cls_score_reshape')
             .softmax(name='rpn_cls_prob'))
        (self.feed('rpn_cls_prob')
             .reshape_layer(len(anchor_scales) * 2 * 2, name='rpn_cls_prob_reshape'))
        (self.feed('rpn_cls_prob_reshape', 'rpn_bbox_pred', 'im_info', 'calib')
             .proposal_layer_3d(_feat_stride[0], 'TEST', name='rois'))
        (self.feed('rois').proposal_transform(target='img', name='roi_data_img'))
        (self.feed('rois').proposal_transform(target='bv', name='roi_data_bv'))
        if self.with_fv:   # before the first roi_pool so that all three views are pooled by ONE launch
            (self.feed('rois').proposal_transform(target='fv', name='roi_data_fv'))
        # ========= RoI Proposal ============  (:103-123)
        (self.feed('conv5_3', 'roi_data_bv')
             .roi_pool(7, 7, 1.0 / 8, name='pool_5')
             .fc(2048, name='fc6_1')
             .fc(2048, name='fc7_1'))
        (self.feed('conv5_3_2', 'roi_data_img')
             .roi_pool(7, 7, 1.0 / 8, name='pool_5_2')
             .fc(2048, name='fc6_2')
             .fc(2048, name='fc7_2'))
        branches = ['fc7_1', 'fc7_2']
        if self.with_fv:
            (self.feed('conv5_3_3', 'roi_data_fv')
                 .roi_pool(7, 7, 1.0 / 8, name='pool_5_3')
                 .fc(2048, name='fc6_3')
                 .fc(2048, name='fc7_3'))
            branches.append('fc7_3')
        (self.feed(*branches)
             .concat(axis=1, name='concat1')
             .fc(n_classes, relu=False, name='cls_score')
             .softmax(name='cls_prob'))
        (self.feed(*branches)
             .concat(axis=1, name='concat2')
             .fc(n_classes * 24, relu=False, name='bbox_pred'))
