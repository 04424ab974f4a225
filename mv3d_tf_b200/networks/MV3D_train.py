"""MV3D training network: same layer names, wiring and placeholders as lib/networks/MV3D_train.py:8-182
(including its re-use of the name 'drop7' for the fused-feature dropout, :175-182, so that `bbox_pred` reads the
4096-wide fused tensor).  `bv_channels` generalises the BEV depth (9 in the reference)."""
from .MV3D_test import MV3D_test
from .network import Network

n_classes = 2  # background, car
_feat_stride = [8, 8]
anchor_scales = [1.0, 1.0]


class MV3D_train(MV3D_test):
    def setup(self):
        self._vgg_trunk('lidar_bv_data', '')     # MV3D_train.py:44-61
        self._vgg_trunk('image_data', '_2', side=1)      # :63-80 (independent of the BEV trunk: its own stream)
        # ========= RPN ============  (:84-110)
        (self.feed('conv5_3')
             .conv(3, 3, 512, 1, 1, name='rpn_conv/3x3')
             .conv(1, 1, len(anchor_scales) * 2 * 2, 1, 1, padding='VALID', relu=False, name='rpn_cls_score'))
        (self.feed('rpn_cls_score', 'gt_boxes_bv', 'gt_boxes_3d', 'im_info')
             .anchor_target_layer(_feat_stride[0], anchor_scales, name='rpn_data'))
        (self.feed('rpn_conv/3x3')
             .conv(1, 1, len(anchor_scales) * 2 * 6, 1, 1, padding='VALID', relu=False, name='rpn_bbox_pred'))
        (self.feed('rpn_cls_score')
             .reshape_layer(2, name='rpn_cls_score_reshape')
             .softmax(name='rpn_cls_prob'))
        (self.feed('rpn_cls_prob')
             .reshape_layer(len(anchor_scales) * 2 * 2, name='rpn_cls_prob_reshape'))
        (self.feed('rpn_cls_prob_reshape', 'rpn_bbox_pred', 'im_info', 'calib')
             .proposal_layer_3d(_feat_stride[0], 'TRAIN', name='rpn_rois'))
        (self.feed('rpn_rois', 'gt_boxes_bv', 'gt_boxes_3d', 'gt_boxes_corners', 'calib')
             .proposal_target_layer_3d(n_classes, name='roi_data_3d'))
        (self.feed('roi_data_3d').proposal_transform(target='img', name='roi_data_img'))
        (self.feed('roi_data_3d').proposal_transform(target='bv', name='roi_data_bv'))
        # ========= RCNN ============  (:158-182)
        (self.feed('conv5_3', 'roi_data_bv')
             .roi_pool(7, 7, 1.0 / 8, name='pool_5')
             .fc(2048, name='fc6_1')
             .dropout(self.keep_prob, name='drop6')
             .fc(2048, name='fc7_1')
             .dropout(self.keep_prob, name='drop7'))
        (self.feed('conv5_3_2', 'roi_data_img')
             .roi_pool(7, 7, 1.0 / 8, name='pool_5_2')
             .fc(2048, name='fc6_2')
             .dropout(self.keep_prob, name='drop6_2')
             .fc(2048, name='fc7_2')
             .dropout(self.keep_prob, name='drop7_2'))
        (self.feed('drop7', 'drop7_2')
             .concat(axis=1, name='concat1')
             .dropout(self.keep_prob, name='drop7')
             .fc(n_classes, relu=False, name='cls_score')
             .softmax(name='cls_prob'))
        (self.feed('drop7')
             .fc(n_classes * 24, relu=False, name='bbox_pred'))  # (x0-x7, y0-y7, z0-z7)
