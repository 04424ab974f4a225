import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import mv3d_oracle as orc
from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
from mv3d_tf_b200.rpn_msr.proposal_layer_tf import ProposalLayer3D
cfg_from_end2end_yml(); cfg.USE_GPU_NMS=False
g = np.load('/root/repo/tests/golden/proposal.npz')
hf, wf = g['prob'].shape[1:3]
layer = ProposalLayer3D(hf, wf, 'TEST', 8, (601,601,1))
st = layer.decode(torch.from_numpy(g['prob'][0]).cuda(), torch.from_numpy(g['deltas'][0]).cuda(), g['calib'])
p3d = st['p3d'].cpu().numpy()
def ulp(a,b):
    a = np.ascontiguousarray(a,np.float32).view(np.int32).astype(np.int64); b=np.ascontiguousarray(b,np.float32).view(np.int32).astype(np.int64)
    a = np.where(a<0, -(a&0x7FFFFFFF), a); b=np.where(b<0, -(b&0x7FFFFFFF), b); return np.abs(a-b)
fin = np.isfinite(g['p3d'])
print('finite eq', np.array_equal(fin, np.isfinite(p3d)))
u = ulp(np.nan_to_num(p3d), np.nan_to_num(g['p3d']))
print('max ulp per col', u.max(0), 'frac rows same', (u.max(1)==0).mean(), 'frac per col same', (u==0).mean(0))
same = u.max(1)==0
pbv_ref = orc.clip_boxes(g['pbv'].copy(), g['im_info'][0,:2])
pbv = st['pbv'].cpu().numpy(); pimg = st['pimg'].cpu().numpy()
print('pbv eq on same', np.array_equal(pbv[same], pbv_ref[same], equal_nan=True), 'pimg eq on same', np.array_equal(pimg[same], g['pimg'][same]))
print('pbv mismatch rows total', (~((pbv==pbv_ref)|(np.isnan(pbv)&np.isnan(pbv_ref))).all(1)).sum(), 'pimg mismatch rows', (pimg!=g['pimg']).any(1).sum(), 'of', len(pbv))
# the oracle on this machine
st2 = orc.proposal_stages(g['prob'], g['deltas'], g['im_info'], g['calib'])
u2 = ulp(np.nan_to_num(st2['p3d']), np.nan_to_num(g['p3d']))
print('oracle-here vs golden: frac same', (u2.max(1)==0).mean())
