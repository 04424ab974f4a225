"""Training data layer (lib/roi_data_layer of the reference: roidb.py, layer.py, minibatch_mv3d.py)."""
