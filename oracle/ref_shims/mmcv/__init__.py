"""mmcv shim (cross_attn.py:9 imports mmcv.cnn.ConvModule)."""
