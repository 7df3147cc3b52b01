"""vec_vad_b200 -- B200 (sm_100a) implementation of the VEC_VAD hot path.

Public surface mirrors the reference:
  * ``vec_vad_b200.unet``      <- model/unet.py   (SelfCompleteNet4 / SelfCompleteNetFull / SelfCompleteNet1raw1of)
  * ``vec_vad_b200.flow_ops``  <- FlowNet2_src/models/components/ops (Correlation / Resample2d / ChannelNorm)
  * ``vec_vad_b200.vad_datasets`` <- vad_datasets.py (cube_to_train_dataset, get_foreground, ...)
All arithmetic runs in libvecvad.so (hand-written CUDA, C ABI in include/vecvad.h).
"""
__version__ = '0.1.0'
