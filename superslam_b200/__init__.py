"""B200-native SuperPoint + LightGlue stereo front-end (drop-in for SuperSLAM's inference layer).

Public host-side mirror of the reference interfaces lives in :mod:`superslam_b200.frontend`;
everything there calls the C-ABI library (include/superslam_b200.h) built from csrc/.
"""
__version__ = "0.1.0"
