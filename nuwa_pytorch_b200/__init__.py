"""nuwa_pytorch_b200 -- B200-native (sm_100a) implementation of the NUWA hot paths.

Public surface mirrors nuwa_pytorch/__init__.py:1-5 of the reference for the in-scope classes
(NUWA, NUWASketch, Sparse3DNA, VQGanVAE).  NUWAVideoAudio and the trainers are out of scope; the
pre-tokenised video-index data format either side of the path lives in `nuwa_pytorch_b200.data`, the fused trainer
step in `nuwa_pytorch_b200.optim`.
"""
from .nuwa import NUWA, NUWASketch, Sparse3DNA
from .vqgan_vae import VQGanVAE

__version__ = "0.1.0"
__all__ = ["NUWA", "NUWASketch", "Sparse3DNA", "VQGanVAE"]
