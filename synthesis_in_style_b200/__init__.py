"""B200-native (sm_100a) implementation of the Synthesis-in-Style synthetic-data hot path:
batched StyleGAN2 generator forward with activation capture -> nearest-centroid labelling -> class masks.

Public surface (mirrors the reference):
    op.fused_leaky_relu, op.FusedLeakyReLU, op.upfirdn2d          networks/stylegan2/op
    model.Generator                                               networks/stylegan2/model.py
    labelling.FactorCatalog, labelling.ClusterSegmenter           segmentation/*
    dataset_creation.{Latents, build_latent_and_noise_generator, generate_images, LabelledPairGenerator}
All compute goes through the C-ABI library lib/libsis_b200.so (include/sis_b200.h); nothing falls back to the
CPU or to PyTorch ops.
"""
from . import _lib  # noqa: F401
from .build import build_library  # noqa: F401

__version__ = '0.1.0'


def library_path() -> str:
    return _lib.LIB_PATH
