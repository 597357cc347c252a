"""affganwriting_b200 — B200-native (sm_100a) generator / discriminator hot path of AFFGanWriting.

Public surface mirrors the reference's Python modules for this path:
    affganwriting_b200.blocks       <->  GAN_word/blocks.py
    affganwriting_b200.modules_tro  <->  GAN_word/modules_tro.py   (GenModel_FC, DisModel, WriterClaModel, ...)
    affganwriting_b200.network_tro  <->  GAN_word/network_tro.py   (ConTranModel step composition, no recogniser)
    affganwriting_b200.load_data    <->  GAN_word/load_data.py     (constants, label_padding)
    affganwriting_b200.install      monkey-patches the reference's modules so main_run.py / tt.* pick these up
All arithmetic runs in csrc/libaffgw.so (C ABI in include/affgw.h); there is no CPU or PyTorch fallback.
"""
from . import _lib
from .ops import check_device_errors, force_simt, precision, set_precision, weights_updated  # noqa: F401

__all__ = ["set_precision", "precision", "force_simt", "check_device_errors", "weights_updated", "launch_count", "lib_path"]


def launch_count():
    """Number of libaffgw kernels launched by this process."""
    return _lib.launch_count()


def lib_path():
    return _lib.LIB_PATH
