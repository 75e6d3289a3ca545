# -*- coding: utf-8 -*-
"""B200-native (sm_100a) generation path of patrickltobing/shallow-wavenet.

``CSWNV`` / ``DSWNV`` keep the reference's constructor, ``state_dict`` layout and
``batch_fast_generate`` call; everything they compute runs in the hand-written CUDA library
behind ``include/swn_b200.h`` (``lib/libswn_b200.so``).
"""
from .nets import CSWNV, DSWNV, encode_mu_law, decode_mu_law, initialize  # noqa: F401
from . import _abi  # noqa: F401

__all__ = ["CSWNV", "DSWNV", "encode_mu_law", "decode_mu_law", "initialize"]
