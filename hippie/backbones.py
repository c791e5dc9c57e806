"""`from hippie.backbones import ResNet18Enc, ResNet18Dec` (reference hippie/model.py:6) on the B200 engine."""
from hippie_b200.backbones import (BasicBlockDec, BasicBlockEnc, ResizeConv1d, ResNet18Dec, ResNet18Enc,  # noqa: F401
                                   test_decoder)
