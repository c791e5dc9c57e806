"""hippie_b200: B200-native (sm_100a) engine for HIPPIE's multimodal cVAE train step and embedding pass.

    from hippie_b200.model import MultiModalCVAE, MultiModalCVAETrainModule   # reference API mirror
    from hippie_b200.engine import Engine                                     # thin wrapper over the C ABI

The arithmetic lives in libhippie_b200.so (include/hippie_b200.h); there is no CPU fallback.
"""
__all__ = ["engine", "model", "dataloading", "parallel"]
