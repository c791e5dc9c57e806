from hippie_b200.model import (FusedAdamW, MultiModalCVAE, MultiModalCVAETrainModule,  # noqa: F401
                               hippieUnimodalCVAE, hippieUnimodalEmbeddingModelCVAE)
