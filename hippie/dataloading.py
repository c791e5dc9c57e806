from hippie_b200.dataloading import (BalancedBatchSampler, EphysDataset, EphysDatasetLabeled,  # noqa: F401
                                     EphysTensorDataset)
