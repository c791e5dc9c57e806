"""Drop-in alias package: `from hippie.model import MultiModalCVAE` resolves to the B200 engine-backed classes
(the reference repository has no hippie/__init__.py; its scripts import hippie.model / hippie.dataloading)."""
