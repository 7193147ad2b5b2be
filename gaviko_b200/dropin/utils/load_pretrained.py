"""Drop-in for the reference's src/utils/load_pretrained.py (same public functions; no timm import, no download)."""
from gaviko_b200.utils.load_pretrained import load_pretrain, load_vanilla_pretrain, load_vanilla_pretrain_with_adapters, mapping_vit  # noqa: F401
