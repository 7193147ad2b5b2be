"""Glue for bench.py's CPU reference arm (oracle only; imports nothing from gaviko_b200)."""
from oracle.shapes import gaviko_state_dict


def sd_for_backbone(backbone):
    return gaviko_state_dict(backbone)
