"""gaviko_b200 — B200-native (sm_100a) drop-in for the GAViKO 3D-ViT forward/backward hot path."""
__version__ = '0.1.0'
