"""Drop-in for /root/reference/src/models/vit.py: same public names, computed by libgvit (sm_100a)."""
from graph_augmented_vision_transformers_b200.modules import (Attention, Block, DropPath, Mlp, PatchEmbed,  # noqa: F401
                                                              PatchGraphLayer, VisionTransformer)
