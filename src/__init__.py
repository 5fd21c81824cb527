"""Import shim: lets the reference's entry points (`from src.models.vit import VisionTransformer`,
/root/reference/scripts/train.py:12 and scripts/evaluate.py:12) resolve to the libgvit-backed modules."""
