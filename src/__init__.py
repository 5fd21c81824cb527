"""Import shim: lets the reference's entry points (`from src.models.vit import VisionTransformer`,
/root/reference/scripts/train.py:12-16 and scripts/evaluate.py:13-16) resolve to the libgvit-backed modules while every
OTHER `src.*` module (src.data, src.training, src.utils.metrics ...) keeps coming from the reference checkout.

Put this repository BEFORE the reference on PYTHONPATH (`PYTHONPATH=<this repo>:<reference>`): `src` is then this package,
and `pkgutil.extend_path` appends the reference's `src/` directory to its search path, so only the sub-modules that
exist here (`src.models.vit`, `src.utils.logging`) are replaced - the rest of the tree is not shadowed."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
