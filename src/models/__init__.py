"""`src.models`: `vit` is the libgvit-backed drop-in; anything else the reference keeps in `src/models/` stays visible."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
