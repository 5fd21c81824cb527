"""`src.utils`: the reference's utilities (metrics, visualization, gradcam, custom_logging ...) stay the reference's own -
this package only adds the module its entry point imports but never shipped (`src.utils.logging`, see logging.py)."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
