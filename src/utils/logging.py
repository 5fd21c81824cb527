"""Repair of a broken import in the reference: /root/reference/scripts/train.py:15 does
`from src.utils.logging import setup_logging, log_system_info, log_dataset_info`, but the file the reference ships is
`src/utils/custom_logging.py` - the script dies at import.  This module forwards the three names to the reference's own
implementation, so `scripts/train.py` starts unchanged when this repository precedes the reference on PYTHONPATH."""
from src.utils.custom_logging import log_dataset_info, log_system_info, setup_logging  # noqa: F401

__all__ = ["setup_logging", "log_system_info", "log_dataset_info"]
