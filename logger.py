"""`ComMU` logger (same name and stream handler as the reference's logger.py)."""
import logging

logger = logging.getLogger("ComMU")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s | %(message)s"))
    logger.addHandler(_h)
logger.setLevel(logging.INFO)
