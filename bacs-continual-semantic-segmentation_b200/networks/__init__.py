"""Network-side piece of the BACS path: the background detector heads."""
from .bg_detector import BgDetector, classification_head

__all__ = ["BgDetector", "classification_head"]
