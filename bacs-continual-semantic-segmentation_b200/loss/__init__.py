"""Loss classes of the BACS path, mirroring the reference's ``loss`` package names."""
from .base_loss import BaseLoss, SeenMap
from .prototypes import Prototypes
from .experience_replay import ExperienceReplay
from .bacs_loss import BACSLoss

__all__ = ["BaseLoss", "SeenMap", "Prototypes", "ExperienceReplay", "BACSLoss"]
