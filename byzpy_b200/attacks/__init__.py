"""Byzantine attack simulators (reference package ``byzpy.attacks``)."""
from .base import Attack
from .empire import EmpireAttack
from .gaussian import GaussianAttack
from .inf import InfAttack
from .label_flip import LabelFlipAttack
from .little import LittleAttack
from .mimic import MimicAttack
from .sign_flip import SignFlipAttack

__all__ = ["Attack", "EmpireAttack", "LittleAttack", "SignFlipAttack", "LabelFlipAttack",
           "GaussianAttack", "InfAttack", "MimicAttack"]
