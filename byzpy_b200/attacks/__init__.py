"""Byzantine attack simulators (counterpart of ``byzpy.attacks``).

Names are resolved from the table below (name -> defining submodule) so that the package namespace
and ``__all__`` cannot drift apart."""
from importlib import import_module as _import_module

_WHERE = {
    "Attack": "base",
    "EmpireAttack": "empire",
    "LittleAttack": "little",
    "SignFlipAttack": "sign_flip",
    "LabelFlipAttack": "label_flip",
    "GaussianAttack": "gaussian",
    "InfAttack": "inf",
    "MimicAttack": "mimic",
}

for _name, _module in _WHERE.items():
    globals()[_name] = getattr(_import_module(f"{__name__}.{_module}"), _name)

__all__ = list(_WHERE)
