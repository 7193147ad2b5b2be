"""Drop-in for the reference's src/model/evp.py (imported unconditionally by the scripts); see gaviko_b200/model/evp.py."""
from gaviko_b200.model.evp import ExplicitVisualPrompting  # noqa: F401
