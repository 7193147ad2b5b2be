"""Drop-in for the reference's src/model/adaptformer.py: the same public names, implemented by gaviko_b200 (see gaviko_b200/dropin/README.md)."""
from gaviko_b200.model.adaptformer import *  # noqa: F401,F403
from gaviko_b200.model import adaptformer as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
