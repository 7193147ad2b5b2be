"""Stand-in for deepspeed (absent here; the branch is disabled in every shipped config, train.py:158-159)."""


class DeepSpeedEngine:
    pass


def initialize(*a, **k):
    raise RuntimeError('deepspeed stand-in: the DeepSpeed branch is out of scope (SURVEY §2 #18)')
