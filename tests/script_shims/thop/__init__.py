"""Stand-in for thop (imported by train.py:24, never called)."""


def profile(*a, **k):
    raise RuntimeError('thop stand-in')
