"""Stand-in for colorama (utils/logging.py:142 of the reference)."""


class _Codes:
    def __getattr__(self, name):
        return ''


Fore = Style = Back = _Codes()


def init(*a, **k):
    pass
