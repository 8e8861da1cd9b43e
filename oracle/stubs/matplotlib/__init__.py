"""oracle/stubs: matplotlib stand-in (plots are not part of any measured path)."""
from . import pyplot  # noqa: F401


def use(*a, **k):
    pass
