from _absorb import Absorb as _Absorb


def subplots(nrows=1, ncols=1, *a, **k):
    n = nrows * ncols
    axs = [_Absorb() for _ in range(n)]
    return _Absorb(), (axs[0] if n == 1 else axs)


def __getattr__(name):
    return _Absorb()
