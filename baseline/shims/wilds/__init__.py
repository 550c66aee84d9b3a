"""Stub: the WILDS package (fMoW loader dependency) is not installed offline; the SEA headline run never calls it."""


def get_dataset(*a, **k):
    raise RuntimeError("wilds is not available in this image")
