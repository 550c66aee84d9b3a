"""Stub for paho-mqtt (imported at module load by the reference's ServerManager; the MPI backend never uses it)."""


class Client:
    def __init__(self, *a, **k):
        raise RuntimeError("paho-mqtt is not available in this image")


def base62(*a, **k):
    return "0"
