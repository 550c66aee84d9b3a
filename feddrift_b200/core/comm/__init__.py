from .base import BaseCommunicationManager
from .inproc import InProcCommunicationManager, World
from .mqtt import LocalBroker, MqttCommManager

__all__ = ["BaseCommunicationManager", "InProcCommunicationManager", "World", "LocalBroker", "MqttCommManager",
           "DistCommunicationManager"]


def __getattr__(name):  # lazy: torch.distributed import only when needed
    if name == "DistCommunicationManager":
        from .dist import DistCommunicationManager
        return DistCommunicationManager
    raise AttributeError(name)
