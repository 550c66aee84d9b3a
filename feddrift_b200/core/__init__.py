"""L1 core runtime: Message / Observer / managers / transports / topology / robustness."""
from .message import DeviceRef, Message
from .observer import Observer
from .managers import ClientManager, ServerManager
from .comm import BaseCommunicationManager, InProcCommunicationManager, LocalBroker, MqttCommManager, World
from .topology import AsymmetricTopologyManager, BaseTopologyManager, SymmetricTopologyManager
from .robustness import RobustAggregator, is_weight_param, vectorize_weight

__all__ = [
    "DeviceRef", "Message", "Observer", "ClientManager", "ServerManager", "BaseCommunicationManager",
    "InProcCommunicationManager", "LocalBroker", "MqttCommManager", "World", "AsymmetricTopologyManager",
    "BaseTopologyManager", "SymmetricTopologyManager", "RobustAggregator", "is_weight_param", "vectorize_weight",
]
