"""Observer ABC (parity: ``fedml_core/distributed/communication/observer.py:4-7``)."""
from abc import ABC, abstractmethod


class Observer(ABC):
    @abstractmethod
    def receive_message(self, msg_type, msg_params) -> None:
        ...
