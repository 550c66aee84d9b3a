"""Message envelope of the FedML-compatible programming model.

Parity target: ``fedml_core/distributed/communication/message.py:5-74`` (dict
backed envelope with ``msg_type / sender / receiver`` header keys and an
arbitrary payload).  The B200 twist: a payload value may be a
:class:`DeviceRef` — a handle to rows of the on-device parameter arena — in
which case no tensor bytes are serialised at all; the transport only moves the
handle and the consumer reads the arena rows in place (zero-copy, stream
ordered).  Plain tensors / state_dicts keep working for the gloo and MQTT
transports.
"""
from __future__ import annotations

import json
from typing import Any, Dict

import numpy as np
import torch


class DeviceRef:
    """Handle to ``rows`` of a device parameter arena (no payload bytes)."""

    __slots__ = ("arena_id", "rows", "event")

    def __init__(self, arena_id: int, rows, event=None):
        self.arena_id = int(arena_id)
        self.rows = list(int(r) for r in rows)
        self.event = event  # optional torch.cuda.Event the consumer waits on

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"DeviceRef(arena={self.arena_id}, rows={self.rows})"


def _jsonable(value: Any) -> Any:
    """Tensors/arrays -> nested lists (the reference's ``is_mobile`` wire form,
    ``fedml_api/distributed/fedavg/utils.py:5-14``)."""
    if isinstance(value, torch.Tensor):
        return value.detach().cpu().tolist()
    if isinstance(value, np.ndarray):
        return value.tolist()
    if isinstance(value, (np.integer,)):
        return int(value)
    if isinstance(value, (np.floating,)):
        return float(value)
    if isinstance(value, dict):
        return {str(k): _jsonable(v) for k, v in value.items()}
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    if isinstance(value, DeviceRef):
        return {"__device_ref__": [value.arena_id, value.rows]}
    return value


class Message:
    MSG_ARG_KEY_OPERATION = "operation"
    MSG_ARG_KEY_TYPE = "msg_type"
    MSG_ARG_KEY_SENDER = "sender"
    MSG_ARG_KEY_RECEIVER = "receiver"

    MSG_OPERATION_SEND = "send"
    MSG_OPERATION_RECEIVE = "receive"
    MSG_OPERATION_BROADCAST = "broadcast"
    MSG_OPERATION_REDUCE = "reduce"

    MSG_ARG_KEY_MODEL_PARAMS = "model_params"

    def __init__(self, type: Any = 0, sender_id: int = 0, receiver_id: int = 0):
        self.type = type
        self.sender_id = sender_id
        self.receiver_id = receiver_id
        self.msg_params: Dict[str, Any] = {
            Message.MSG_ARG_KEY_TYPE: type,
            Message.MSG_ARG_KEY_SENDER: sender_id,
            Message.MSG_ARG_KEY_RECEIVER: receiver_id,
        }

    # -- construction from a received payload ---------------------------------
    def init(self, msg_params: Dict[str, Any]) -> "Message":
        self.msg_params = msg_params
        self._sync_header()
        return self

    def init_from_json_string(self, json_string: str) -> "Message":
        self.msg_params = json.loads(json_string)
        self._sync_header()
        return self

    def _sync_header(self) -> None:
        self.type = self.msg_params.get(Message.MSG_ARG_KEY_TYPE, self.type)
        self.sender_id = self.msg_params.get(Message.MSG_ARG_KEY_SENDER, self.sender_id)
        self.receiver_id = self.msg_params.get(Message.MSG_ARG_KEY_RECEIVER, self.receiver_id)

    # -- accessors ------------------------------------------------------------
    def get_sender_id(self):
        return self.msg_params[Message.MSG_ARG_KEY_SENDER]

    def get_receiver_id(self):
        return self.msg_params[Message.MSG_ARG_KEY_RECEIVER]

    def get_type(self):
        return self.msg_params[Message.MSG_ARG_KEY_TYPE]

    def add_params(self, key: str, value: Any) -> None:
        self.msg_params[key] = value

    add = add_params

    def get_params(self) -> Dict[str, Any]:
        return self.msg_params

    def get(self, key: str, default: Any = None) -> Any:
        return self.msg_params.get(key, default)

    # -- wire forms -----------------------------------------------------------
    def to_string(self) -> Dict[str, Any]:
        """The pickled-dict wire form used by the p2p transports."""
        return self.msg_params

    def to_json(self) -> str:
        return json.dumps(_jsonable(self.msg_params))

    def get_content(self) -> str:
        return f"{self.get_type()}: {self.msg_params}"

    def __repr__(self) -> str:  # pragma: no cover
        keys = [k for k in self.msg_params if k not in (self.MSG_ARG_KEY_TYPE, self.MSG_ARG_KEY_SENDER,
                                                        self.MSG_ARG_KEY_RECEIVER)]
        return f"Message(type={self.get_type()}, {self.get_sender_id()}->{self.get_receiver_id()}, keys={keys})"
