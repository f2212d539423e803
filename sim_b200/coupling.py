"""``Connector`` and ``Message`` value types (sim/src/simulator/coupling.rs:9-123)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class Connector:
    """coupling.rs:9-17; ``Connector::new(id, source_id, target_id, source_port, target_port)`` (:21-35)."""
    id: str
    source_id: str
    target_id: str
    source_port: str
    target_port: str

    @classmethod
    def new(cls, id, source_id, target_id, source_port, target_port):  # noqa: A002
        return cls(id, source_id, target_id, source_port, target_port)

    def to_dict(self) -> dict:
        # serde: #[serde(rename = "sourceID")] / "targetID" (coupling.rs:11-14)
        return {"id": self.id, "sourceID": self.source_id, "targetID": self.target_id,
                "sourcePort": self.source_port, "targetPort": self.target_port}


@dataclass
class Message:
    """coupling.rs:64-71; ``Message::new(source_id, source_port, target_id, target_port, time, content)``."""
    source_id: str
    source_port: str
    target_id: str
    target_port: str
    time: float
    content: str

    @classmethod
    def new(cls, source_id, source_port, target_id, target_port, time, content):
        return cls(source_id, source_port, target_id, target_port, float(time), content)
