from .arena import ClientArena, ModelBank

__all__ = ["ClientArena", "ModelBank"]
