from .checkpoint_manager import CheckpointManager

__all__ = ["CheckpointManager"]
