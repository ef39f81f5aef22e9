"""look2hear.system mirror for the hot path: the sharded separation runner (inference over a batch
split across ranks).  The training step (`AudioLightningModule.training_step`,
audio_litmodule.py:83-124) needs the backward kernels and is not part of this build yet."""
from .sharding import shard_bounds, separate_sharded

__all__ = ["shard_bounds", "separate_sharded"]
