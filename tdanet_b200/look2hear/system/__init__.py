"""look2hear.system mirror for the hot path: the sharded separation runner (inference over a batch
split across ranks) and the long-form chunk / separate / stitch runner (audio_test_css.py).  The training step (`AudioLightningModule.training_step`,
audio_litmodule.py:83-124) needs the backward kernels and is not part of this build yet."""
from .longform import css_segments, separate_long
from .sharding import shard_bounds, separate_sharded

__all__ = ["shard_bounds", "separate_sharded", "css_segments", "separate_long"]
