"""look2hear.system mirror for the hot path: the sharded separation runner (inference over a batch
split across ranks), the long-form chunk / separate / stitch runner (audio_test_css.py) and the training
step (`AudioLightningModule.training_step`, audio_litmodule.py:83-124, plus the Trainer's backward /
gradient mean / clip / Adam) on the CUDA forward and backward kernels."""
from .longform import css_segments, separate_long
from .sharding import shard_bounds, separate_sharded
from .streaming import separate_pipelined
from .training import AudioLightningModule, FlatParameters, TrainingStep
from .checkpoint import ReduceLROnPlateau
from . import checkpoint

__all__ = ["shard_bounds", "separate_sharded", "separate_pipelined", "css_segments", "separate_long", "AudioLightningModule",
           "FlatParameters", "TrainingStep", "ReduceLROnPlateau", "checkpoint"]
