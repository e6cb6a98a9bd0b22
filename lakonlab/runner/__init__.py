from .dynamic_iter_based_runner import CheckpointHook, DynamicIterBasedRunnerMod, TextLoggerHook  # noqa: F401
from .checkpoint import adapter_from_checkpoint, exists_ckpt, load_checkpoint  # noqa: F401
