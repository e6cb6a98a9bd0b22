from .batch_parallel import shard_batch, gather_latents  # noqa: F401
