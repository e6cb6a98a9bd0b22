from .train import train_model  # noqa: F401
