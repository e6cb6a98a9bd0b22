from .ema_hook import ExponentialMovingAverageHookMod  # noqa: F401
