from .config import Config, ConfigDict, DictAction  # noqa: F401
