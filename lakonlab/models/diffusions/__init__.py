from .arcflow import ArcFlowImitationDataFree  # noqa: F401
