from .arcflow import ArcFlowPolicy

POLICY_CLASSES = dict(ArcFlow=ArcFlowPolicy)
