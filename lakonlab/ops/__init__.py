"""`lakonlab.ops` — in the reference this package holds its only native code (gmflow_ops, unused by ArcFlow);
here it is the slot for the sm_100a extension (BASELINE.json north_star). Operators are thin re-exports."""
from arcflow_b200.ops import (attention, gemm, ln_modulate, rmsnorm_rope, sampler_step, small_linear,  # noqa: F401
                              timestep_embed)
from arcflow_b200 import AfbError  # noqa: F401
