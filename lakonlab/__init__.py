"""lakonlab — the reference's plugin surface for the ArcFlow denoising hot path, backed by arcflow_b200.

Only the path SURVEY.md §8 scopes is mirrored: `lakonlab.pipelines` (ArcFluxPipeline, load_arcflow_adapter),
`lakonlab.ops` (the native operators), `lakonlab.parallel` (batch-parallel sharding + all-gather).
Imports need torch only — no diffusers / peft / mmcv.
"""
__version__ = "0.1.0"
