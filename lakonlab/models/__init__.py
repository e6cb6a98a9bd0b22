from .builder import build_model  # noqa: F401
from .latent_diffusion_text_image import LatentDiffusionTextImage  # noqa: F401
