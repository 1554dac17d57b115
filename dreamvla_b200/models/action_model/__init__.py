from .action_model import ActionModel, DiT_models  # noqa: F401
from .gaussian_diffusion import GaussianDiffusion, SpacedDiffusion, create_diffusion, space_timesteps  # noqa: F401
