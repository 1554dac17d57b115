from .action_model import ActionModel, ActionModelFM, DiT_models  # noqa: F401
from .gaussian_diffusion import FMDiffusion, GaussianDiffusion, SpacedDiffusion, create_diffusion, space_timesteps  # noqa: F401
