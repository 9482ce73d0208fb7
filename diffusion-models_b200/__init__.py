"""diffusion_models_b200 -- B200-native (sm_100a) implementation of the denoising hot path of
lbarseghyan/diffusion-models: `Unet` + `DenoisingDiffusion.sample()/ddim_sample()/p_sample_loop()` and the
image-/text-conditional and latent-diffusion variants, behind the reference's own Python surface.

    from diffusion_models_b200 import Unet, DenoisingDiffusion          # was: from denoising_diffusion import ...
    from diffusion_models_b200.image_conditional import Unet, ImageConditionalDenoisingDiffusion
    from diffusion_models_b200.text_conditional import Unet, TextConditionalDenoisingDiffusion
    from diffusion_models_b200.latent import LatentDiffusion, ImageConditionalLatentDiffusion

All compute goes through libddm_b200.so (include/ddm_b200.h); there is no CPU or PyTorch fallback.
"""
from .unet import Unet
from .diffusion import DenoisingDiffusion, GaussianDiffusion, ModelPrediction
from .image_conditional import ImageConditionalDenoisingDiffusion
from .text_conditional import TextConditionalDenoisingDiffusion
from .latent import LatentDiffusion, ImageConditionalLatentDiffusion, TextConditionalLatentDiffusion
from .ddim_sampler import DDIMSampler
from .learned_gaussian import LearnedGaussianDiffusion
from .distributed import sample_sharded, shard_bounds, gather_samples
from .vae import VQDecoder
from .trainer_adapter import attach_fast_sampler, FastSamplerBinding
from . import image_conditional, text_conditional, latent, sampling, trainer_adapter, vae, _lib

__version__ = "0.1.0"
__all__ = ["Unet", "DenoisingDiffusion", "GaussianDiffusion", "ModelPrediction", "ImageConditionalDenoisingDiffusion",
           "TextConditionalDenoisingDiffusion", "LatentDiffusion", "ImageConditionalLatentDiffusion",
           "TextConditionalLatentDiffusion", "DDIMSampler", "LearnedGaussianDiffusion", "sample_sharded", "shard_bounds", "gather_samples", "attach_fast_sampler", "FastSamplerBinding", "VQDecoder"]
