"""B200-native reverse-diffusion sampler for MoleculeDiffusionTransformer models (sm_100a only)."""
from .diffusion import ADPM2Sampler, AEulerSampler, KarrasSampler, KarrasSchedule, XDiffusion_x, build_iter_scalars
from .generative import QMDiffusion, QMDiffusionForward
from .graphmodel import AnalogDiffusionFull, AnalogDiffusionSparse
from .screening import generate_and_score, is_novel, reverse_tokenize, tokens_to_forward_conditioning, vocabulary_table
from .unet_params import UNet1dParams, UNetCFG1dParams, UNetConfig, XUNet1d

__all__ = ["QMDiffusion", "QMDiffusionForward", "AnalogDiffusionSparse", "AnalogDiffusionFull", "UNet1dParams", "XUNet1d", "UNetConfig", "UNetCFG1dParams",
           "ADPM2Sampler", "AEulerSampler", "KarrasSampler", "KarrasSchedule", "XDiffusion_x", "build_iter_scalars", "generate_and_score",
           "tokens_to_forward_conditioning", "reverse_tokenize", "vocabulary_table", "is_novel"]
__version__ = "0.1.0"
