"""Drop-in `diffusion` package.  `create_diffusion` keeps the reference's keyword names and defaults
(/root/reference/diffusion/__init__.py:10-20: they are the API `sample.py:76` / `train.py:157` call), and returns this
repo's `SpacedDiffusion`, whose per-step arithmetic runs in libosudit.so."""
from . import gaussian_diffusion as gd  # kept importable under the reference's alias
from .gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
from .respace import SpacedDiffusion, space_timesteps

__all__ = ["create_diffusion", "SpacedDiffusion", "space_timesteps", "gd"]


def _loss_type(use_kl: bool, rescaled: bool, use_l1: bool) -> LossType:
    """KL wins over everything; otherwise L1 or MSE, in its plain or "rescaled" flavour."""
    if use_kl:
        return LossType.RESCALED_KL
    plain, scaled = (LossType.L1, LossType.RESCALED_L1) if use_l1 else (LossType.MSE, LossType.RESCALED_MSE)
    return scaled if rescaled else plain


def _variance_type(learn_sigma: bool, sigma_small: bool) -> ModelVarType:
    """The model predicts an interpolation weight between the two fixed variances unless told otherwise."""
    if learn_sigma:
        return ModelVarType.LEARNED_RANGE
    return ModelVarType.FIXED_SMALL if sigma_small else ModelVarType.FIXED_LARGE


def create_diffusion(timestep_respacing, noise_schedule="linear", use_kl=False, sigma_small=False, predict_xstart=False,
                     learn_sigma=True, rescale_learned_sigmas=False, diffusion_steps=1000, use_l1=False):
    """Build the (possibly respaced) diffusion process.

    `timestep_respacing`: "" / None keeps all `diffusion_steps`, "100" keeps 100 of them, "ddimN" uses DDIM striding,
    a comma list gives per-section counts (respace.py).  `noise_schedule`: "linear" or "squaredcos_cap_v2".
    The scripts use the defaults (epsilon prediction, learned-range variance) with `use_l1=True` for training; that
    combination is the one implemented natively, other enum values are accepted here and raise when used."""
    keep_all = timestep_respacing is None or timestep_respacing == ""
    kept = space_timesteps(diffusion_steps, [diffusion_steps] if keep_all else timestep_respacing)
    return SpacedDiffusion(
        use_timesteps=kept,
        betas=get_named_beta_schedule(noise_schedule, diffusion_steps),
        model_mean_type=ModelMeanType.START_X if predict_xstart else ModelMeanType.EPSILON,
        model_var_type=_variance_type(learn_sigma, sigma_small),
        loss_type=_loss_type(use_kl, rescale_learned_sigmas, use_l1))
