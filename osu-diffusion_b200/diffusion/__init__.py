"""Drop-in `diffusion` package: `create_diffusion` with the reference's signature
(/root/reference/diffusion/__init__.py:10-47)."""
from . import gaussian_diffusion as gd
from .respace import space_timesteps
from .respace import SpacedDiffusion


def create_diffusion(
    timestep_respacing,
    noise_schedule="linear",
    use_kl=False,
    sigma_small=False,
    predict_xstart=False,
    learn_sigma=True,
    rescale_learned_sigmas=False,
    diffusion_steps=1000,
    use_l1=False,
):
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_L1 if use_l1 else gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.L1 if use_l1 else gd.LossType.MSE
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    if timestep_respacing is None or timestep_respacing == "":
        timestep_respacing = [diffusion_steps]
    return SpacedDiffusion(
        use_timesteps=space_timesteps(diffusion_steps, timestep_respacing),
        betas=gd.get_named_beta_schedule(noise_schedule, diffusion_steps),
        model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
        model_var_type=var_type,
        loss_type=loss_type,
    )
