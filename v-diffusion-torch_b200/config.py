"""Config handling of the reference's sampling driver: defaults merge (utils.py:193-201), the dataset
constants generate.py reads from DATA_INFO (datasets.py:96-149), and the model / diffusion construction of
generate.py:48-98."""
import json

# channels / resolution / num_classes / multitags only (datasets.py:96-149); loaders are out of scope
DATA_INFO = {
    "mnist": {"channels": 1, "resolution": (32, 32), "num_classes": 10},
    "cifar10": {"channels": 3, "resolution": (32, 32), "num_classes": 10},
    "celeba": {"channels": 3, "resolution": (64, 64), "num_classes": 40, "multitags": True},
}


def fill_with_defaults(config, defaults):
    """Recursively copy keys that are missing or None (utils.py:193-201)."""
    for k, v in defaults.items():
        if isinstance(v, dict):
            fill_with_defaults(config.setdefault(k, dict()), v)
        elif config.get(k) is None:
            config[k] = v
    return config


def load_config(config_path, default_config_path):
    with open(config_path) as f:
        config = json.load(f)
    with open(default_config_path) as f:
        defaults = json.load(f)
    return fill_with_defaults(config, defaults)


def build_from_config(config, use_cfg, w_guide, sample_timesteps, uncond=False):
    """generate.py:54-98: returns (diffusion, model, image_shape_without_batch)."""
    from .diffusion import GaussianDiffusion, get_logsnr_schedule
    from .unet import UNet
    info = DATA_INFO[config["data"]["name"]]
    in_channels, res = info["channels"], info["resolution"][0]
    multitags = info.get("multitags", False)
    num_classes = info["num_classes"] if use_cfg else 0
    w = (0. if uncond else w_guide) if use_cfg else 0
    d = dict(config["diffusion"])
    logsnr_fn = get_logsnr_schedule(d.pop("logsnr_schedule"), d.pop("logsnr_min"), d.pop("logsnr_max"),
                                    rescale=d.pop("allow_rescale"))
    d["sample_timesteps"] = sample_timesteps
    d.pop("train_timesteps")
    diffusion = GaussianDiffusion(logsnr_fn=logsnr_fn, w_guide=w, **d)
    out_channels = (2 if d.get("model_out_type", "both") == "both" else 1) * in_channels
    model = UNet(out_channels=out_channels, num_classes=num_classes, multitags=multitags, **config["model"])
    return diffusion, model, (in_channels, res, res)
