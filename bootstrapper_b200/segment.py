"""`bs segment` configuration resolution and dispatch — the drop-in boundary.

Same key set, defaults, override order and error behaviour as the reference's
bootstrapper/segment.py (DEFAULTS :10-62, get_seg_config :95-135, run_segmentation :138-163);
pinned by tests/golden/seg_config.json, which was produced by executing the reference file.
The ws, cc and mws methods (single shot and blockwise) run on the CUDA path.
"""
import ast
import copy

_NBH = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -9, 0], [0, 0, -9], [-3, 0, 0], [0, -27, 0], [0, 0, -27]]

DEFAULTS = {
    "ws": dict(fragments_in_xy=True, min_seed_distance=10, seed_eps=None, epsilon_agglomerate=0.0,
               filter_fragments=0.1, remove_debris=64, thresholds=[0.2, 0.35, 0.5], merge_function="mean",
               sigma=None, noise_eps=None, bias=None),
    "mws": dict(aff_neighborhood=_NBH, bias=[-0.4] * 3 + [-0.7] * 6, sigma=None, noise_eps=0.001,
                strides=[[1, 1, 1]] * 3 + [[2, 9, 9]] * 3 + [[3, 27, 27]] * 3, randomized_strides=True,
                filter_fragments=0.1, remove_debris=64, min_seed_distance=None, global_bias=[1.0, -0.5]),
    "cc": dict(threshold=0.5, sigma=None, noise_eps=None, remove_debris=64),
}

COORD_KEYS = ("roi_offset", "roi_shape", "block_shape", "context")


def parse_params(text):
    """`-p key=value` values go through literal_eval; anything unparsable stays a string."""
    try:
        return ast.literal_eval(text)
    except Exception:  # noqa: BLE001  (the reference swallows every parse failure)
        return text


def parse_shape(value):
    """Coordinates: list of ints, or a space / comma separated string; None and "roi" pass through."""
    if value is None or value == "roi":
        return value
    if isinstance(value, str):
        value = value.replace(",", " ").split()
    return [int(v) for v in value]


def get_method_params(method, params):
    out = {}
    for item in params:
        key, value = item.split("=")
        if key not in DEFAULTS[method]:
            raise ValueError(f"Invalid {method} parameter {key}")
        out[key] = parse_params(value)
    return out


def resolve_config(config, method, **kwargs):
    """The merge of get_seg_config on an already-loaded TOML dict."""
    config = copy.deepcopy(config)
    for key, value in kwargs.items():
        if key != "param" and value is not None:
            config["context" if key == "block_context" else key] = value
    params = {**copy.deepcopy(DEFAULTS[method]), **config.get(f"{method}_params", {}),
              **get_method_params(method, kwargs.get("param", ()))}
    for key in [k for k in config if k.endswith("_params")]:
        del config[key]
    for key in COORD_KEYS:
        if key in config:
            config[key] = parse_shape(config[key])
    if config.get("blockwise", False):
        if method == "cc":
            raise ValueError("Blockwise connected components is not supported!")
        if "db" not in config:
            raise ValueError("Blockwise requires a database config!")
        if "lut_dir" not in config:
            config["lut_dir"] = config["seg_dataset_prefix"].replace("segmentations", "luts")
    return {**config, **params}


def get_seg_config(config_file, method, **kwargs):
    import toml
    with open(config_file, "r") as f:
        config = toml.load(f)
    return resolve_config(config, method, **kwargs)


def run_segmentation(config_file, mode="ws", **kwargs):
    config = get_seg_config(config_file, mode, **kwargs)
    if mode == "ws":
        from .post.watershed import watershed_segmentation
        return watershed_segmentation(config)
    if mode == "cc":
        from .post.connected_components import cc_segmentation
        return cc_segmentation(config)
    if mode == "mws":
        from .post.watershed_mutex import mutex_watershed_segmentation
        return mutex_watershed_segmentation(config)
    raise ValueError(f"Unknown segmentation mode: {mode}")
