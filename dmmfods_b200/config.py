"""EasyDict-style configuration with the reference's key names and defaults.

Mirror of the config part of dmmfods/utils/Dense_U_Net_lidar_helper.py (create_config :84-211,
get_config :213-223, load_config/save_config :60-82, set_current_run :225-228) so that the
reference's agents/ training loop and the model factories read the same `config.<section>.<key>`
values.  `easydict` is not a dependency: `EasyDict` below is a small attribute-dict with the same
observable behaviour (attribute AND item access, recursive conversion of nested dicts).
"""
import json
import os
from datetime import datetime
from os.path import isfile, join
from pathlib import Path


class EasyDict(dict):
    """dict whose keys are also attributes; nested dicts (also inside lists/tuples) convert recursively."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        src = dict(d or {})
        src.update(kwargs)
        for k, v in src.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        elif isinstance(value, (list, tuple)):
            value = type(value)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x
                                for x in value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, other=None, **kwargs):
        for k, v in dict(other or {}, **kwargs).items():
            setattr(self, k, v)

    def pop(self, k, *default):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *default)


edict = EasyDict

# config.model defaults, helper:111-123
MODEL_DEFAULTS = {
    "growth_rate": 32,
    "block_config": (6, 12, 24, 16),
    "num_init_features": 64,
    "stream_1_in_channels": 3,        # rgb
    "stream_2_in_channels": 1,        # lidar (0 = single stream)
    "concat_before_block_num": 2,
    "num_layers_before_blocks": 4,
    "bn_size": 4,
    "drop_rate": 0,
    "num_classes": 3,
    "memory_efficient": False,
}


def load_json_file(filepath):
    """helper:24-38."""
    if not isfile(filepath):
        raise FileNotFoundError
    with open(filepath, "r") as jf:
        return json.load(jf)


def save_json_file(filepath, save_file, indent=None):
    """helper:40-54."""
    with open(filepath, "w") as jf:
        json.dump(save_file, jf, indent=indent)
    print("Successfully saved " + filepath)
    return 1


def load_config(loading_dir, file_name):
    """helper:60-73: the json config if present, else None."""
    json_file = join(loading_dir, file_name)
    return load_json_file(json_file) if isfile(json_file) else None


def save_config(config, file_name="config.json"):
    """helper:75-82."""
    Path(config.dir.configs).mkdir(exist_ok=True)
    save_json_file(os.path.join(config.dir.configs, file_name), config, indent=4)


def create_config(host_dir):
    """helper:84-211: same sections, keys and default values."""
    if not host_dir:
        host_dir = "/content/drive/My Drive/Colab Notebooks/DeepCV_Packages"
    config = {"dir": {"hosting": host_dir}}
    config["scripts"] = {
        "model": "Dense_U_Net_lidar.py",
        "utils": "Dense_U_Net_lidar_helper.py",
        "agent": "Dense_U_Net_lidar_Agent.py",
        "dataset": "WaymoData.py",
        "setup": "Setup.ipynb",
    }
    config["model"] = dict(MODEL_DEFAULTS)
    config["loss"] = {
        "alpha": 1, "gamma": 2, "logits": True, "reduce": False,
        "skip_v_every_n_its": False, "skip_p_every_n_its": False, "skip_b_every_n_its": False,
    }
    config["loader"] = {
        "mode": "train", "batch_size": None, "pin_memory": True, "num_workers": 4,
        "async_loading": True, "drop_last": False,
    }
    config["optimizer"] = {
        "type": "Adam", "learning_rate": 1e-3, "beta1": 0.9, "beta2": 0.999, "eps": 1e-08,
        "amsgrad": False, "weight_decay": 0,
        "lr_scheduler": {"want": False, "every_n_epochs": 30, "gamma": 0.1},
    }
    config["dataset"] = {
        "batch_size": 32,
        "label": {"1": "TYPE_VEHICLE", "2": "TYPE_PEDESTRIAN", "4": "TYPE_CYCLIST"},
        "images": {"original.size": (3, 1920, 1280), "size": (3, 192, 128)},
        "datatypes": ["images", "lidar", "labels", "heat_maps"],
        "file_list_name": "file_list.json",
    }
    config["agent"] = {
        "seed": 123, "max_epoch": 100, "iou_threshold": 0.7,
        "checkpoint": {
            "epoch": "epoch", "train_iteration": "train_iteration", "val_iteration": "val_iteration",
            "best_val_iou": "best_val_iou", "state_dict": "state_dict", "optimizer": "optimizer",
        },
        "best_checkpoint_name": "best_checkpoint.pth.tar",
    }
    root = join(host_dir, "DMMFODS", "dmmfods")
    config["dir"]["root"] = root
    for sub in ("agents", "graphs", "utils", "datasets", "configs", "experiments"):
        config["dir"][sub] = join(root, sub)
    config["dir"]["graphs"] = {"models": join(config["dir"]["graphs"], "models")}
    config["dir"]["data"] = {"root": join(host_dir, "data"), "file_lists": join(root, "data")}
    current_run = datetime.now().strftime("%Y-%m-%d-%H-%M")
    config["dir"]["current_run"] = {
        "summary": join(config["dir"]["experiments"], current_run, "summary"),
        "checkpoints": join(config["dir"]["experiments"], current_run, "checkpoints"),
    }
    return config


def get_config(host_dir="", file_name="config.json"):
    """helper:213-223: load `<host>/DMMFODS/dmmfods/configs/config.json` or create the default."""
    config = load_config(join(host_dir, "DMMFODS", "dmmfods", "configs"), file_name)
    if config is None:
        config = create_config(host_dir)
    return EasyDict(config)


def set_current_run(config, current_run):
    """helper:225-228."""
    def swap(path):
        return "/" + os.path.join(*path.split("/")[:-2], current_run, path.split("/")[-1])
    config.dir.current_run.summary = swap(config.dir.current_run.summary)
    config.dir.current_run.checkpoints = swap(config.dir.current_run.checkpoints)
    return config
