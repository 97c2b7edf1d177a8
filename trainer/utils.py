"""Drop-in for the hot-path parts of trainer/utils.py: smooothing_loss (fused kernel) and ReplayBuffer / get_config."""
import yaml

import _ctagan_path  # noqa: F401
from ctagan.nn import smooothing_loss  # noqa: F401
from ctagan.replay import ReplayBuffer  # noqa: F401


def get_config(config):
    with open(config, "r") as stream:
        return yaml.safe_load(stream)
