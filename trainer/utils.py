"""Drop-in for the hot-path parts of trainer/utils.py: smooothing_loss and ReplayBuffer (host-side, Python `random`
driven exactly like trainer/utils.py:120-140 so seeded runs pick the same history slots)."""
import random

import torch
import yaml

import _ctagan_path  # noqa: F401
from ctagan.nn import smooothing_loss  # noqa: F401


class ReplayBuffer:
    def __init__(self, max_size=50):
        assert max_size > 0, "Empty buffer or trying to create a black hole. Be careful."
        self.max_size = max_size
        self.data = []

    def push_and_pop(self, data):
        out = []
        for element in data.detach():
            element = element.unsqueeze(0)
            if len(self.data) < self.max_size:
                self.data.append(element)
                out.append(element)
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.max_size - 1)
                out.append(self.data[i].clone())
                self.data[i] = element
            else:
                out.append(element)
        return torch.cat(out)


def get_config(config):
    with open(config, "r") as stream:
        return yaml.safe_load(stream)
