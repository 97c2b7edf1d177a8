"""History buffer of generated slices for the discriminator updates (host-side logic of trainer/utils.py:120-140):
the first `max_size` fakes pass through; afterwards each new fake is swapped with a random stored one with probability
1/2, driven by Python's `random` exactly like the reference so seeded runs pick the same slots."""
import random

import torch


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
