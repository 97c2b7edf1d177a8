"""History buffer of generated slices for the discriminator updates (trainer/utils.py:120-140): the first `max_size` fakes pass
through; afterwards each new fake is swapped with a random stored one with probability 1/2, driven by Python's `random` exactly
like the reference, so seeded runs pick the same slots.

The reference interleaves the random draws with tensor work on the host.  The draws do not depend on the data, so they are
split off here: `plan()` makes the host decisions (same `random` call sequence), `apply()` moves the slices on the device with
index tensors -- stream-ordered, allocation-stable and therefore capturable in a CUDA graph.  That is what lets a whole Cyc
iteration (generator phase AND both discriminator phases) be one graph, with the discriminator phases running beside the
generator backward instead of after a host round trip."""
import random

import torch


class ReplayBuffer:
    def __init__(self, max_size=50):
        assert max_size > 0, "Empty buffer or trying to create a black hole. Be careful."
        self.max_size = max_size
        self.size = 0
        self.pool = None            # [max_size + 1, C, H, W]; the last slot absorbs the writes of "pass through" decisions

    # -- host: decisions -----------------------------------------------------------------------------------------------
    def plan(self, n: int):
        """Decisions for n new elements, consuming `random` like the reference.  Returns (src, dst): src[b] = pool slot whose OLD
        content is returned for element b (scratch slot = return the new element itself), dst[b] = slot the new element is stored
        in (scratch slot = not stored)."""
        scratch = self.max_size
        src, dst = [], []
        for _ in range(n):
            if self.size < self.max_size:
                src.append(scratch); dst.append(self.size)
                self.size += 1
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.max_size - 1)
                src.append(i); dst.append(i)
            else:
                src.append(scratch); dst.append(scratch)
        return src, dst

    # -- device: data movement -----------------------------------------------------------------------------------------
    def ensure_pool(self, like: torch.Tensor):
        if self.pool is None or self.pool.shape[1:] != like.shape[1:] or self.pool.dtype != like.dtype or self.pool.device != like.device:
            assert self.size == 0 or self.pool is None, "ReplayBuffer: element shape changed while the buffer holds data"
            self.pool = torch.zeros((self.max_size + 1, *like.shape[1:]), dtype=like.dtype, device=like.device)
        return self.pool

    def apply(self, data: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
        """data [n, C, H, W]; src, dst int64 device tensors [n] from plan().  Elements are processed in order (a later element of the
        same call may receive an earlier one back, as in the reference's loop)."""
        pool = self.ensure_pool(data)
        data = data.detach()
        out = torch.empty_like(data)
        for b in range(data.shape[0]):
            new = data[b:b + 1]
            s, d = src[b:b + 1], dst[b:b + 1]
            old = pool.index_select(0, s)
            torch.where((s != self.max_size).view(1, 1, 1, 1), old, new, out=out[b:b + 1])
            pool.index_copy_(0, d, new)
        return out

    def push_and_pop(self, data):
        src, dst = self.plan(data.shape[0])
        dev = data.device
        return self.apply(data, torch.tensor(src, dtype=torch.int64, device=dev), torch.tensor(dst, dtype=torch.int64, device=dev))

    @property
    def data(self):
        """The stored elements as a list of [1, C, H, W] tensors (the reference's attribute)."""
        return [] if self.pool is None else [self.pool[i:i + 1] for i in range(self.size)]
