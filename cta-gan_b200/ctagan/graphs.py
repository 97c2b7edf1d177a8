"""CUDA-graph execution of the trainer iteration bodies.

A Cyc step at batch 1 is ~1700 kernel launches for ~1.3 TFLOP: launch latency, not arithmetic, bounds an eager loop.  The
iteration body is therefore captured once (torch.cuda.CUDAGraph: all libctagan kernels, the NCCL gradient all-reduce and the
capturable Adam updates are stream-ordered and allocation-free on replay) and replayed per step.  Host-side logic that the
reference keeps on the CPU (ReplayBuffer with Python `random`, trainer/utils.py:120-140) stays on the host BETWEEN graphs:
but its random decisions do not depend on the data, so the host makes them before the replay and uploads them as an index tensor
(replay.py); the Cyc step is then ONE graph whose branches are the two generator chains, the two discriminator updates (forked as
soon as their fake exists, running beside the generator backward) and the weight-gradient lanes.

Weight packing (fp32 master -> bf16 [O][kh][kw][I]) is part of the captured graphs.  Single-graph trainers invalidate the
packed-weight cache right before capture, so every replay re-packs from the current master weights at its start.  The Cyc
step instead re-packs each network right AFTER its optimizer step (forced, so the kernels are captured there; the
discriminators only once the generator backward, the last reader of their packed weights, is done) -- the next step then
starts on its convolutions at once.  Weights edited out of band between replays (load_state_dict, in-place copies) are
detected through the parameters' version counters and re-packed before the next replay.
"""
from __future__ import annotations

import torch

from . import engine as E
from . import ops


class GraphedTrainer:
    def __init__(self, trainer, enabled: bool = True, warmup: int = 3, replay_first: bool = True):
        """warmup: eager iterations run before the capture (allocator / NCCL / optimizer-state warm-up; they are real training
        iterations on the first batch).  replay_first=False makes the first call consist of exactly those eager iterations plus the
        capture -- with warmup=1 every batch is then trained on exactly once (the epoch loop of trainer.train())."""
        self.t = trainer
        self.enabled = enabled
        self.warmup = max(int(warmup), 1)
        self.replay_first = replay_first
        self._graphs = None
        self._launches = 0
        self._sig = None
        self.is_cyc = hasattr(trainer, "phase_G")

    # -- public ----------------------------------------------------------------------------------------------------------
    def step_host(self, batch):
        """One iteration from a host (pinned) batch: H2D copy into the static inputs, then the (graphed) step."""
        tensors = self.t.load_batch(batch)
        return self._run(tensors, copy_inputs=False)

    def step_device(self, tensors):
        """One iteration from device-resident tensors."""
        return self._run(tensors, copy_inputs=True)

    def launches_per_step(self) -> int:
        return self._launches

    def _weights_signature(self):
        """Changes whenever a parameter is edited out of band (load_state_dict, `.copy_`, `.data` swaps): tensor versions and storage
        addresses of every parameter.  The captured Adam kernels update parameters in place without touching either."""
        sig = 0
        for m in self.t.__dict__.values():
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    sig = (sig * 1000003 + p._version * 31 + p.data_ptr()) & 0xFFFFFFFFFFFF
        return sig

    def refresh_weights(self):
        """Re-pack every network from its master weights now (after load_state_dict or any out-of-band edit between replays)."""
        for m in self.t.__dict__.values():
            if isinstance(m, torch.nn.Module) and hasattr(m, "prepack"):
                m.prepack(force=True)

    # -- internals -------------------------------------------------------------------------------------------------------
    def _run(self, tensors, copy_inputs):
        t = self.t
        if not self.enabled:
            return t.step(tensors=list(tensors))
        static = [t.inputs[k] for k in t.data_keys]
        if copy_inputs:
            for dst, src in zip(static, tensors):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
        if self._graphs is None:
            self._capture(static)
            if not self.replay_first:
                return t.last_losses         # the eager warm-up iteration was this call's iteration; the capture itself executes nothing
        E.invalidate_weight_cache()          # eager users after us must not trust capture-time packed weights
        sig = self._weights_signature()
        if sig != self._sig:                 # weights were edited between replays (e.g. load_state_dict): the Cyc graph re-packs only AFTER its
            self.refresh_weights()           # optimizer steps, so re-pack from the master weights now
            self._sig = self._weights_signature()
        if self.is_cyc:
            self._sel.copy_(t.plan_replay(static[0].shape[0]), non_blocking=True)   # this step's ReplayBuffer decisions (host RNG)
            self._graphs[0].replay()
            t.last_losses = {"loss_G": self._loss_G, "loss_D_A": self._loss_DA, "loss_D_B": self._loss_DB}
        else:
            self._graphs[0].replay()
            t.last_losses = self._losses
        t.step_count += 1
        return t.last_losses

    def _capture(self, static):
        t = self.t
        cur = torch.cuda.current_stream()
        side = ops.named_stream("graph.warmup")
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(self.warmup):                 # allocator / NCCL / optimizer-state warm-up, off the capture stream
                t.step(tensors=static)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        warm_losses = {k: v.detach().clone() for k, v in t.last_losses.items()}      # the capture below re-binds last_losses to graph outputs
        if not self.is_cyc:
            E.invalidate_weight_cache()
        if self.is_cyc:
            self.refresh_weights()
        n0 = ops.launch_count()
        if self.is_cyc:
            real_A, real_B = static
            B = real_A.shape[0]
            # capture must not consume ReplayBuffer state or host RNG: it runs with "pass through" decisions on the scratch slot
            scratch = t.fake_A_buffer.max_size
            self._sel = torch.full((4, B), scratch, dtype=torch.int64, device=real_A.device)
            assert t.fake_A_buffer.pool is not None and t.fake_B_buffer.pool is not None, "warm-up steps create the device pools"
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._loss_G, self._loss_DA, self._loss_DB = t.phase_fn()(real_A, real_B, self._sel)
            self._graphs = (g,)
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._losses = t.step(tensors=static)
            t.step_count -= 1
            self._graphs = (g,)
        self._launches = ops.launch_count() - n0
        self._sig = self._weights_signature()
        t.last_losses = warm_losses          # what the caller of a replay_first=False first call sees: the eager iteration's losses
