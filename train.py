#!/usr/bin/python3
"""python train.py --config Yaml/CycleGan.yaml [--mode train|test]   (under torchrun for multi-GPU data parallel).

Same dispatch as the reference's train.py:31-45 (config['name'] picks the trainer), seed 42 as train.py:22-28,49.
`--mode` replaces the reference's "edit the file to toggle train()/test()" (train.py:44-45); `name: RegGan` reaches
Reg_Trainer, which the reference imports but never dispatches to."""
import argparse
import os
import random

import numpy as np
import torch

from trainer import Cyc_Trainer, Hd_Trainer_x1, Hd_Trainer_x2, P2p_Trainer, Reg_Trainer
from trainer.utils import get_config


def seed_everything(seed):
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="Yaml/HdGan.yaml", help="Path to the config file.")
    parser.add_argument("--mode", choices=["train", "test"], default="train")
    opts = parser.parse_args()
    config = get_config(opts.config)
    trainers = {"CycleGan": Cyc_Trainer, "P2p": P2p_Trainer, "RegGan": Reg_Trainer,
                "HdGan": Hd_Trainer_x2 if config.get("stage", 1) == 2 else Hd_Trainer_x1}
    trainer = trainers[config["name"]](config)
    getattr(trainer, opts.mode)()


if __name__ == "__main__":
    seed_everything(seed=42)
    main()
