"""pruner.prune() wall seconds on a full-size random-init BLIP-2 (the first item of BASELINE.json's metric).
usage: prune_wall.py [wanda] [ecoflap] [sparsegpt]     (torchrun for N > 1: stage 1 is sharded over layers)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import prune_wall  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1:] or ["wanda", "ecoflap"]
    print(json.dumps(prune_wall(which, verbose=True)))
