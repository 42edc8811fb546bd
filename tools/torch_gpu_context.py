"""the reference step on THIS GPU through torch + cuDNN, including torch.compile (train.py:140-146).
    python tools/torch_gpu_context.py        (prints one JSON object; compile can take minutes)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ste_gan_b200.synthetic import synthetic_batch
host = [synthetic_batch(16, 100, seed=i) for i in range(4)]
print(json.dumps(bench.torch_gpu_context(torch, host, True)))
