"""Throughput vs number of streams and the interpreter's GIL switch interval (dev tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from geotransformer_b200.config import make_cfg
from geotransformer_b200.engine import RegistrationEngine
from geotransformer_b200.loss import Evaluator
from geotransformer_b200.model import create_model
from geotransformer_b200.synth import make_pair
from geotransformer_b200.weights import synthetic_state_dict

cfg = make_cfg('3dmatch')
model = create_model(cfg)
model.load_state_dict(synthetic_state_dict(model, 7351))
model = model.cuda().eval()
limits = [38, 36, 36, 38]
keys = ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')
n = 64
pairs = [{k: torch.from_numpy(make_pair('3dmatch20k', i)[k]).pin_memory() for k in keys} for i in range(n)]
ev = Evaluator(cfg)
for interval in (5e-3, 2e-4, 2e-5):
    sys.setswitchinterval(interval)
    for s in (3, 4, 5, 6, 8):
        eng = RegistrationEngine(model, cfg, limits, num_streams=s, evaluator=ev)
        eng.register(pairs)          # warm: every scratch buffer at its final size
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.register(pairs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'switch interval {interval:g}s streams {s}: {n / dt:7.1f} pairs/s', flush=True)
        eng.close()
