"""Per-stage GPU time (CUDA events on the lane's main stream) and host wall time of one batch, batch mode, one lane.  Dev tool.
    python tools/batch_stage_times.py [batch] [lanes] [side_streams]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from geotransformer_b200.config import make_cfg
from geotransformer_b200.engine import RegistrationEngine
from geotransformer_b200.loss import Evaluator
from geotransformer_b200.model import create_model
from geotransformer_b200.synth import make_pair
from geotransformer_b200.weights import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
LANES = int(sys.argv[2]) if len(sys.argv) > 2 else 1
SIDES = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = make_cfg('3dmatch')
model = create_model(cfg)
model.load_state_dict(synthetic_state_dict(model, 7351))
model = model.cuda().eval()
keys = ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')
pairs = [{k: torch.from_numpy(make_pair('3dmatch20k', i)[k]).cuda() for k in keys} for i in range(B * 6 * LANES)]
eng = RegistrationEngine(model, cfg, cfg.neighbor_limits, num_streams=LANES, evaluator=Evaluator(cfg), batch_size=B, side_streams=SIDES, pin_cpu=True)
eng.register(pairs[:B * 2 * LANES])
torch.cuda.synchronize()
eng.stage_times = {}
t0 = time.perf_counter()
eng.register(pairs[B * 2 * LANES:])
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
n = len(pairs) - B * 2 * LANES
print(f'batch {B}, lanes {LANES}, side streams {SIDES}: {n} pairs in {wall:.1f} ms = {wall / n:.3f} ms/pair = {n / wall * 1e3:.0f} pairs/s')
tot = 0.0
for k, v in eng.stage_times.items():
    print(f'  {k:34s} {np.mean(v):8.3f} ms per batch = {np.mean(v) / B:7.3f} ms/pair ({len(v)} batches)')
    tot += np.mean(v)
print(f'  sum {tot:.3f} ms per batch = {tot / B:.3f} ms/pair (GPU time of the main stream incl. waits for the host at the collate read-backs)')
# host time of the same work with the GPU idle-waiting excluded: time the Python side with launches only (no sync inside a batch
# except the collate read-backs) -> rough upper bound by running with CUDA_LAUNCH_BLOCKING unset and measuring per-batch wall
eng.close()
