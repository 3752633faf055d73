"""Small profiling target: a few batches of the registration hot path (batch mode) with a modest memory footprint, so that
ncu's save/restore between replay passes stays cheap.  Dev tool.
    ncu <options> python tools/ncu_target.py [batch] [batches] [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from geotransformer_b200.config import make_cfg
from geotransformer_b200.engine import RegistrationEngine
from geotransformer_b200.loss import Evaluator
from geotransformer_b200.model import create_model
from geotransformer_b200.synth import WORKLOADS, make_pair
from geotransformer_b200.weights import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 3
workload = sys.argv[3] if len(sys.argv) > 3 else '3dmatch20k'
cfg = make_cfg(WORKLOADS[workload][0])
limits = cfg.neighbor_limits or [27, 75, 147, 157, 119][:cfg.backbone.num_stages]
model = create_model(cfg)
model.load_state_dict(synthetic_state_dict(model, 7351))
model = model.cuda().eval()
keys = ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')
pairs = [{k: torch.from_numpy(make_pair(workload, i)[k]).cuda() for k in keys} for i in range(B * NB)]
eng = RegistrationEngine(model, cfg, limits, num_streams=1, evaluator=Evaluator(cfg), batch_size=B, side_streams=2)
res = eng.register(pairs)
torch.cuda.synchronize()
eng.close()
print('registered', len(res), 'pairs; last RRE', res[-1]['metrics']['RRE'])
