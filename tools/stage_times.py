"""Per-stage GPU time of one pair (CUDA events, warm), single stream.  Dev tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from geotransformer_b200 import functional as GF
from geotransformer_b200.config import make_cfg
from geotransformer_b200.model import create_model
from geotransformer_b200.synth import make_pair
from geotransformer_b200.utils.data import registration_collate_fn_stack_mode
from geotransformer_b200.weights import synthetic_state_dict
import time
cfg = make_cfg('3dmatch'); model = create_model(cfg); model.load_state_dict(synthetic_state_dict(model, 7351)); model = model.cuda().eval()
pairs = [make_pair('3dmatch20k', i) for i in range(6)]
dev = [{k: torch.from_numpy(p[k]).cuda() for k in ('ref_points','src_points','ref_feats','src_feats','transform')} for p in pairs]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
marks = []
orig = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        s = ev(); r = f(*a, **k); e = ev(); marks.append((label, s, e)); return r
    setattr(obj, name, g)
wrap(model.backbone, 'forward', 'backbone'); wrap(model.transformer.embedding, 'forward', 'gse(indices+embed)')
wrap(model.transformer.transformer, 'forward_stacked', 'transformer layers'); wrap(model.coarse_matching, 'forward', 'coarse matching')
wrap(model.optimal_transport, 'forward', 'sinkhorn'); wrap(model.fine_matching, 'forward', 'lgr')
for name in ('point_to_node_partition', 'gather_patches', 'patch_scores', 'linear', 'kpconv', 'group_norm', 'maxpool', 'upsample_concat', 'attention', 'head_project', 'add_layernorm', 'gse_indices', 'gse_embed', 'l2_normalize'):
    wrap(GF, name, name)
tot = {}
for it, d in enumerate(dev):
    marks.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0 = ev()
    data = registration_collate_fn_stack_mode([d], 4, 0.025, 0.0625, [38,36,36,38])
    s1 = ev()
    out = model(data)
    s2 = ev(); torch.cuda.synchronize(); wall = (time.perf_counter()-t0)*1e3
    if it >= 2:
        tot.setdefault('collate', []).append(s0.elapsed_time(s1)); tot.setdefault('model total', []).append(s1.elapsed_time(s2)); tot.setdefault('wall', []).append(wall)
        for lab, s, e in marks: tot.setdefault(lab, []).append(s.elapsed_time(e))
for k, v in tot.items():
    n = len(v) / 4
    print(f'{k:24s} {sum(v)/4:8.3f} ms per pair ({n:.0f} calls)')
