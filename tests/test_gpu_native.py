"""The C++ stage drivers (native.cu) must reproduce the per-op module path bit for bit: same kernels, same order."""
import pytest
import torch

from geotransformer_b200.model import enable_native
from geotransformer_b200.synth import make_pair
from geotransformer_b200.utils.data import registration_collate_fn_stack_mode

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('workload,cfg_name,limits', [('demo2k', '3dmatch', [38, 36, 36, 38]), ('modelnet717', 'modelnet', [13, 21, 27])])
def test_native_forward_is_bitwise_identical(workload, cfg_name, limits, models):
    cfg, sd, model = models(cfg_name)
    model = model.cuda().eval()
    if hasattr(model, '_native'):
        del model._native
    pair = make_pair(workload, 0)
    dd = {k: pair[k] for k in ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')}
    b = cfg.backbone
    data = registration_collate_fn_stack_mode([dd], b.num_stages, b.init_voxel_size, b.init_radius, limits)
    taps0, taps1 = {}, {}
    out0 = model(data, taps=taps0)
    enable_native(model)
    try:
        out1 = model(data, taps=taps1)
    finally:
        del model._native
    assert torch.equal(taps0['feats_c'], taps1['feats_c']) and torch.equal(taps0['feats_f'], taps1['feats_f'])
    for k in ('ref_feats_c', 'src_feats_c', 'matching_scores', 'ref_corr_points', 'src_corr_points', 'corr_scores', 'estimated_transform',
              'ref_node_corr_indices', 'src_node_corr_indices'):
        assert torch.equal(out0[k], out1[k]), k


def test_native_kitti_backbone(models):
    """5 stages, 3 decoders (decoder4, decoder3 Unary + decoder2 Last)"""
    from geotransformer_b200.synth import WORKLOADS, _ground
    cfg, sd, model = models('kitti')
    model = model.cuda().eval()
    WORKLOADS['kitti6k'] = ('kitti', _ground, dict(n=6000, R=22.0, sigma=0.05), 10.0)
    pair = make_pair('kitti6k', 0)
    dd = {k: pair[k] for k in ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')}
    data = registration_collate_fn_stack_mode([dd], 5, 0.3, 1.275, [27, 75, 147, 157, 119])
    want = model.backbone(data['features'], data)
    enable_native(model)
    try:
        got = model._native.backbone_forward(data['features'], data)
    finally:
        del model._native
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        assert torch.equal(a, b)
