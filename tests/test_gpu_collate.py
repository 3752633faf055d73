"""GPU collate kernels vs the plain-C oracle: values AND order bit-exact for grid_subsample; radius_search identical
including order on continuous data, identical up to exact-tie order otherwise (see oracle/collate_oracle.c header)."""
import numpy as np
import pytest
import torch

from geotransformer_b200.synth import make_pair
from oracle import collate_oracle as co
from oracle import geo_oracle

pytestmark = pytest.mark.gpu


def _stack(pair):
    pts = torch.from_numpy(np.concatenate([pair['ref_points'], pair['src_points']]))
    lens = torch.tensor([len(pair['ref_points']), len(pair['src_points'])])
    return pts, lens


@pytest.mark.parametrize('workload,voxel', [('demo2k', 0.05), ('modelnet717', 0.1), ('3dmatch20k', 0.05), ('kitti60k', 0.6)])
def test_grid_subsample_pyramid_bit_exact(workload, voxel):
    from geotransformer_b200.modules.ops import grid_subsample
    pts, lens = _stack(make_pair(workload, 0))
    gp, gl = pts.cuda(), lens
    for lvl in range(3):
        op, ol = co.grid_subsampling(pts, lens, voxel)
        sp, sl = grid_subsample(gp, gl, voxel)
        assert torch.equal(sl.cpu(), ol), f'level {lvl}: lengths {sl.tolist()} vs {ol.tolist()}'
        assert sp.shape == op.shape
        assert torch.equal(sp.cpu(), op), f'level {lvl}: {(sp.cpu() != op).any(dim=1).sum().item()} rows differ (values or order)'
        pts, lens, gp, gl, voxel = op, ol, sp, sl.cpu(), voxel * 2


def test_grid_subsample_many_clouds_and_tiny():
    from geotransformer_b200.modules.ops import grid_subsample
    g = torch.Generator().manual_seed(3)
    sizes = [1, 2, 13, 14, 15, 29, 30, 31, 500, 1, 3000]
    pts = torch.rand(sum(sizes), 3, generator=g) * 0.7 - 0.2
    lens = torch.tensor(sizes)
    for voxel in (0.03, 0.11, 5.0):
        op, ol = co.grid_subsampling(pts, lens, voxel)
        sp, sl = grid_subsample(pts.cuda(), lens, voxel)
        assert torch.equal(sl.cpu(), ol)
        assert torch.equal(sp.cpu(), op)


def test_grid_subsample_rejects_bad_input():
    from geotransformer_b200.modules.ops import grid_subsample
    pts = torch.rand(10, 3).cuda()
    with pytest.raises(RuntimeError):
        grid_subsample(pts, torch.tensor([10, 0]), 0.1)          # empty cloud (the reference would read points[0])
    with pytest.raises(RuntimeError):
        grid_subsample(pts.double(), torch.tensor([10]), 0.1)    # dtype check of the reference (CHECK_IS_FLOAT)
    with pytest.raises(RuntimeError):
        grid_subsample(pts, torch.tensor([10], dtype=torch.int32), 0.1)


@pytest.mark.parametrize('workload', ['demo2k', 'modelnet717'])
def test_radius_search_all_tables(workload):
    """all 3S-2 searches of the collate (self / sub / up) against the oracle, full reference width and limited width."""
    from geotransformer_b200.modules.ops import radius_search
    from geotransformer_b200 import ext
    from geotransformer_b200.config import make_cfg
    pair = make_pair(workload, 0)
    cfg = make_cfg(pair['config'])
    S = cfg.backbone.num_stages
    limits = cfg.neighbor_limits or [13, 21, 27]
    data = geo_oracle.collate_pair(pair, cfg, limits)
    r = cfg.backbone.init_radius
    for i in range(S):
        p, l = data['points'][i], data['lengths'][i]
        cases = [('self', p, p, l, l, r, limits[i], data['neighbors'][i])]
        if i < S - 1:
            sp, sl = data['points'][i + 1], data['lengths'][i + 1]
            cases.append(('sub', sp, p, sl, l, r, limits[i], data['subsampling'][i]))
            cases.append(('up', p, sp, l, sl, r * 2, limits[i + 1], data['upsampling'][i]))
        for name, q, s, ql, sl_, rad, lim, want in cases:
            got = radius_search(q.cuda(), s.cuda(), ql, sl_, rad, lim)
            assert got.is_contiguous() and got.dtype == torch.int64
            assert got.shape == want.shape, f'{name}[{i}] shape {tuple(got.shape)} vs {tuple(want.shape)}'
            if name == 'self':
                assert torch.equal(got.cpu(), want), f'{name}[{i}]: {(got.cpu() != want).any(dim=1).sum().item()} rows differ'
            else:   # exact ties exist (barycentre of a 2-point voxel): both sides break them by index here
                assert torch.equal(got.cpu(), want), f'{name}[{i}] differs'
            if i == S - 1:      # full-width reference call signature
                full = ext.radius_neighbors(q.cuda(), s.cuda(), ql, sl_, rad)
                want_full = co.radius_neighbors(q, s, ql, sl_, rad)
                assert torch.equal(full.cpu(), want_full)
        r *= 2


def test_radius_search_dense_overflow_path():
    """more than 256 neighbours per query exercises the large-capacity second pass"""
    from geotransformer_b200 import ext
    g = torch.Generator().manual_seed(5)
    pts = torch.rand(1500, 3, generator=g) * 0.2
    lens = torch.tensor([900, 600])
    got = ext.radius_neighbors(pts.cuda(), pts.cuda(), lens, lens, 0.12)
    want = co.radius_neighbors(pts, pts, lens, lens, 0.12)
    assert got.shape == want.shape and got.shape[1] > 256
    assert torch.equal(got.cpu(), want)


def test_radius_search_query_outside_support_box():
    from geotransformer_b200 import ext
    g = torch.Generator().manual_seed(6)
    s = torch.rand(300, 3, generator=g)
    q = torch.cat([torch.rand(50, 3, generator=g), torch.rand(50, 3, generator=g) + 5.0, torch.rand(20, 3, generator=g) - 0.1])
    ql, sl = torch.tensor([70, 50]), torch.tensor([200, 100])
    got = ext.radius_neighbors(q.cuda(), s.cuda(), ql, sl, 0.3)
    want = co.radius_neighbors(q, s, ql, sl, 0.3)
    assert torch.equal(got.cpu(), want)


def test_cpu_tensor_drop_in_roundtrip():
    """literal drop-in of the reference's CPU call sites: CPU tensors in, CPU tensors out, computed on the GPU"""
    from geotransformer_b200 import ext
    g = torch.Generator().manual_seed(7)
    pts = torch.rand(400, 3, generator=g)
    lens = torch.tensor([250, 150])
    sp, sl = ext.grid_subsampling(pts, lens, 0.1)
    op, ol = co.grid_subsampling(pts, lens, 0.1)
    assert not sp.is_cuda and torch.equal(sp, op) and torch.equal(sl, ol)


def test_calibrate_neighbors_matches_reference(golden, models):
    """GPU collate at the histogram width + device histogram == the limits computed by the real reference (fixture)"""
    from geotransformer_b200.utils.data import calibrate_neighbors_stack_mode, registration_collate_fn_stack_mode
    cfg, _, _ = models('3dmatch')
    b = cfg.backbone
    keys = ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')
    pairs = [{k: make_pair('demo2k', i)[k] for k in keys} for i in range(3)]
    gold = golden('calibration')
    for thr in (2000, 150):
        got = calibrate_neighbors_stack_mode(pairs, registration_collate_fn_stack_mode, b.num_stages, b.init_voxel_size, b.init_radius,
                                             sample_threshold=thr)
        assert np.array_equal(np.asarray(got), gold[f'limits_threshold_{thr}']), (got, gold[f'limits_threshold_{thr}'])
