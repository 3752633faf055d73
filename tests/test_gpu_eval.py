"""Ground-truth superpoint correspondences (matching.py:231-315) and the Evaluator (loss.py:95-159) on the GPU against
the oracle and against the fixtures written by the real reference."""
import math

import numpy as np
import pytest
import torch

from geotransformer_b200 import functional as GF
from geotransformer_b200.config import make_cfg
from geotransformer_b200.loss import Evaluator
from geotransformer_b200.modules.registration import get_node_correspondences
from geotransformer_b200.synth import make_pair
from geotransformer_b200.utils.data import registration_collate_fn_stack_mode
from oracle import geo_oracle as G

pytestmark = pytest.mark.gpu


def _dense(idx, ov, m, n):
    d = torch.zeros(m, n)
    d[idx[:, 0], idx[:, 1]] = ov
    return d


def _compare(got, want, m, n, k):
    """pairs in the same (row-major) order with the same overlaps; a point pair whose squared distance is within float
    noise of pos_radius^2 may flip one count (1/(2K) of overlap): at most 0.5% of the superpoint pairs may do so"""
    gi, go = got[0].cpu(), got[1].cpu()
    wi, wo = want
    assert gi.dtype == torch.int64 and gi.ndim == 2 and gi.shape[1] == 2
    lin = gi[:, 0] * n + gi[:, 1]
    assert bool((lin[1:] > lin[:-1]).all()), 'pairs must come in row-major nonzero order'
    if torch.equal(gi, wi):
        assert float((go - wo).abs().max()) <= 1e-6 if go.numel() else True
        return 0
    diff = (_dense(gi, go, m, n) - _dense(wi, wo, m, n)).abs()
    nbad = int((diff > 1e-6).sum())
    assert nbad <= max(1, (wi.shape[0] + 199) // 200) and float(diff.max()) <= 1.0 / k + 1e-6, (nbad, float(diff.max()))
    return nbad


@pytest.mark.parametrize('m,n,k,masked', [(37, 41, 16, False), (120, 97, 64, True), (1, 1, 8, True), (64, 64, 128, True)])
def test_node_correspondences_vs_oracle(m, n, k, masked):
    g = torch.Generator().manual_seed(m * 1000 + n)
    ref_nodes = torch.rand(m, 3, generator=g) * 2.0
    src_nodes0 = torch.rand(n, 3, generator=g) * 2.0
    ref_knn = ref_nodes[:, None] + 0.25 * torch.randn(m, k, 3, generator=g)
    src_knn0 = src_nodes0[:, None] + 0.25 * torch.randn(n, k, 3, generator=g)
    ang = 0.7
    T = torch.eye(4)
    T[:3, :3] = torch.tensor([[math.cos(ang), -math.sin(ang), 0], [math.sin(ang), math.cos(ang), 0], [0, 0, 1.0]])
    T[:3, 3] = torch.tensor([0.3, -0.2, 0.1])
    Tinv = torch.linalg.inv(T)
    src_nodes, src_knn = G.apply_transform(src_nodes0, Tinv), G.apply_transform(src_knn0, Tinv)
    if masked:
        ref_masks, src_masks = torch.rand(m, generator=g) > 0.1, torch.rand(n, generator=g) > 0.1
        ref_km, src_km = torch.rand(m, k, generator=g) > 0.3, torch.rand(n, k, generator=g) > 0.3
        ref_km[:, 0] = True
        src_km[:, 0] = True
    else:
        ref_masks, src_masks = torch.ones(m, dtype=torch.bool), torch.ones(n, dtype=torch.bool)
        ref_km, src_km = torch.ones(m, k, dtype=torch.bool), torch.ones(n, k, dtype=torch.bool)
    r = 0.12
    want = G.get_node_correspondences(ref_nodes, src_nodes, ref_knn, src_knn, T, r, ref_masks, src_masks, ref_km, src_km)
    c = lambda t: t.cuda()
    if masked:
        got = get_node_correspondences(c(ref_nodes), c(src_nodes), c(ref_knn), c(src_knn), c(T), r, c(ref_masks), c(src_masks),
                                       c(ref_km), c(src_km))
    else:       # masks omitted = all valid (matching.py:268-275)
        got = get_node_correspondences(c(ref_nodes), c(src_nodes), c(ref_knn), c(src_knn), c(T), r)
    assert want[0].shape[0] > 0 or m == 1
    _compare(got, want, m, n, k)


def test_node_correspondences_rejects_bad_arguments():
    z = torch.zeros(4, 3).cuda()
    p = torch.zeros(4, 8, 3).cuda()
    with pytest.raises(RuntimeError):
        get_node_correspondences(z.cpu(), z, p, p, torch.eye(4).cuda(), 0.1)          # no CPU path
    with pytest.raises(ValueError):
        get_node_correspondences(z, z, p, p[:, :4].contiguous(), torch.eye(4).cuda(), 0.1)
    with pytest.raises(ValueError):
        get_node_correspondences(z, z, p, p, torch.eye(4).cuda(), 0.1, ref_masks=torch.ones(5, dtype=torch.bool).cuda())


@pytest.mark.parametrize('workload,cfg_name', [('demo2k', '3dmatch'), ('modelnet717', 'modelnet')])
def test_forward_gt_node_corr_and_metrics_match_reference(workload, cfg_name, golden, models):
    cfg, sd, model = models(cfg_name)
    model = model.cuda().eval()
    gold = golden(workload)
    pair = make_pair(workload, 0)
    dd = {k: pair[k] for k in ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')}
    data = registration_collate_fn_stack_mode([dd], cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                              cfg.backbone.init_radius, gold['neighbor_limits'].tolist())
    for key in ('neighbors', 'subsampling', 'upsampling'):
        data[key] = [torch.from_numpy(gold[f'{key}_{i}'].astype(np.int64)).cuda() for i in range(len(data[key]))]
    data['forced_node_corr'] = (torch.from_numpy(gold['ref_node_corr_indices']).cuda(),
                                torch.from_numpy(gold['src_node_corr_indices']).cuda(),
                                torch.from_numpy(gold['node_corr_scores']).cuda())
    out = model(data)
    m, n = out['ref_points_c'].shape[0], out['src_points_c'].shape[0]
    want = (torch.from_numpy(gold['gt_node_corr_indices']), torch.from_numpy(gold['gt_node_corr_overlaps']))
    nbad = _compare((out['gt_node_corr_indices'], out['gt_node_corr_overlaps']), want, m, n, cfg.model.num_points_in_patch)
    print(f'{workload}: {want[0].shape[0]} gt superpoint pairs, {nbad} differ by one borderline point')

    # the Evaluator on the REFERENCE's outputs must give the reference's numbers, for all three experiment variants
    src_points = torch.from_numpy(pair['src_points']).cuda()
    ref_out = {'gt_node_corr_indices': want[0].cuda(), 'gt_node_corr_overlaps': want[1].cuda(), 'src_points': src_points}
    for k in ('ref_node_corr_indices', 'src_node_corr_indices', 'ref_corr_points', 'src_corr_points', 'estimated_transform'):
        ref_out[k] = torch.from_numpy(gold[k]).cuda()
    for variant in ('3dmatch', 'kitti', 'modelnet'):
        suffix = '' if variant == cfg_name else '_' + variant
        names, vals = gold['metric_names' + suffix].tolist(), gold['metric_values' + suffix]
        res = Evaluator(make_cfg(variant))(ref_out, data)
        assert sorted(res) == names
        for name, v in zip(names, vals):
            tol = 2e-3 if name == 'RRE' else 1e-5         # RRE = acos of an fp32 trace close to 3
            assert abs(float(res[name]) - v) <= tol * max(1.0, abs(v)), (variant, name, float(res[name]), v)
    # ... and on our own outputs it must agree with the oracle's Evaluator fed the same outputs
    ours = Evaluator(cfg)(out, data)
    cpu = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    o = G.evaluate(cfg, cpu, torch.from_numpy(pair['transform']))
    for name in o:
        tol = 2e-3 if name == 'RRE' else 1e-5
        assert abs(float(ours[name]) - float(o[name])) <= tol * max(1.0, abs(float(o[name]))), (name, float(ours[name]), float(o[name]))


def test_evaluator_empty_sets_are_nan():
    cfg = make_cfg('3dmatch')
    e = torch.empty
    out = {'gt_node_corr_indices': e((0, 2), dtype=torch.int64).cuda(), 'gt_node_corr_overlaps': e((0,)).cuda(),
           'ref_node_corr_indices': e((0,), dtype=torch.int64).cuda(), 'src_node_corr_indices': e((0,), dtype=torch.int64).cuda(),
           'ref_corr_points': e((0, 3)).cuda(), 'src_corr_points': e((0, 3)).cuda(), 'estimated_transform': torch.eye(4).cuda(),
           'src_points': torch.rand(100, 3).cuda()}
    res = Evaluator(cfg)(out, {'transform': torch.eye(4).cuda()})
    assert math.isnan(float(res['PIR'])) and math.isnan(float(res['IR']))           # torch: mean of an empty tensor
    assert float(res['RRE']) == 0.0 and float(res['RTE']) == 0.0 and float(res['RMSE']) == 0.0 and float(res['RR']) == 1.0
    m = GF.evaluate(out['gt_node_corr_indices'], out['gt_node_corr_overlaps'], out['ref_node_corr_indices'],
                    out['src_node_corr_indices'], out['ref_corr_points'], out['src_corr_points'], torch.eye(4).cuda(),
                    torch.eye(4).cuda(), out['src_points'], 1, 0.0, 1.0, 0.0, 5.0, 2.0)
    assert math.isnan(float(m[4])) and float(m[5]) == 1.0                             # KITTI: no RMSE


def test_engine_streams_and_metrics(models):
    """RegistrationEngine: several pairs in flight on separate streams give the same results as one pair at a time, in order"""
    from geotransformer_b200.engine import RegistrationEngine
    cfg, sd, model = models('3dmatch')
    model = model.cuda().eval()
    limits = [38, 36, 36, 38]
    pairs = []
    for i in range(5):
        p = make_pair('demo2k', i)
        pairs.append({k: p[k] for k in ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')})
    ev = Evaluator(cfg)
    seq = RegistrationEngine(model, cfg, limits, num_streams=1, evaluator=ev)
    par = RegistrationEngine(model, cfg, limits, num_streams=3, evaluator=ev)
    a = seq.register(pairs, keep_outputs=True)
    b = par.register(pairs)
    assert par.register([]) == []
    seq.close(); par.close()
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.equal(x['estimated_transform'], y['estimated_transform']), i
        assert x['num_corr'] == y['num_corr'] and x['num_superpoints'] == y['num_superpoints']
        for k in x['metrics']:
            assert x['metrics'][k] == y['metrics'][k] or (math.isnan(x['metrics'][k]) and math.isnan(y['metrics'][k]))
        out = x['output_dict']
        cpu = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        o = G.evaluate(cfg, cpu, torch.from_numpy(pairs[i]['transform']))
        for name in o:
            tol = 2e-3 if name == 'RRE' else 1e-5
            assert abs(x['metrics'][name] - float(o[name])) <= tol * max(1.0, abs(float(o[name]))), (i, name)
    assert len({tuple(x['estimated_transform'].flatten().tolist()) for x in a}) == 5        # five different pairs


@pytest.mark.parametrize('batch_size', [1, 2])
def test_tester_loop_writes_reference_npz(models, tmp_path, batch_size):
    """SingleTester-style loop: one <scene>/<ref>_<src>.npz per pair with the arrays test.py:73-92 writes, metrics summary;
    batch_size 2 = two pairs per forward (3 pairs: one full batch + a trailing single pair)"""
    from geotransformer_b200.tester import RegistrationTester, NPZ_OUTPUT_KEYS
    cfg, sd, model = models('3dmatch')
    model = model.cuda().eval()
    dataset = []
    for i in range(3):
        p = make_pair('demo2k', i)
        d = {k: p[k] for k in ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')}
        d.update(scene_name='synthetic_scene', ref_frame=2 * i, src_frame=2 * i + 1, overlap=0.5)
        dataset.append(d)
    lines = []
    tester = RegistrationTester(cfg, model, [38, 36, 36, 38], output_dir=str(tmp_path), num_streams=2 if batch_size == 1 else 1, chunk=2,
                                batch_size=batch_size)
    summary, per_pair = tester.run(dataset, log=lines.append)
    tester.close()
    assert len(per_pair) == 3 and len(lines) == 3 and set(summary) == {'PIR', 'IR', 'RRE', 'RTE', 'RMSE', 'RR'}
    for i, entry in enumerate(per_pair):
        z = np.load(entry['file'])
        assert entry['file'].endswith(f'synthetic_scene/{2 * i}_{2 * i + 1}.npz')
        assert set(z.files) == set(NPZ_OUTPUT_KEYS) | {'transform', 'overlap'}
        assert np.array_equal(z['estimated_transform'], entry['estimated_transform'].numpy())
        assert np.array_equal(z['transform'], dataset[i]['transform'])
        assert z['gt_node_corr_indices'].shape[1] == 2 and z['ref_corr_points'].shape[0] == entry['num_corr']
    assert abs(summary['RRE'] - np.mean([p['metrics']['RRE'] for p in per_pair])) < 1e-6
