"""The oracle itself (CPU): plain-C collate restatement vs the real reference build (oracle/_ref), and the
torch restatement of the forward vs the committed reference fixtures."""
import numpy as np
import pytest
import torch

from geotransformer_b200.synth import make_pair
from oracle import collate_oracle as co
from oracle import geo_oracle as G
from oracle import ref_ext


def _stack(pair):
    pts = torch.from_numpy(np.concatenate([pair['ref_points'], pair['src_points']]))
    return pts, torch.tensor([len(pair['ref_points']), len(pair['src_points'])])


@pytest.mark.skipif(not ref_ext.available(), reason='oracle/_ref not built')
@pytest.mark.parametrize('workload,voxel', [('demo2k', 0.05), ('3dmatch20k', 0.05), ('modelnet717', 0.1)])
def test_c_restatement_matches_reference_build(workload, voxel):
    pts, lens = _stack(make_pair(workload, 1))
    for _ in range(3):
        a, al = ref_ext.grid_subsampling(pts, lens, voxel)
        b, bl = co.grid_subsampling(pts, lens, voxel)
        assert torch.equal(al, bl) and torch.equal(a, b)          # values AND unordered_map order
        pts, lens, voxel = a, al, voxel * 2
    na = ref_ext.radius_neighbors(pts, pts, lens, lens, voxel * 1.25)
    nb = co.radius_neighbors(pts, pts, lens, lens, voxel * 1.25)
    assert torch.equal(na, nb)


def test_grid_subsample_golden_order(golden):
    """against the committed fixture (reference output), no reference build needed"""
    gold = golden('demo2k')
    pts, lens = _stack(make_pair('demo2k', 0))
    v = 0.05
    for i in range(1, 4):
        pts, lens = co.grid_subsampling(pts, lens, v)
        assert np.array_equal(pts.numpy(), gold[f'points_{i}']) and lens.tolist() == gold[f'lengths_{i}'].tolist()
        v *= 2


def test_rehash_schedule_edge_sizes():
    """clouds whose voxel counts straddle the libstdc++ rehash thresholds (13/14, 29/30, 59/60, 127/128)"""
    if not ref_ext.available():
        pytest.skip('oracle/_ref not built')
    g = torch.Generator().manual_seed(0)
    for n in (1, 2, 12, 13, 14, 28, 29, 30, 58, 59, 60, 126, 127, 128, 129, 257, 258, 542):
        pts = torch.rand(n, 3, generator=g) * 100.0          # one point per voxel with high probability
        a, al = ref_ext.grid_subsampling(pts, torch.tensor([n]), 0.5)
        b, bl = co.grid_subsampling(pts, torch.tensor([n]), 0.5)
        assert torch.equal(a, b), n


def test_forward_restatement_matches_reference_fixture(golden, models):
    """torch restatement vs the real reference on the ModelNet-shape pair (teacher-forced neighbour tables)"""
    cfg, sd, _ = models('modelnet')
    gold = golden('modelnet717')
    pair = make_pair('modelnet717', 0)
    data = G.collate_pair(pair, cfg, gold['neighbor_limits'].tolist())
    for i in range(1, cfg.backbone.num_stages):
        assert np.array_equal(data['points'][i].numpy(), gold[f'points_{i}'])
    for key in ('neighbors', 'subsampling', 'upsampling'):
        for i in range(len(data[key])):
            want = torch.from_numpy(gold[f'{key}_{i}'].astype(np.int64))
            q = data['points'][i + (1 if key == 'subsampling' else 0)]
            s = data['points'][i + (1 if key == 'upsampling' else 0)]
            assert torch.equal(G.canonical_neighbors(q, s, data[key][i]), G.canonical_neighbors(q, s, want))
            data[key][i] = want
    with torch.no_grad():
        out = G.forward(sd, cfg, data)
    assert np.abs(out['ref_feats_c'].numpy() - gold['ref_feats_c']).max() < 1e-5
    assert np.array_equal(out['ref_node_corr_indices'].numpy(), gold['ref_node_corr_indices'])
    assert np.array_equal(out['ref_corr_points'].numpy(), gold['ref_corr_points'])
    assert np.abs(out['estimated_transform'].numpy() - gold['estimated_transform']).max() < 1e-5
    # ground-truth superpoint correspondences (matching.py:231-315) and the Evaluator (loss.py:95-159)
    assert np.array_equal(out['gt_node_corr_indices'].numpy(), gold['gt_node_corr_indices'])
    assert np.abs(out['gt_node_corr_overlaps'].numpy() - gold['gt_node_corr_overlaps']).max() < 1e-6
    metrics = G.evaluate(cfg, out, data['transform'])
    assert sorted(metrics) == gold['metric_names'].tolist()
    for name, v in zip(gold['metric_names'].tolist(), gold['metric_values']):
        assert abs(float(metrics[name]) - v) < (1e-3 if name == 'RRE' else 1e-5), name


def test_evaluator_variants_match_reference_fixture(golden, models):
    """the three Evaluator variants (3DMatch / KITTI / ModelNet, loss.py:95-159) of the restatement on the demo pair vs the
    numbers the real reference Evaluators produced on the same outputs"""
    from geotransformer_b200.config import make_cfg
    cfg, sd, _ = models('3dmatch')
    gold = golden('demo2k')
    pair = make_pair('demo2k', 0)
    data = G.collate_pair(pair, cfg, gold['neighbor_limits'].tolist())
    for key in ('neighbors', 'subsampling', 'upsampling'):
        for i in range(len(data[key])):
            data[key][i] = torch.from_numpy(gold[f'{key}_{i}'].astype(np.int64))      # the reference's exact-tie order
    with torch.no_grad():
        out = G.forward(sd, cfg, data)
    assert np.array_equal(out['gt_node_corr_indices'].numpy(), gold['gt_node_corr_indices'])
    for variant, suffix in (('3dmatch', ''), ('kitti', '_kitti'), ('modelnet', '_modelnet')):
        metrics = G.evaluate(make_cfg(variant), out, data['transform'])
        assert sorted(metrics) == gold['metric_names' + suffix].tolist()
        for name, v in zip(gold['metric_names' + suffix].tolist(), gold['metric_values' + suffix]):
            assert abs(float(metrics[name]) - v) < (1e-3 if name == 'RRE' else 1e-5), (variant, name)


def test_calibration_restatement_matches_reference_fixture(golden, models):
    """calibrate_neighbors_stack_mode (utils/data.py:190-217): restatement vs the limits the real reference computed"""
    cfg, _, _ = models('3dmatch')
    keys = ('ref_points', 'src_points', 'ref_feats', 'src_feats', 'transform')
    pairs = [{k: make_pair('demo2k', i)[k] for k in keys} for i in range(3)]
    gold = golden('calibration')
    for thr in (2000, 150):
        assert np.array_equal(G.calibrate_neighbors(pairs, cfg, sample_threshold=thr), gold[f'limits_threshold_{thr}'])


def test_sinkhorn_marginals_property():
    g = torch.Generator().manual_seed(1)
    s = torch.randn(3, 20, 20, generator=g)
    rm, cm = torch.rand(3, 20, generator=g) > 0.2, torch.rand(3, 20, generator=g) > 0.2
    out = G.optimal_transport(torch.tensor(1.0), s, rm, cm, 100).exp()
    rows = out[:, :-1, :].sum(dim=2)
    assert (rows[rm] - 1).abs().max() < 1e-3


def test_procrustes_degenerate_cases():
    src = torch.randn(1, 10, 3)
    T = G.weighted_procrustes(src, src + 1.0, torch.zeros(1, 10))
    assert torch.equal(T[0], torch.eye(4))


@pytest.mark.parametrize('c,n,sigma_d,extent', [(256, 100, 0.2, 2.0), (128, 90, 4.8, 20.0)])
def test_tabulated_structure_embedding_model_vs_oracle(c, n, sigma_d, extent):
    """oracle/gse_table_model.py (the arithmetic of csrc/gse_table.cu: fp64-built table of proj(sinusoid(x)) on a 1/256 grid,
    fp16 forward differences, fp32 lerp) reproduces GeometricStructureEmbedding.forward to ~3e-6: the tabulation error is an
    order of magnitude below the 2e-5 of the split-precision tensor-core kernels it replaces"""
    import math
    from oracle.gse_table_model import structure_embedding_tabulated
    g = torch.Generator().manual_seed(n)
    pts = torch.rand(n, 3, generator=g) * extent
    sd = {'e.embedding.div_term': torch.exp(torch.arange(0, c, 2).float() * (-np.log(10000.0) / c)),
          'e.proj_d.weight': torch.randn(c, c, generator=g) / math.sqrt(c), 'e.proj_d.bias': torch.randn(c, generator=g) * 0.1,
          'e.proj_a.weight': torch.randn(c, c, generator=g) / math.sqrt(c), 'e.proj_a.bias': torch.randn(c, generator=g) * 0.1}
    want = G.structure_embedding(sd, 'e.', pts, sigma_d, 15, 3)
    got = structure_embedding_tabulated(sd, 'e.', pts, sigma_d, 15, 3)
    assert float((got - want).abs().max()) < 1e-5
    coarse = structure_embedding_tabulated(sd, 'e.', pts, sigma_d, 15, 3, inv_step=64)
    assert float((coarse - want).abs().max()) < 5e-5


@pytest.mark.skipif(not ref_ext.available(), reason='oracle/_ref not built')
def test_c_restatement_matches_reference_build_on_adversarial_small_inputs():
    """hypothesis: several ragged clouds per batch, coordinates quantised to a coarse lattice (many points per voxel, exact-distance
    ties, duplicated points), negative coordinates, 1-point clouds.  grid_subsampling: values AND unordered_map order bit for bit;
    radius_neighbors: identical rows up to the order inside exact-distance tie groups (std::sort is unstable in both)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.lists(st.integers(1, 40), min_size=1, max_size=4), st.integers(0, 2 ** 31 - 1), st.sampled_from([0.0, 0.05, 0.25]),
           st.sampled_from([0.3, 0.5, 1.0]))
    def check(lengths, seed, lattice, voxel):
        g = torch.Generator().manual_seed(seed)
        n = sum(lengths)
        pts = (torch.rand(n, 3, generator=g) - 0.5) * 4.0
        if lattice > 0:
            pts = torch.round(pts / lattice) * lattice
        pts = pts.contiguous()
        lens = torch.tensor(lengths)
        a, al = ref_ext.grid_subsampling(pts, lens, voxel)
        b, bl = co.grid_subsampling(pts, lens, voxel)
        assert torch.equal(al, bl) and torch.equal(a, b)
        r = voxel * 1.5
        na = ref_ext.radius_neighbors(pts, pts, lens, lens, r)
        nb = co.radius_neighbors(pts, pts, lens, lens, r)
        assert na.shape == nb.shape
        assert torch.equal(G.canonical_neighbors(pts, pts, na), G.canonical_neighbors(pts, pts, nb))
        # sub-sampled queries against the full support (the 'subsampling' tables of the collate)
        nq = ref_ext.radius_neighbors(a, pts, al, lens, r)
        nr = co.radius_neighbors(a, pts, al, lens, r)
        assert nq.shape == nr.shape
        assert torch.equal(G.canonical_neighbors(a, pts, nq), G.canonical_neighbors(a, pts, nr))

    check()


def test_boundary_2_op_restatements_match_the_real_reference_live():
    """oracle/pin_ops_live.py in its own process (it imports the unmodified reference package and patches Tensor.cuda): the oracle's
    pairwise_distance / knn_partition / get_point_to_node_indices / point_to_node_partition / ball_query_partition / apply_transform
    are bit-identical to geotransformer.modules.ops on seeded inputs.  Only where /root/reference exists (the build container)."""
    import os
    import subprocess
    import sys
    from oracle import ref_harness
    if not (ref_harness.available() and ref_ext.available()):
        pytest.skip('needs /root/reference and oracle/_ref (build container only)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'oracle.pin_ops_live'], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'pinned:' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_golden_fixtures_regenerate_bit_identically_from_the_real_reference(tmp_path):
    """python -m oracle.make_golden (the REAL reference package imported from /root/reference, CPU) into a scratch directory gives,
    array for array, ALL committed fixtures tests/golden/{demo2k,modelnet717,kitti4k,calibration}.npz -- and the script itself asserts that the restatement
    oracle/geo_oracle.py equals the reference run.  Only in the build container (the GPU box has no /root/reference)."""
    import os
    import subprocess
    import sys
    from oracle import ref_harness
    if not (ref_harness.available() and ref_ext.available()):
        pytest.skip('needs /root/reference and oracle/_ref (build container only)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GEOB200_GOLDEN_OUT=str(tmp_path))
    names = ('demo2k', 'modelnet717', 'kitti4k', 'calibration')
    r = subprocess.run([sys.executable, '-m', 'oracle.make_golden'] + list(names), cwd=root, env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for w in names:
        new, old = np.load(tmp_path / f'{w}.npz'), np.load(os.path.join(root, 'tests', 'golden', f'{w}.npz'))
        assert sorted(new.files) == sorted(old.files)
        for k in new.files:
            a, b = new[k], old[k]
            assert a.shape == b.shape and a.dtype == b.dtype, (w, k)
            assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == 'f' else np.array_equal(a, b), (w, k)


def test_restatement_matches_the_real_reference_on_other_pairs_live():
    """the committed fixtures pin the restatement on pair 0 of each workload; here the real reference and oracle/geo_oracle.py run
    LIVE on further seeded pairs of all three models (collate tables up to tie order, features, matching scores, correspondences
    as a set when two coarse scores tie to 1e-5, transform, gt superpoint pairs, all three Evaluator variants) -- make_golden's
    own assertions, nothing written.  Build container only."""
    import os
    import subprocess
    import sys
    from oracle import ref_harness
    if not (ref_harness.available() and ref_ext.available()):
        pytest.skip('needs /root/reference and oracle/_ref (build container only)')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cases = ['check:demo2k:1', 'check:demo2k:2', 'check:modelnet717:1', 'check:modelnet717:3', 'check:kitti4k:1']
    r = subprocess.run([sys.executable, '-m', 'oracle.make_golden'] + cases, cwd=root, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and r.stdout.count('restatement == reference') == len(cases), r.stdout[-3000:] + r.stderr[-3000:]
