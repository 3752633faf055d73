"""TEST INFRASTRUCTURE ONLY -- packs the stage-boundary tensors of one forward (the real reference's, when run by
oracle/make_golden.py in the build container; the pinned restatement oracle/geo_oracle.py's, when a -m gpu parity test runs
it live at BASELINE.json's full sizes) into the flat dict layout of tests/golden/*.npz, so that fixture-based and live
parity tests share one checker (tests/test_gpu_e2e.py::check_forward)."""
import numpy as np


def sample_rows(t, n=64):
    t = t.detach()
    idx = np.unique(np.linspace(0, t.shape[0] - 1, num=min(n, t.shape[0])).astype(np.int64))
    return idx, t[idx].numpy()


def pack(data, taps, out, node_corr_scores, limits):
    """data: collated dict (CPU tensors); taps: geo_oracle.forward taps; out: output_dict the fixture is made of."""
    g = {}
    for i, (p, l) in enumerate(zip(data['points'], data['lengths'])):
        if i > 0:
            g[f'points_{i}'] = p.numpy()
        g[f'lengths_{i}'] = l.numpy()
    for key in ('neighbors', 'subsampling', 'upsampling'):
        for i, t in enumerate(data[key]):
            g[f'{key}_{i}'] = t.numpy().astype(np.uint16 if int(t.max()) < 65536 else np.int32)   # sentinel = number of support rows
    for k in ('feats_c', 'feats_f', 'ref_embeddings'):
        t = taps[k]
        idx, rows = sample_rows(t.reshape(t.shape[0], -1) if k != 'ref_embeddings' else t.reshape(-1, t.shape[-1]), 96)
        g[k + '_rows'], g[k + '_sample'] = idx, rows
        g[k + '_sum'] = np.array([t.double().sum().item(), t.double().abs().sum().item()])
    for k in ('ref_feats_c', 'src_feats_c', 'estimated_transform', 'corr_scores', 'ref_corr_points', 'src_corr_points',
              'ref_node_corr_indices', 'src_node_corr_indices'):
        g[k] = out[k].detach().numpy()
    g['gt_node_corr_indices'] = out['gt_node_corr_indices'].numpy()
    g['gt_node_corr_overlaps'] = out['gt_node_corr_overlaps'].numpy()
    g['node_corr_scores'] = node_corr_scores.numpy()
    idx, rows = sample_rows(out['matching_scores'].reshape(out['matching_scores'].shape[0], -1), 16)
    g['matching_scores_rows'], g['matching_scores_sample'] = idx, rows
    for k in ('ref_node_knn_indices', 'src_node_knn_indices'):
        g[k] = taps[k].numpy().astype(np.int32)
    g['neighbor_limits'] = np.array(limits)
    return g
