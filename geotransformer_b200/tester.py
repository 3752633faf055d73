"""Test-time loop over a dataset of pairs.

Reference: ``geotransformer/engine/single_tester.py:39-74`` (the loop: to_cuda -> test_step -> eval_step -> after_test_step ->
summary) and ``experiments/*/test.py:40-92`` (test_step = model forward, eval_step = Evaluator, after_test_step = one
``<output_dir>/<scene_name>/<ref_frame>_<src_frame>.npz`` per pair with the arrays ``eval.py`` consumes).  Differences: the
collate runs on the GPU in this process (CUDA cannot be used in forked DataLoader workers), ``num_streams`` pairs are in
flight at once (`RegistrationEngine`), and the six metrics come from one device kernel per pair.
"""
import os

import numpy as np
import torch

from .engine import RegistrationEngine
from .loss import Evaluator

# arrays written per pair, in the order of experiments/geotransformer.3dmatch.*/test.py:73-92
NPZ_OUTPUT_KEYS = ('ref_points', 'src_points', 'ref_points_f', 'src_points_f', 'ref_points_c', 'src_points_c', 'ref_feats_c',
                   'src_feats_c', 'ref_node_corr_indices', 'src_node_corr_indices', 'ref_corr_points', 'src_corr_points',
                   'corr_scores', 'gt_node_corr_indices', 'gt_node_corr_overlaps', 'estimated_transform')
METRICS = ('PIR', 'IR', 'RRE', 'RTE', 'RMSE', 'RR')


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class RegistrationTester:
    def __init__(self, cfg, model, neighbor_limits, output_dir=None, num_streams=4, chunk=16, device=None, batch_size=1):
        """batch_size > 1: that many pairs per forward (GeoTransformer.forward_batch) on each of the num_streams lanes"""
        self.cfg, self.output_dir, self.chunk = cfg, output_dir, max(1, int(chunk), int(batch_size) * int(num_streams))
        self.engine = RegistrationEngine(model, cfg, neighbor_limits, num_streams=num_streams, device=device, evaluator=Evaluator(cfg),
                                         batch_size=batch_size)

    def after_test_step(self, data_dict, output_dict):
        """experiments/*/test.py:65-92"""
        scene, ref_id, src_id = data_dict.get('scene_name', 'scene'), data_dict.get('ref_frame', 0), data_dict.get('src_frame', 1)
        os.makedirs(os.path.join(self.output_dir, str(scene)), exist_ok=True)
        arrays = {k: _np(output_dict[k]) for k in NPZ_OUTPUT_KEYS}
        arrays['transform'] = _np(data_dict['transform'])
        arrays['overlap'] = data_dict.get('overlap', np.float32('nan'))
        path = os.path.join(self.output_dir, str(scene), f'{ref_id}_{src_id}.npz')
        np.savez_compressed(path, **arrays)
        return path

    def run(self, dataset, log=None):
        """dataset: indexable of dicts with ref_points/src_points/ref_feats/src_feats/transform (+ scene_name, ref_frame,
        src_frame, overlap as the reference datasets provide).  Returns (summary of mean metrics, per-pair results)."""
        per_pair = []
        n = len(dataset)
        for base in range(0, n, self.chunk):
            items = [dataset[i] for i in range(base, min(n, base + self.chunk))]
            tensors = [{k: v for k, v in it.items() if isinstance(v, (np.ndarray, torch.Tensor))} for it in items]
            results = self.engine.register(tensors, keep_outputs=self.output_dir is not None)
            for it, res in zip(items, results):
                entry = {'metrics': res['metrics'], 'num_corr': res['num_corr'], 'estimated_transform': res['estimated_transform']}
                if self.output_dir is not None:
                    entry['file'] = self.after_test_step(it, res.pop('output_dict'))
                per_pair.append(entry)
                if log is not None:       # single_tester.py:62-66 / test.py:55-63 summary string
                    msg = ', '.join(f'{k}: {res["metrics"][k]:.3f}' for k in METRICS)
                    log(f"{it.get('scene_name', 'scene')}, id0: {it.get('ref_frame', 0)}, id1: {it.get('src_frame', 1)}, {msg}, "
                        f"nCorr: {res['num_corr']}")
        summary = {k: float(np.mean([p['metrics'][k] for p in per_pair])) for k in METRICS} if per_pair else {}
        return summary, per_pair

    def close(self):
        self.engine.close()
