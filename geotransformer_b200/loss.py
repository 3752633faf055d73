"""Evaluator (reference ``experiments/<exp>/loss.py:95-159``): PIR, IR, RRE, RTE, RMSE, RR of one registered pair.

Same constructor/forward contract as the reference's three Evaluator classes (3DMatch, KITTI, ModelNet differ in how
RMSE and RR are defined; ``cfg.name`` selects).  All six numbers come from ONE kernel launch (`geob200_evaluate`); the
result dict holds 0-dim device tensors like the reference's.  KITTI has no RMSE entry (loss.py:140-151 there).
"""
import torch
import torch.nn as nn

from . import functional as GF


class Evaluator(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.mode = GF.EVAL_MODES[cfg.name]
        e = cfg.eval
        self.acceptance_overlap = e.acceptance_overlap
        self.acceptance_radius = e.acceptance_radius
        self.acceptance_rmse = getattr(e, 'rmse_threshold', 0.0) if self.mode == 0 else 0.0
        self.acceptance_rre = e.rre_threshold
        self.acceptance_rte = e.rte_threshold

    @torch.no_grad()
    def metrics_tensor(self, output_dict, data_dict, out=None):
        """(8,) device tensor [PIR, IR, RRE, RTE, RMSE, RR, #corr, #gt_node_corr] -- no host sync."""
        return GF.evaluate(output_dict['gt_node_corr_indices'], output_dict['gt_node_corr_overlaps'],
                           output_dict['ref_node_corr_indices'], output_dict['src_node_corr_indices'],
                           output_dict['ref_corr_points'], output_dict['src_corr_points'], data_dict['transform'],
                           output_dict['estimated_transform'], output_dict['src_points'], self.mode,
                           self.acceptance_overlap, self.acceptance_radius, self.acceptance_rmse, self.acceptance_rre,
                           self.acceptance_rte, out=out)

    def forward(self, output_dict, data_dict):
        m = self.metrics_tensor(output_dict, data_dict)
        res = {'PIR': m[0], 'IR': m[1], 'RRE': m[2], 'RTE': m[3], 'RMSE': m[4], 'RR': m[5]}
        if self.mode == 1:
            del res['RMSE']
        return res
