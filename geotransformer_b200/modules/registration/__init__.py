"""reference ``geotransformer/modules/registration/procrustes.py:6-91``."""
import torch
import torch.nn as nn

from ... import functional as GF


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-5, return_transform=False):
    squeeze_first = src_points.ndim == 2
    if squeeze_first:
        src_points, ref_points = src_points.unsqueeze(0), ref_points.unsqueeze(0)
        if weights is not None:
            weights = weights.unsqueeze(0)
    T = GF.weighted_procrustes(src_points.contiguous(), ref_points.contiguous(),
                               None if weights is None else weights.contiguous(), weight_thresh, eps)
    if return_transform:
        return T.squeeze(0) if squeeze_first else T
    R, t = T[:, :3, :3], T[:, :3, 3]
    return (R.squeeze(0), t.squeeze(0)) if squeeze_first else (R, t)


class WeightedProcrustes(nn.Module):
    def __init__(self, weight_thresh=0.0, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh, self.eps, self.return_transform = weight_thresh, eps, return_transform

    def forward(self, src_points, tgt_points, weights=None):
        return weighted_procrustes(src_points, tgt_points, weights=weights, weight_thresh=self.weight_thresh, eps=self.eps,
                                   return_transform=self.return_transform)


@torch.no_grad()
def get_node_correspondences(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks=None,
                             src_masks=None, ref_knn_masks=None, src_knn_masks=None):
    """reference ``geotransformer/modules/registration/matching.py:231-315``: ``(corr_indices (C,2), corr_overlaps (C,))``."""
    res = GF.node_correspondences(ref_nodes.contiguous(), src_nodes.contiguous(), ref_knn_points.contiguous(),
                                  src_knn_points.contiguous(), transform.contiguous(), pos_radius, ref_masks, src_masks,
                                  ref_knn_masks, src_knn_masks)
    return GF.finish_node_correspondences(*res)
