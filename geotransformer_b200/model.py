"""GeoTransformer registration model (inference forward) on the B200 path.

Reference: ``experiments/*/model.py:18-217``.  Same attribute names (``backbone``, ``transformer``,
``coarse_matching``, ``fine_matching``, ``optimal_transport``) and hence the same ``state_dict`` keys; same
``forward(data_dict) -> output_dict`` contract, including ``gt_node_corr_indices/overlaps`` (model.py:112-126, computed
whenever ``data_dict`` carries the ground-truth ``transform``; the reference requires it).
"""
import torch
import torch.nn as nn

from . import functional as GF
from .backbone import KPConvFPN
from .modules.geotransformer import GeometricTransformer, SuperPointMatching, LocalGlobalRegistration
from .modules.sinkhorn import LearnableLogOptimalTransport


class GeoTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.num_points_in_patch = cfg.model.num_points_in_patch
        self.matching_radius = cfg.model.ground_truth_matching_radius
        self.fine_level = cfg.model.fine_level            # 1 for 3DMatch/KITTI (model.py:77,80), 0 for ModelNet
        b = cfg.backbone
        self.backbone = KPConvFPN(b.input_dim, b.output_dim, b.init_dim, b.kernel_size, b.init_radius, b.init_sigma,
                                  b.group_norm, num_stages=b.num_stages,
                                  finest_decoder=2 if self.fine_level == 1 else 1)
        g = cfg.geotransformer
        self.transformer = GeometricTransformer(g.input_dim, g.output_dim, g.hidden_dim, g.num_heads, g.blocks, g.sigma_d,
                                                g.sigma_a, g.angle_k, reduction_a=g.reduction_a)
        self.coarse_matching = SuperPointMatching(cfg.coarse_matching.num_correspondences,
                                                  cfg.coarse_matching.dual_normalization)
        f = cfg.fine_matching
        self.fine_matching = LocalGlobalRegistration(
            f.topk, f.acceptance_radius, mutual=f.mutual, confidence_threshold=f.confidence_threshold,
            use_dustbin=f.use_dustbin, use_global_score=f.use_global_score,
            correspondence_threshold=f.correspondence_threshold, correspondence_limit=f.correspondence_limit,
            num_refinement_steps=f.num_refinement_steps)
        self.optimal_transport = LearnableLogOptimalTransport(cfg.model.num_sinkhorn_iterations)

    @torch.no_grad()
    def forward(self, data_dict, taps=None):
        out = {}
        marks = data_dict.get('_stage_events')            # profiling hook: list receiving (label, CUDA event) pairs
        def mark(label):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((label, e))
        mark('start')
        feats = data_dict['features']
        lens = data_dict['lengths']
        fl = self.fine_level
        # lengths are needed on the host to slice ref/src (the reference does three .item() syncs, model.py:76-78);
        # the collate keeps host copies so no sync happens here
        lens_h = data_dict.get('lengths_host')
        if lens_h is None:
            lens_h = [l.tolist() for l in lens]
        nc, nf, n0 = int(lens_h[-1][0]), int(lens_h[fl][0]), int(lens_h[0][0])
        points_c, points_f, points = data_dict['points'][-1], data_dict['points'][fl], data_dict['points'][0]
        ref_c, src_c = points_c[:nc], points_c[nc:]
        ref_f, src_f = points_f[:nf], points_f[nf:]
        out.update(ref_points_c=ref_c, src_points_c=src_c, ref_points_f=ref_f, src_points_f=src_f,
                   ref_points=points[:n0], src_points=points[n0:])

        K = self.num_points_in_patch
        _, ref_node_masks, ref_knn_idx, ref_knn_masks = GF.point_to_node_partition(ref_f, ref_c, K)
        _, src_node_masks, src_knn_idx, src_knn_masks = GF.point_to_node_partition(src_f, src_c, K)
        if taps is not None:
            taps.update(ref_node_masks=ref_node_masks, src_node_masks=src_node_masks, ref_node_knn_indices=ref_knn_idx,
                        src_node_knn_indices=src_knn_idx, ref_node_knn_masks=ref_knn_masks, src_node_knn_masks=src_knn_masks)

        # ground-truth superpoint correspondences (reference model.py:106-126).  Launched here, read back after the last
        # host sync of the forward so that it costs no extra synchronisation.
        gt_pending = None
        transform = data_dict.get('transform')
        if transform is not None:
            ar_r = GF.scratch_arange(ref_c.shape[0], ref_c.device, 'ar_ref')
            ar_s = GF.scratch_arange(src_c.shape[0], src_c.device, 'ar_src')
            _, _, ref_all_pts = GF.gather_patches(ar_r, ref_knn_idx, ref_knn_masks, ref_f)
            _, _, src_all_pts = GF.gather_patches(ar_s, src_knn_idx, src_knn_masks, src_f)
            gt_pending = GF.node_correspondences(ref_c, src_c, ref_all_pts, src_all_pts, transform, self.matching_radius,
                                                 ref_node_masks, src_node_masks, ref_knn_masks, src_knn_masks)

        mark('partition+gt')
        native = getattr(self, '_native', None)          # NativeModel: backbone / transformer as one C call each
        feats_list = native.backbone_forward(feats, data_dict) if native is not None else self.backbone(feats, data_dict)
        feats_c, feats_f = feats_list[-1], feats_list[0]
        mark('backbone')
        if taps is not None:
            taps['feats_c'], taps['feats_f'] = feats_c, feats_f

        ref_fc, src_fc = self.transformer(ref_c, src_c, feats_c[:nc], feats_c[nc:], native=native)
        mark('transformer')
        ref_fc_n, src_fc_n = GF.l2_normalize(ref_fc), GF.l2_normalize(src_fc)
        ref_ff, src_ff = feats_f[:nf], feats_f[nf:]
        out.update(ref_feats_c=ref_fc_n, src_feats_c=src_fc_n, ref_feats_f=ref_ff, src_feats_f=src_ff)

        # fewer than num_correspondences rows exist only when #valid ref x #valid src superpoints is smaller (tiny clouds):
        # the count stays on the device, the padding rows become empty patches (no fine correspondences, so LGR is
        # unaffected) and the per-patch outputs are trimmed after the forward's last host sync
        ref_corr, src_corr, node_scores, corr_count = self.coarse_matching(ref_fc_n, src_fc_n, ref_node_masks, src_node_masks,
                                                                           defer_count=True)
        forced = data_dict.get('forced_node_corr')       # test hook: teacher-forced coarse correspondences
        if forced is not None:
            ref_corr, src_corr, node_scores = forced
            corr_count = None
        out.update(ref_node_corr_indices=ref_corr, src_node_corr_indices=src_corr, node_corr_scores=node_scores)

        rk_idx, rk_masks, rk_pts = GF.gather_patches(ref_corr, ref_knn_idx, ref_knn_masks, ref_f)
        sk_idx, sk_masks, sk_pts = GF.gather_patches(src_corr, src_knn_idx, src_knn_masks, src_f)
        out.update(ref_node_corr_knn_points=rk_pts, src_node_corr_knn_points=sk_pts, ref_node_corr_knn_masks=rk_masks,
                   src_node_corr_knn_masks=sk_masks)

        scores = GF.patch_scores(ref_ff, src_ff, rk_idx, sk_idx)
        if taps is not None:
            taps['matching_scores_raw'] = scores
        scores = self.optimal_transport(scores, rk_masks, sk_masks)
        out['matching_scores'] = scores
        mark('matching+sinkhorn')

        rc, sc, cs, T = self.fine_matching(rk_pts, sk_pts, rk_masks, sk_masks, scores, node_scores)
        out.update(ref_corr_points=rc, src_corr_points=sc, corr_scores=cs, estimated_transform=T)
        mark('lgr')
        if gt_pending is not None:
            out['gt_node_corr_indices'], out['gt_node_corr_overlaps'] = GF.finish_node_correspondences(*gt_pending)
        if corr_count is not None:
            kk = int(corr_count.item())          # already complete: LGR synchronised the stream
            if kk < ref_corr.shape[0]:
                for key in ('ref_node_corr_indices', 'src_node_corr_indices', 'node_corr_scores', 'ref_node_corr_knn_points',
                            'src_node_corr_knn_points', 'ref_node_corr_knn_masks', 'src_node_corr_knn_masks', 'matching_scores'):
                    out[key] = out[key][:kk]
        return out


def enable_native(model):
    """Attach the C++ stage drivers (geotransformer_b200.native): same kernels and results, ~10x less host time per pair.
    Call after the weights are loaded and the model is on its device."""
    from .native import NativeModel
    model._native = NativeModel(model)
    if not getattr(model, '_native_hook', False):
        # NativeModel snapshots pointers AND derived copies (fused q|k|v weights, transposes): rebuild it whenever new
        # weights are loaded, otherwise the raw parameters would update in place while the derived copies stayed stale
        def _rebuild(module, incompatible_keys):
            if getattr(module, '_native', None) is not None:
                module._native = NativeModel(module)
        model.register_load_state_dict_post_hook(_rebuild)
        model._native_hook = True
    return model


def create_model(cfg):
    return GeoTransformer(cfg)
