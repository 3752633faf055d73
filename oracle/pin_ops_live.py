"""TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Pins the Boundary-2 op restatements of oracle/geo_oracle.py
against the REAL reference functions of ``geotransformer.modules.ops`` (imported unmodified through oracle/ref_harness.py) on
seeded random inputs: ``pairwise_distance`` (pairwise_distance.py:4-31), ``knn_partition`` / ``get_point_to_node_indices`` /
``point_to_node_partition`` / ``ball_query_partition`` (pointcloud_partition.py:9-107,159-175), ``apply_transform``
(transformation.py:7-60).  Run in its own process (the harness patches ``Tensor.cuda`` and ``sys.modules``):

    python -m oracle.pin_ops_live        # prints 'pinned: ...' and exits 0
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import geo_oracle as G, ref_harness  # noqa: E402


def main():
    ref_harness.install()
    from geotransformer.modules import ops as R
    checked = []
    for seed, n, m in ((0, 500, 40), (1, 64, 64), (2, 2000, 7)):
        g = torch.Generator().manual_seed(seed)
        pts = torch.rand(n, 3, generator=g) * 2.0
        nodes = pts[torch.randperm(n, generator=g)[:m]].contiguous() + 0.01 * torch.randn(m, 3, generator=g)
        feats_a = torch.nn.functional.normalize(torch.randn(m, 32, generator=g), dim=1)
        feats_b = torch.nn.functional.normalize(torch.randn(n, 32, generator=g), dim=1)
        assert torch.equal(R.pairwise_distance(nodes, pts), G.pairwise_distance(nodes, pts))
        assert torch.equal(R.pairwise_distance(feats_a, feats_b, normalized=True), G.pairwise_distance(feats_a, feats_b, normalized=True))
        for k in (1, 8, 33):
            kk = min(k, n)
            assert torch.equal(R.knn_partition(pts, nodes, kk), G.knn_partition(pts, nodes, kk))
            d_r, i_r = R.knn_partition(pts, nodes, kk, return_distance=True)
            d_o, i_o = G.knn_partition(pts, nodes, kk, return_distance=True)
            assert torch.equal(i_r, i_o) and torch.equal(d_r, d_o)
        assert torch.equal(R.get_point_to_node_indices(pts, nodes), G.get_point_to_node_indices(pts, nodes))
        i_r, c_r = R.get_point_to_node_indices(pts, nodes, return_counts=True)
        i_o, c_o = G.get_point_to_node_indices(pts, nodes, return_counts=True)
        assert torch.equal(i_r, i_o) and torch.equal(c_r, c_o)
        for limit in (4, 16):
            got, want = G.point_to_node_partition(pts, nodes, limit), R.point_to_node_partition(pts, nodes, limit)
            assert len(got) == len(want) == 4 and all(torch.equal(a, b) for a, b in zip(got, want))
            for radius in (0.05, 0.3):
                got, want = G.ball_query_partition(pts, nodes, radius, limit, return_count=True), R.ball_query_partition(pts, nodes, radius, limit, return_count=True)
                assert all(torch.equal(a, b) for a, b in zip(got, want))
        T = torch.eye(4)
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        T[:3, :3], T[:3, 3] = q, torch.randn(3, generator=g)
        assert torch.equal(R.apply_transform(pts, T), G.apply_transform(pts, T))
        checked.append((n, m))
    print('pinned: pairwise_distance, knn_partition, get_point_to_node_indices, point_to_node_partition, ball_query_partition, '
          f'apply_transform == the real reference on {checked}')


if __name__ == '__main__':
    main()
