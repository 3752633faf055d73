"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-pair registration forward.

This is the checker the CUDA path is compared against; it is never imported by ``geotransformer_b200``.
It is a functional re-telling (state_dict in, tensors out) of the reference modules, written to perform the
same fp32 torch operations in the same order so that, on CPU, it reproduces the reference bit for bit
(pinned by ``oracle/make_golden.py`` against the real reference imported from /root/reference, and on the GPU
box by the committed fixtures in ``tests/golden/``).

Reference map (all paths under /root/reference):
  collate                  geotransformer/utils/data.py:13-77,139-189
  kpconv                   geotransformer/modules/kpconv/kpconv.py:79-122
  group_norm / unary / ... geotransformer/modules/kpconv/modules.py:33-225, functional.py:6-67
  backbone                 experiments/*/backbone.py (3dmatch :48-87, kitti :76-124, modelnet :38-73)
  point_to_node_partition  geotransformer/modules/ops/pointcloud_partition.py:60-107
  pairwise_distance        geotransformer/modules/ops/pairwise_distance.py:4-31
  structure embedding      geotransformer/modules/geotransformer/geotransformer.py:27-72,
                           geotransformer/modules/transformer/positional_embedding.py:8-34
  transformer              geotransformer/modules/transformer/{rpe_transformer.py:18-131,
                           vanilla_transformer.py:15-129, output_layer.py:6-21, conditional_transformer.py:73-117},
                           geotransformer/modules/geotransformer/geotransformer.py:114-155
  superpoint matching      geotransformer/modules/geotransformer/superpoint_matching.py:13-50
  optimal transport        geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66
  LGR + procrustes         geotransformer/modules/geotransformer/local_global_registration.py:49-235,
                           geotransformer/modules/registration/procrustes.py:6-73,
                           geotransformer/modules/ops/transformation.py:7-60
  model assembly           experiments/*/model.py:69-212
"""

import numpy as np
import torch
import torch.nn.functional as F

from . import collate_oracle


# ------------------------------------------------------------------------------------------------ collate

def grid_subsample(points, lengths, voxel_size, impl=collate_oracle):
    return impl.grid_subsampling(points, lengths, voxel_size)


def radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit, impl=collate_oracle):
    idx = impl.radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)
    if neighbor_limit > 0:
        idx = idx[:, :neighbor_limit]
    return idx.contiguous()  # the reference's to_cuda() densifies the slice (utils/torch.py:113-123)


def collate_pair(pair, cfg, neighbor_limits, impl=collate_oracle):
    """registration_collate_fn_stack_mode for one pair (utils/data.py:139-189): stack [ref; src]."""
    ref, src = torch.from_numpy(pair['ref_points']), torch.from_numpy(pair['src_points'])
    points = torch.cat([ref, src], dim=0)
    lengths = torch.LongTensor([ref.shape[0], src.shape[0]])
    feats = torch.cat([torch.from_numpy(pair['ref_feats']), torch.from_numpy(pair['src_feats'])], dim=0)
    d = precompute_data_stack_mode(points, lengths, cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                             cfg.backbone.init_radius, neighbor_limits, impl)
    d['features'] = feats
    d['transform'] = torch.from_numpy(pair['transform'])
    d['batch_size'] = 1
    return d


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits, impl=collate_oracle):
    """utils/data.py:13-77.  `voxel_size *= 2` runs for every i including i == 0 (data.py:22-28), so stage
    i >= 1 is subsampled with voxel_size * 2**i; the radius doubles per stage (data.py:31-69)."""
    pts, lens = [], []
    v = voxel_size
    p, l = points, lengths
    for i in range(num_stages):
        if i > 0:
            p, l = grid_subsample(p, l, v, impl)
        pts.append(p)
        lens.append(l)
        v *= 2
    neighbors, subsampling, upsampling = [], [], []
    r = radius
    for i in range(num_stages):
        neighbors.append(radius_search(pts[i], pts[i], lens[i], lens[i], r, neighbor_limits[i], impl))
        if i < num_stages - 1:
            subsampling.append(radius_search(pts[i + 1], pts[i], lens[i + 1], lens[i], r, neighbor_limits[i], impl))
            upsampling.append(radius_search(pts[i], pts[i + 1], lens[i], lens[i + 1], r * 2, neighbor_limits[i + 1], impl))
        r *= 2
    return {'points': pts, 'lengths': lens, 'neighbors': neighbors, 'subsampling': subsampling,
            'upsampling': upsampling}


def calibrate_neighbors(pairs, cfg, keep_ratio=0.8, sample_threshold=2000, impl=collate_oracle):
    """utils/data.py:190-217 calibrate_neighbors_stack_mode over a list of raw pairs: the keep_ratio quantile of the
    neighbourhood sizes per stage, histogrammed up to ceil(4/3 pi (r/voxel + 1)^3)."""
    b = cfg.backbone
    hist_n = int(np.ceil(4 / 3 * np.pi * (b.init_radius / b.init_voxel_size + 1) ** 3))
    hists = np.zeros((b.num_stages, hist_n), dtype=np.int64)
    for pair in pairs:
        data = collate_pair(pair, cfg, [hist_n] * b.num_stages, impl=impl)
        for s, nbr in enumerate(data['neighbors']):
            counts = (nbr.numpy() < nbr.shape[0]).sum(axis=1)
            hists[s] += np.bincount(counts, minlength=hist_n)[:hist_n]
        if hists.sum(axis=1).min() > sample_threshold:
            break
    cum = np.cumsum(hists.T, axis=0)
    return np.sum(cum < keep_ratio * cum[hist_n - 1, :], axis=0)


def canonical_neighbors(q_points, s_points, table):
    """Re-orders every row of a neighbour table by (fp32 reference distance, index): removes the implementation-defined
    order inside exact-distance tie groups (radius_neighbors_cpu.cpp sorts with an unstable std::sort)."""
    ns = s_points.shape[0]
    pad = torch.cat([s_points, torch.full((1, 3), 1e9)], dim=0)
    diff = q_points.unsqueeze(1) - pad[table]
    d = diff[..., 0] * diff[..., 0]
    d = d + diff[..., 1] * diff[..., 1]
    d = d + diff[..., 2] * diff[..., 2]
    d = torch.where(table >= ns, torch.full_like(d, float('inf')), d)
    key_idx = table.argsort(dim=1, stable=True)
    d_sorted_by_idx = torch.gather(d, 1, key_idx)
    order2 = d_sorted_by_idx.argsort(dim=1, stable=True)
    return torch.gather(torch.gather(table, 1, key_idx), 1, order2)


# ------------------------------------------------------------------------------------------------ backbone

def _gather_rows(table, index):
    """ops/index_select.py:4-31 with dim=0."""
    out = table.index_select(0, index.reshape(-1))
    return out.view(*index.shape, *table.shape[1:])


def kpconv(sd, pre, s_feats, q_points, s_points, nbr, sigma):
    """kpconv.py:79-122.  sd[pre+'weights'] (K,Cin,Cout), bias, kernel_points (K,3)."""
    kp, W = sd[pre + 'kernel_points'], sd[pre + 'weights']
    pad_pts = torch.cat([s_points, torch.zeros_like(s_points[:1]) + 1e6], 0)
    rel = _gather_rows(pad_pts, nbr) - q_points.unsqueeze(1)                     # (M,H,3)
    diff = rel.unsqueeze(2) - kp                                                 # (M,H,K,3)
    sq = torch.sum(diff ** 2, dim=3)
    infl = torch.clamp(1 - torch.sqrt(sq) / sigma, min=0.0).transpose(1, 2)      # (M,K,H)
    pad_feats = torch.cat((s_feats, torch.zeros_like(s_feats[:1])), 0)
    nf = _gather_rows(pad_feats, nbr)                                            # (M,H,C)
    wf = torch.matmul(infl, nf).permute(1, 0, 2)                                 # (K,M,C)
    out = torch.sum(torch.matmul(wf, W), dim=0)                                  # (M,Cout)
    n_valid = torch.sum(torch.gt(torch.sum(nf, dim=-1), 0.0), dim=-1)
    n_valid = torch.max(n_valid, torch.ones_like(n_valid))
    out = out / n_valid.unsqueeze(1)
    if (pre + 'bias') in sd:
        out = out + sd[pre + 'bias']
    return out


def group_norm(sd, pre, x, groups):
    """modules.py:33-50: nn.GroupNorm over the whole stacked (1,C,N) tensor."""
    y = F.group_norm(x.transpose(0, 1).unsqueeze(0), groups, sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)
    return y.squeeze(0).transpose(0, 1).squeeze()


def unary(sd, pre, x, groups, relu=True, norm=True):
    """modules.py:53-104 (UnaryBlock / LastUnaryBlock)."""
    x = F.linear(x, sd[pre + 'mlp.weight'], sd[pre + 'mlp.bias'])
    if norm:
        x = group_norm(sd, pre + 'norm.', x, groups)
    if relu:
        x = F.leaky_relu(x, 0.1)
    return x


def maxpool(x, nbr):
    """functional.py:54-67."""
    x = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return _gather_rows(x, nbr).max(1)[0]


def nearest_upsample(x, up):
    """functional.py:6-22."""
    x = torch.cat((x, torch.zeros_like(x[:1])), 0)
    return _gather_rows(x, up[:, 0])


def conv_block(sd, pre, feats, q_pts, s_pts, nbr, sigma, groups):
    """modules.py:107-148."""
    x = kpconv(sd, pre + 'KPConv.', feats, q_pts, s_pts, nbr, sigma)
    return F.leaky_relu(group_norm(sd, pre + 'norm.', x, groups), 0.1)


def residual_block(sd, pre, feats, q_pts, s_pts, nbr, sigma, groups, strided):
    """modules.py:151-225 (bottleneck; presence of unary1 / unary_shortcut is read off the state_dict)."""
    x = feats
    if (pre + 'unary1.mlp.weight') in sd:
        x = unary(sd, pre + 'unary1.', x, groups)
    x = kpconv(sd, pre + 'KPConv.', x, q_pts, s_pts, nbr, sigma)
    x = F.leaky_relu(group_norm(sd, pre + 'norm_conv.', x, groups), 0.1)
    x = unary(sd, pre + 'unary2.', x, groups, relu=False)
    sc = maxpool(feats, nbr) if strided else feats
    if (pre + 'unary_shortcut.mlp.weight') in sd:
        sc = unary(sd, pre + 'unary_shortcut.', sc, groups, relu=False)
    return F.leaky_relu(x + sc, 0.1)


def backbone(sd, cfg, feats, data, pre='backbone.', taps=None):
    """KPConv-FPN of the three experiments as one generic loop over stages."""
    S, g = cfg.backbone.num_stages, cfg.backbone.group_norm
    fine = 2 if cfg.name in ('3dmatch', 'kitti') else 1           # 1-based level of the finest decoder output
    pts, nb, sub, up = data['points'], data['neighbors'], data['subsampling'], data['upsampling']
    sig = cfg.backbone.init_sigma
    enc = []
    x = conv_block(sd, pre + 'encoder1_1.', feats, pts[0], pts[0], nb[0], sig, g)
    if taps is not None:
        taps['encoder1_1'] = x
    x = residual_block(sd, pre + 'encoder1_2.', x, pts[0], pts[0], nb[0], sig, g, False)
    if taps is not None:
        taps['encoder1_2'] = x
    enc.append(x)
    for s in range(2, S + 1):
        x = residual_block(sd, pre + f'encoder{s}_1.', x, pts[s - 1], pts[s - 2], sub[s - 2], sig, g, True)
        if taps is not None:
            taps[f'encoder{s}_1'] = x
        sig = sig * 2
        x = residual_block(sd, pre + f'encoder{s}_2.', x, pts[s - 1], pts[s - 1], nb[s - 1], sig, g, False)
        x = residual_block(sd, pre + f'encoder{s}_3.', x, pts[s - 1], pts[s - 1], nb[s - 1], sig, g, False)
        if taps is not None:
            taps[f'encoder{s}_3'] = x
        enc.append(x)
    outs = [enc[-1]]
    latent = enc[-1]
    for lvl in range(S - 1, fine - 1, -1):                        # decoder{lvl}
        latent = torch.cat([nearest_upsample(latent, up[lvl - 1]), enc[lvl - 1]], dim=1)
        last = lvl == fine
        latent = unary(sd, pre + f'decoder{lvl}.', latent, g, relu=not last, norm=not last)
        if taps is not None:
            taps[f'decoder{lvl}'] = latent
        outs.append(latent)
    outs.reverse()
    return outs                                                   # [fine ... coarse]


# ------------------------------------------------------------------------------------------------ partition

def pairwise_distance(x, y, normalized=False):
    """ops/pairwise_distance.py:4-31 (channel-last)."""
    xy = torch.matmul(x, y.transpose(-1, -2))
    if normalized:
        sq = 2.0 - 2.0 * xy
    else:
        x2 = torch.sum(x ** 2, dim=-1).unsqueeze(-1)
        y2 = torch.sum(y ** 2, dim=-1).unsqueeze(-2)
        sq = x2 - 2 * xy + y2
    return sq.clamp(min=0.0)


def point_to_node_partition(points, nodes, point_limit):
    """ops/pointcloud_partition.py:60-107."""
    sq = pairwise_distance(nodes, points)                         # (M,N)
    point_to_node = sq.min(dim=0)[1]
    node_masks = torch.zeros(nodes.shape[0], dtype=torch.bool)
    node_masks.index_fill_(0, point_to_node, True)
    match = torch.zeros_like(sq, dtype=torch.bool)
    match[point_to_node, torch.arange(points.shape[0])] = True
    sq = sq.masked_fill(~match, 1e12)
    knn = sq.topk(k=point_limit, dim=1, largest=False)[1]
    knn_nodes = point_to_node[knn]
    knn_masks = torch.eq(knn_nodes, torch.arange(nodes.shape[0]).unsqueeze(1).expand(-1, point_limit))
    knn = knn.masked_fill(~knn_masks, points.shape[0])
    return point_to_node, node_masks, knn, knn_masks


def knn_partition(points, nodes, k, return_distance=False):
    """ops/pointcloud_partition.py:35-57."""
    k = min(k, points.shape[0])
    sq = pairwise_distance(nodes, points)
    knn_sq, knn = sq.topk(dim=1, k=k, largest=False)
    return (torch.sqrt(knn_sq), knn) if return_distance else knn


def get_point_to_node_indices(points, nodes, return_counts=False):
    """ops/pointcloud_partition.py:9-32."""
    indices = pairwise_distance(points, nodes).min(dim=1)[1]
    if return_counts:
        u, c = torch.unique(indices, return_counts=True)
        sizes = torch.zeros(nodes.shape[0], dtype=torch.long)
        sizes[u] = c
        return indices, sizes
    return indices


def ball_query_partition(points, nodes, radius, point_limit, return_count=False):
    """ops/pointcloud_partition.py:159-175."""
    d, idx = knn_partition(points, nodes, point_limit, return_distance=True)
    masks = torch.lt(d, radius)
    idx = torch.where(masks, idx, torch.full_like(idx, points.shape[0]))
    return (idx, masks, masks.sum(1)) if return_count else (idx, masks)


# ------------------------------------------------------------------------------------------------ transformer

def embedding_indices(points, sigma_d, sigma_a, angle_k):
    """geotransformer.py:27-55 for one cloud: d_indices (N,N), a_indices (N,N,k)."""
    n = points.shape[0]
    dist = torch.sqrt(pairwise_distance(points, points))
    d_idx = dist / sigma_d
    knn = dist.topk(k=angle_k + 1, dim=1, largest=False)[1][:, 1:]             # (N,k)
    ref_v = points[knn] - points.unsqueeze(1)                                     # (N,k,3)  p_knn - p_i
    anc_v = points.unsqueeze(0) - points.unsqueeze(1)                             # (N,N,3)  p_j - p_i
    ref_e = ref_v.unsqueeze(1).expand(n, n, angle_k, 3)
    anc_e = anc_v.unsqueeze(2).expand(n, n, angle_k, 3)
    sin_v = torch.linalg.norm(torch.cross(ref_e, anc_e, dim=-1), dim=-1)
    cos_v = torch.sum(ref_e * anc_e, dim=-1)
    factor_a = 180.0 / (sigma_a * np.pi)
    return d_idx, torch.atan2(sin_v, cos_v) * factor_a


def sinusoid(idx, div_term):
    """positional_embedding.py:27-33: interleaved [sin w0, cos w0, sin w1, ...]."""
    om = idx.reshape(-1, 1, 1) * div_term.view(1, -1, 1)
    emb = torch.cat([torch.sin(om), torch.cos(om)], dim=2)
    return emb.view(*idx.shape, 2 * div_term.shape[0])


def structure_embedding(sd, pre, points, sigma_d, sigma_a, angle_k):
    """geotransformer.py:57-72 (reduction 'max'): E (N,N,C)."""
    d_idx, a_idx = embedding_indices(points, sigma_d, sigma_a, angle_k)
    div = sd[pre + 'embedding.div_term']
    d_emb = F.linear(sinusoid(d_idx, div), sd[pre + 'proj_d.weight'], sd[pre + 'proj_d.bias'])
    a_emb = F.linear(sinusoid(a_idx, div), sd[pre + 'proj_a.weight'], sd[pre + 'proj_a.bias']).max(dim=2)[0]
    return d_emb + a_emb


def _heads(x, h):
    return x.view(x.shape[0], h, x.shape[1] // h).permute(1, 0, 2)               # (H,N,c)


def _attention_tail(sd, pre, hidden, inp):
    """linear + residual LayerNorm (rpe_transformer.py:99-103 / vanilla :97-101), FFN (output_layer.py:15-21)."""
    c = inp.shape[1]
    h = F.linear(hidden, sd[pre + 'attention.linear.weight'], sd[pre + 'attention.linear.bias'])
    x = F.layer_norm(h + inp, (c,), sd[pre + 'attention.norm.weight'], sd[pre + 'attention.norm.bias'])
    y = F.relu(F.linear(x, sd[pre + 'output.expand.weight'], sd[pre + 'output.expand.bias']))
    y = F.linear(y, sd[pre + 'output.squeeze.weight'], sd[pre + 'output.squeeze.bias'])
    return F.layer_norm(x + y, (c,), sd[pre + 'output.norm.weight'], sd[pre + 'output.norm.bias'])


def rpe_self_layer(sd, pre, x, emb, num_heads):
    """rpe_transformer.py:36-103 with memory = input, no masks."""
    a = pre + 'attention.attention.'
    c = x.shape[1]
    q = _heads(F.linear(x, sd[a + 'proj_q.weight'], sd[a + 'proj_q.bias']), num_heads)
    k = _heads(F.linear(x, sd[a + 'proj_k.weight'], sd[a + 'proj_k.bias']), num_heads)
    v = _heads(F.linear(x, sd[a + 'proj_v.weight'], sd[a + 'proj_v.bias']), num_heads)
    p = F.linear(emb, sd[a + 'proj_p.weight'], sd[a + 'proj_p.bias'])           # (N,M,C)
    p = p.view(p.shape[0], p.shape[1], num_heads, c // num_heads).permute(2, 0, 1, 3)
    s_p = torch.einsum('hnc,hnmc->hnm', q, p)
    s_e = torch.einsum('hnc,hmc->hnm', q, k)
    s = F.softmax((s_e + s_p) / (c // num_heads) ** 0.5, dim=-1)
    hid = torch.matmul(s, v).permute(1, 0, 2).reshape(x.shape[0], c)
    return _attention_tail(sd, pre, hid, x)


def cross_layer(sd, pre, x, mem, num_heads):
    """vanilla_transformer.py:50-101, no masks."""
    a = pre + 'attention.attention.'
    c = x.shape[1]
    q = _heads(F.linear(x, sd[a + 'proj_q.weight'], sd[a + 'proj_q.bias']), num_heads)
    k = _heads(F.linear(mem, sd[a + 'proj_k.weight'], sd[a + 'proj_k.bias']), num_heads)
    v = _heads(F.linear(mem, sd[a + 'proj_v.weight'], sd[a + 'proj_v.bias']), num_heads)
    s = F.softmax(torch.einsum('hnc,hmc->hnm', q, k) / (c // num_heads) ** 0.5, dim=-1)
    hid = torch.matmul(s, v).permute(1, 0, 2).reshape(x.shape[0], c)
    return _attention_tail(sd, pre, hid, x)


def geometric_transformer(sd, cfg, ref_points, src_points, ref_feats, src_feats, pre='transformer.', taps=None):
    """geotransformer.py:114-155 + conditional_transformer.py:97-117 (sequential cross updates)."""
    g = cfg.geotransformer
    e0 = structure_embedding(sd, pre + 'embedding.', ref_points, g.sigma_d, g.sigma_a, g.angle_k)
    e1 = structure_embedding(sd, pre + 'embedding.', src_points, g.sigma_d, g.sigma_a, g.angle_k)
    f0 = F.linear(ref_feats, sd[pre + 'in_proj.weight'], sd[pre + 'in_proj.bias'])
    f1 = F.linear(src_feats, sd[pre + 'in_proj.weight'], sd[pre + 'in_proj.bias'])
    if taps is not None:
        taps['ref_embeddings'], taps['src_embeddings'] = e0, e1
    for i, blk in enumerate(g.blocks):
        lp = pre + f'transformer.layers.{i}.'
        if blk == 'self':
            f0 = rpe_self_layer(sd, lp, f0, e0, g.num_heads)
            f1 = rpe_self_layer(sd, lp, f1, e1, g.num_heads)
        else:
            f0 = cross_layer(sd, lp, f0, f1, g.num_heads)
            f1 = cross_layer(sd, lp, f1, f0, g.num_heads)
        if taps is not None:
            taps[f'layer{i}_ref'], taps[f'layer{i}_src'] = f0, f1
    f0 = F.linear(f0, sd[pre + 'out_proj.weight'], sd[pre + 'out_proj.bias'])
    f1 = F.linear(f1, sd[pre + 'out_proj.weight'], sd[pre + 'out_proj.bias'])
    return f0, f1


# ------------------------------------------------------------------------------------------------ matching

def superpoint_matching(ref_feats, src_feats, ref_masks, src_masks, num_corr, dual=True):
    """superpoint_matching.py:13-50."""
    ri = torch.nonzero(ref_masks, as_tuple=True)[0]
    si = torch.nonzero(src_masks, as_tuple=True)[0]
    sc = torch.exp(-pairwise_distance(ref_feats[ri], src_feats[si], normalized=True))
    if dual:
        sc = (sc / sc.sum(dim=1, keepdim=True)) * (sc / sc.sum(dim=0, keepdim=True))
    k = min(num_corr, sc.numel())
    val, idx = sc.view(-1).topk(k=k, largest=True)
    return ri[idx // sc.shape[1]], si[idx % sc.shape[1]], val


def optimal_transport(alpha, scores, row_masks, col_masks, num_iter, inf=1e12):
    """learnable_sinkhorn.py:13-66."""
    b, nr, nc = scores.shape
    prm = torch.zeros(b, nr + 1, dtype=torch.bool)
    prm[:, :nr] = ~row_masks
    pcm = torch.zeros(b, nc + 1, dtype=torch.bool)
    pcm[:, :nc] = ~col_masks
    psm = torch.logical_or(prm.unsqueeze(2), pcm.unsqueeze(1))
    pc = alpha.expand(b, nr, 1)
    pr = alpha.expand(b, 1, nc + 1)
    ps = torch.cat([torch.cat([scores, pc], dim=-1), pr], dim=1)
    ps = ps.masked_fill(psm, -inf)
    nvr = row_masks.float().sum(1)
    nvc = col_masks.float().sum(1)
    norm = -torch.log(nvr + nvc)
    log_mu = torch.empty(b, nr + 1)
    log_mu[:, :nr] = norm.unsqueeze(1)
    log_mu[:, nr] = torch.log(nvc) + norm
    log_mu[prm] = -inf
    log_nu = torch.empty(b, nc + 1)
    log_nu[:, :nc] = norm.unsqueeze(1)
    log_nu[:, nc] = torch.log(nvr) + norm
    log_nu[pcm] = -inf
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(num_iter):
        u = log_mu - torch.logsumexp(ps + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(ps + u.unsqueeze(2), dim=1)
    return ps + u.unsqueeze(2) + v.unsqueeze(1) - norm.unsqueeze(1).unsqueeze(2)


def weighted_procrustes(src, ref, weights, eps=1e-5):
    """procrustes.py:43-73 (batched, weight_thresh 0): returns (B,4,4)."""
    if src.ndim == 2:
        src, ref, weights = src.unsqueeze(0), ref.unsqueeze(0), weights.unsqueeze(0)
        squeeze = True
    else:
        squeeze = False
    b = src.shape[0]
    w = torch.where(torch.lt(weights, 0.0), torch.zeros_like(weights), weights)
    w = (w / (torch.sum(w, dim=1, keepdim=True) + eps)).unsqueeze(2)
    sc = torch.sum(src * w, dim=1, keepdim=True)
    rc = torch.sum(ref * w, dim=1, keepdim=True)
    H = (src - sc).permute(0, 2, 1) @ (w * (ref - rc))
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    eye = torch.eye(3).unsqueeze(0).repeat(b, 1, 1)
    eye[:, -1, -1] = torch.sign(torch.det(V @ Ut))
    R = V @ eye @ Ut
    t = (rc.permute(0, 2, 1) - R @ sc.permute(0, 2, 1)).squeeze(2)
    T = torch.eye(4).unsqueeze(0).repeat(b, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return T.squeeze(0) if squeeze else T


def apply_transform(points, T):
    """transformation.py:37-49."""
    if T.ndim == 2:
        return torch.matmul(points, T[:3, :3].transpose(-1, -2)) + T[:3, 3]
    return torch.matmul(points, T[:, :3, :3].transpose(-1, -2)) + T[:, None, :3, 3]


def correspondence_matrix(score_mat, ref_masks, src_masks, k, conf, mutual=True):
    """local_global_registration.py:49-83 (no dustbin)."""
    mask = torch.logical_and(ref_masks.unsqueeze(2), src_masks.unsqueeze(1))
    b, n, m = score_mat.shape
    bi = torch.arange(b)
    rs, ri = score_mat.topk(k=k, dim=2)
    rmat = torch.zeros_like(score_mat)
    rmat[bi.view(b, 1, 1).expand(-1, n, k), torch.arange(n).view(1, n, 1).expand(b, -1, k), ri] = rs
    rc = torch.gt(rmat, conf)
    ss, si = score_mat.topk(k=k, dim=1)
    smat = torch.zeros_like(score_mat)
    smat[bi.view(b, 1, 1).expand(-1, k, m), si, torch.arange(m).view(1, 1, m).expand(b, k, -1)] = ss
    scm = torch.gt(smat, conf)
    cm = torch.logical_and(rc, scm) if mutual else torch.logical_or(rc, scm)
    return torch.logical_and(cm, mask)


def local_global_registration(cfg, ref_knn_pts, src_knn_pts, ref_masks, src_masks, score_mat, taps=None):
    """local_global_registration.py:137-235."""
    fm = cfg.fine_matching
    score = torch.exp(score_mat)
    cm = correspondence_matrix(score, ref_masks, src_masks, fm.topk, fm.confidence_threshold, fm.mutual)
    score = score * cm.float()
    bi, ri, si = torch.nonzero(cm, as_tuple=True)
    ref_c, src_c, sc = ref_knn_pts[bi, ri], src_knn_pts[bi, si], score[bi, ri, si]
    if taps is not None:
        taps['corr_batch_indices'] = bi
    # per-patch chunks with >= correspondence_threshold entries
    bounds = [0] + (torch.nonzero(bi[1:] != bi[:-1], as_tuple=True)[0] + 1).tolist() + [bi.shape[0]]
    chunks = [(x, y) for x, y in zip(bounds[:-1], bounds[1:]) if y - x >= fm.correspondence_threshold]
    if len(chunks) > 0:
        mx = max(y - x for x, y in chunks)
        br = torch.zeros(len(chunks), mx, 3)
        bs = torch.zeros(len(chunks), mx, 3)
        bw = torch.zeros(len(chunks), mx)
        for i, (x, y) in enumerate(chunks):
            br[i, :y - x], bs[i, :y - x], bw[i, :y - x] = ref_c[x:y], src_c[x:y], sc[x:y]
        Ts = weighted_procrustes(bs, br, bw)
        aligned = apply_transform(src_c.unsqueeze(0), Ts)
        res_all = torch.linalg.norm(ref_c.unsqueeze(0) - aligned, dim=2)
        inl = torch.lt(res_all, fm.acceptance_radius)
        best = inl.sum(dim=1).argmax()
        cur = sc * inl[best].float()
        margins = [float((res_all[best] - fm.acceptance_radius).abs().min())]
        if taps is not None:
            taps['patch_transforms'], taps['inlier_counts'], taps['best_index'] = Ts, inl.sum(dim=1), best
            # conditioning of every per-patch Kabsch problem (test diagnostics only): ratio of the two largest singular values
            # of the weighted, centred source points -- near 0 for (almost) collinear correspondences, whose rotation about
            # the line is undetermined and differs between SVD implementations by far more than rounding
            w = bw / (bw.sum(dim=1, keepdim=True) + 1e-5)
            cen = bs - (bs * w.unsqueeze(2)).sum(dim=1, keepdim=True)
            sv = torch.linalg.svdvals(cen * w.sqrt().unsqueeze(2))
            taps['patch_conditioning'] = sv[:, 1] / sv[:, 0].clamp(min=1e-12)
    else:
        T = weighted_procrustes(src_c, ref_c, sc)
        res = torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1)
        cur = sc * torch.lt(res, fm.acceptance_radius).float()
        margins = [float((res - fm.acceptance_radius).abs().min())]
    T = weighted_procrustes(src_c, ref_c, cur)
    for _ in range(fm.num_refinement_steps - 1):
        res = torch.linalg.norm(ref_c - apply_transform(src_c, T), dim=1)
        margins.append(float((res - fm.acceptance_radius).abs().min()))
        cur = sc * torch.lt(res, fm.acceptance_radius).float()
        T = weighted_procrustes(src_c, ref_c, cur)
    if taps is not None:
        # distance of the closest residual to the hard inlier threshold over the hypothesis test and every refinement
        # step: below float noise the selected inlier set (hence T) is not reproducible between implementations
        taps['threshold_margin'] = min(margins)
    return ref_c, src_c, sc, T


# ------------------------------------------------------------------------------------------------ assembly

def forward(sd, cfg, data, taps=None):
    """experiments/*/model.py:69-212 at inference; gt_node_corr_* (model.py:112-126) when data carries 'transform'."""
    out = {}
    fl = cfg.model.fine_level
    lens, pts = data['lengths'], data['points']
    nc, nf, n0 = int(lens[-1][0]), int(lens[fl][0]), int(lens[0][0])
    pc, pf, p0 = pts[-1], pts[fl], pts[0]
    ref_c, src_c, ref_f, src_f = pc[:nc], pc[nc:], pf[:nf], pf[nf:]
    out.update(ref_points_c=ref_c, src_points_c=src_c, ref_points_f=ref_f, src_points_f=src_f,
               ref_points=p0[:n0], src_points=p0[n0:])
    K = cfg.model.num_points_in_patch
    _, r_nm, r_knn, r_km = point_to_node_partition(ref_f, ref_c, K)
    _, s_nm, s_knn, s_km = point_to_node_partition(src_f, src_c, K)
    r_pad = torch.cat([ref_f, torch.zeros_like(ref_f[:1])], dim=0)
    s_pad = torch.cat([src_f, torch.zeros_like(src_f[:1])], dim=0)
    r_knn_pts, s_knn_pts = r_pad[r_knn], s_pad[s_knn]
    if taps is not None:
        taps.update(ref_node_masks=r_nm, src_node_masks=s_nm, ref_node_knn_indices=r_knn, src_node_knn_indices=s_knn,
                    ref_node_knn_masks=r_km, src_node_knn_masks=s_km)

    if data.get('transform') is not None:
        gi, go = get_node_correspondences(ref_c, src_c, r_knn_pts, s_knn_pts, torch.as_tensor(data['transform']),
                                          cfg.model.ground_truth_matching_radius, r_nm, s_nm, r_km, s_km)
        out.update(gt_node_corr_indices=gi, gt_node_corr_overlaps=go)

    feats_list = backbone(sd, cfg, data['features'], data, taps=taps)
    feats_c, feats_f = feats_list[-1], feats_list[0]
    if taps is not None:
        taps['feats_c'], taps['feats_f'] = feats_c, feats_f
    rf, sf = geometric_transformer(sd, cfg, ref_c, src_c, feats_c[:nc], feats_c[nc:], taps=taps)
    rfn, sfn = F.normalize(rf, p=2, dim=1), F.normalize(sf, p=2, dim=1)
    out.update(ref_feats_c=rfn, src_feats_c=sfn, ref_feats_f=feats_f[:nf], src_feats_f=feats_f[nf:])

    cmc = cfg.coarse_matching
    r_idx, s_idx, node_scores = superpoint_matching(rfn, sfn, r_nm, s_nm, cmc.num_correspondences, cmc.dual_normalization)
    out.update(ref_node_corr_indices=r_idx, src_node_corr_indices=s_idx, node_corr_scores=node_scores)

    rk_i, sk_i = r_knn[r_idx], s_knn[s_idx]
    rk_m, sk_m = r_km[r_idx], s_km[s_idx]
    rk_p, sk_p = r_knn_pts[r_idx], s_knn_pts[s_idx]
    rf_pad = torch.cat([feats_f[:nf], torch.zeros_like(feats_f[:1])], dim=0)
    sf_pad = torch.cat([feats_f[nf:], torch.zeros_like(feats_f[:1])], dim=0)
    ms = torch.einsum('bnd,bmd->bnm', rf_pad[rk_i], sf_pad[sk_i]) / feats_f.shape[1] ** 0.5
    if taps is not None:
        taps['matching_scores_raw'] = ms
    ms = optimal_transport(sd['optimal_transport.alpha'], ms, rk_m, sk_m, cfg.model.num_sinkhorn_iterations)
    out.update(ref_node_corr_knn_points=rk_p, src_node_corr_knn_points=sk_p, ref_node_corr_knn_masks=rk_m,
               src_node_corr_knn_masks=sk_m, matching_scores=ms)
    sm = ms if cfg.fine_matching.use_dustbin else ms[:, :-1, :-1]
    rc, sc_, cs, T = local_global_registration(cfg, rk_p, sk_p, rk_m, sk_m, sm, taps=taps)
    out.update(ref_corr_points=rc, src_corr_points=sc_, corr_scores=cs, estimated_transform=T)
    return out


def get_node_correspondences(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks,
                             src_masks, ref_knn_masks, src_knn_masks):
    """modules/registration/matching.py:231-315: GT superpoint pairs = patches sharing at least one point pair closer
    than pos_radius after the GT transform; overlap = mean of the two covered fractions."""
    src_nodes = apply_transform(src_nodes, transform)
    src_knn_points = apply_transform(src_knn_points, transform)
    node_ok = ref_masks[:, None] & src_masks[None, :]
    r_d = torch.linalg.norm(ref_knn_points - ref_nodes[:, None], dim=-1).masked_fill(~ref_knn_masks, 0.0).max(1)[0]
    s_d = torch.linalg.norm(src_knn_points - src_nodes[:, None], dim=-1).masked_fill(~src_knn_masks, 0.0).max(1)[0]
    centre = torch.sqrt(pairwise_distance(ref_nodes, src_nodes))
    cand = ((r_d[:, None] + s_d[None, :] + pos_radius - centre) > 0) & node_ok
    ri, si = torch.nonzero(cand, as_tuple=True)
    overlaps = torch.zeros(ri.shape[0])
    for lo in range(0, ri.shape[0], 4096):                       # bounded (B,K,K) chunks
        r, s_ = ri[lo:lo + 4096], si[lo:lo + 4096]
        rm, sm_ = ref_knn_masks[r], src_knn_masks[s_]
        d = pairwise_distance(ref_knn_points[r], src_knn_points[s_])
        d = d.masked_fill(~(rm[:, :, None] & sm_[:, None, :]), 1e12)
        hit = d < pos_radius ** 2
        rc = torch.count_nonzero(hit.sum(-1), dim=-1).float() / rm.sum(-1).float()
        sc = torch.count_nonzero(hit.sum(-2), dim=-1).float() / sm_.sum(-1).float()
        overlaps[lo:lo + 4096] = (rc + sc) / 2
    keep = overlaps > 0
    return torch.stack([ri[keep], si[keep]], dim=1), overlaps[keep]


def evaluate(cfg, out, transform):
    """experiments/<exp>/loss.py:95-159 Evaluator.forward (3DMatch / KITTI / ModelNet variants by cfg.name)."""
    e = cfg.eval
    T_gt = torch.as_tensor(transform)
    gi, go = out['gt_node_corr_indices'], out['gt_node_corr_overlaps']
    gi = gi[go > e.acceptance_overlap]
    gmap = torch.zeros(out['ref_points_c'].shape[0], out['src_points_c'].shape[0])
    gmap[gi[:, 0], gi[:, 1]] = 1.0
    res = {'PIR': gmap[out['ref_node_corr_indices'], out['src_node_corr_indices']].mean()}
    d = torch.linalg.norm(out['ref_corr_points'] - apply_transform(out['src_corr_points'], T_gt), dim=1)
    res['IR'] = (d < e.acceptance_radius).float().mean()
    T = out['estimated_transform']
    x = 0.5 * (torch.trace(T[:3, :3].T @ T_gt[:3, :3]) - 1.0)
    rre = 180.0 * torch.arccos(x.clamp(min=-1.0, max=1.0)) / np.pi
    rte = torch.linalg.norm(T_gt[:3, 3] - T[:3, 3])
    res['RRE'], res['RTE'] = rre, rte
    src = out['src_points']
    if cfg.name == '3dmatch':
        re = apply_transform(src, torch.inverse(T_gt) @ T)
        res['RMSE'] = torch.linalg.norm(re - src, dim=1).mean()
        res['RR'] = (res['RMSE'] < e.rmse_threshold).float()
    else:
        if cfg.name == 'modelnet':
            res['RMSE'] = torch.linalg.norm(apply_transform(src, T) - apply_transform(src, T_gt), dim=1).mean()
        res['RR'] = ((rre < e.rre_threshold) & (rte < e.rte_threshold)).float()
    return res


def registration_error(gt, est):
    """utils/registration.py:51 compute_registration_error / modules/registration/metrics.py isotropic error:
    RRE in degrees, RTE in the cloud's unit."""
    gt, est = np.asarray(gt, dtype=np.float64), np.asarray(est, dtype=np.float64)
    x = 0.5 * (np.trace(est[:3, :3].T @ gt[:3, :3]) - 1.0)
    rre = np.degrees(np.arccos(np.clip(x, -1.0, 1.0)))
    rte = np.linalg.norm(gt[:3, 3] - est[:3, 3])
    return float(rre), float(rte)
