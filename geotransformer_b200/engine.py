"""Throughput front end: registers a batch of point-cloud pairs, several pairs in flight at once.

Pairs are independent.  Most kernels of one pair leave SMs idle (a few hundred CTAs, deep levels
with ~600 points), so the engine runs ``num_streams`` pairs concurrently, each on its own CUDA stream driven by its own
host thread (ctypes and torch release the GIL while launching / synchronising).  Per pair it performs: H2D of the raw
clouds (pinned staging) -> GPU collate -> model forward -> D2H of the estimated transform.  This is the caller-facing API
bench.py measures as `e2e` (the SingleTester-compatible loop, is built on it).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from .utils.data import registration_collate_fn_stack_mode


class RegistrationEngine:
    def __init__(self, model, cfg, neighbor_limits, num_streams=4, device=None):
        self.model, self.cfg, self.limits = model, cfg, neighbor_limits
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_streams = num_streams
        self.streams = [torch.cuda.Stream(self.device) for _ in range(num_streams)]
        self.pool = ThreadPoolExecutor(max_workers=num_streams)
        self.t_host = [torch.empty((4, 4), dtype=torch.float32).pin_memory() for _ in range(num_streams)]

    def _one(self, slot, pair, start_event, keep):
        torch.cuda.set_device(self.device)
        stream = self.streams[slot]
        with torch.cuda.stream(stream), _lib.stream_scope(stream.cuda_stream):
            if start_event is not None:
                stream.wait_event(start_event)
            b = self.cfg.backbone
            data = registration_collate_fn_stack_mode([pair], b.num_stages, b.init_voxel_size, b.init_radius, self.limits,
                                                      device=self.device)
            out = self.model(data)
            self.t_host[slot].copy_(out['estimated_transform'], non_blocking=True)
            done = torch.cuda.Event()
            done.record(stream)
            done.synchronize()                      # this thread only; the other streams keep running
            res = {'estimated_transform': self.t_host[slot].clone(), 'num_corr': int(out['ref_corr_points'].shape[0]),
                   'num_superpoints': (int(out['ref_points_c'].shape[0]), int(out['src_points_c'].shape[0]))}
            if keep:
                res['output_dict'] = out
        return res, done

    def register(self, pairs, start_event=None, keep_outputs=False):
        """pairs: list of dicts with ref_points/src_points/ref_feats/src_feats/transform (numpy, CPU or CUDA tensors).
        Returns one result dict per pair, in order."""
        results = [None] * len(pairs)
        for base in range(0, len(pairs), self.num_streams):
            chunk = pairs[base:base + self.num_streams]
            futs = [self.pool.submit(self._one, s, p, start_event, keep_outputs) for s, p in enumerate(chunk)]
            for i, f in enumerate(futs):
                res, done = f.result()
                torch.cuda.current_stream().wait_event(done)
                results[base + i] = res
        return results

    def close(self):
        self.pool.shutdown()
