"""Self-attention micro-benchmark: a batch of clouds (N superpoints each, C = 256, H = 4) through geob200_attention_batched,
TMA-staged path vs the lanes<->channels cp.async path.  CUDA events, warm.  Dev tool.
    python tools/att_bench.py [clouds] [N] [reps]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from geotransformer_b200 import _lib as L

NC = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 320
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 20
C, H = 256, 4


class Item(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('q', 'k', 'v', 'qp', 'qb', 'embed', 'out')] + [('n_query', ctypes.c_int64), ('n_key', ctypes.c_int64)]


dev = torch.device('cuda')
rows = NC * N
qkv = torch.randn(rows, 3 * C, device=dev)
qp = torch.randn(rows, H, C, device=dev) * 0.2
qb = torch.randn(rows, H, device=dev)
E = torch.randn(NC, N, N, C, device=dev)
out = torch.empty(rows, C, device=dev)
items = (Item * NC)()
for c in range(NC):
    o = c * N
    items[c] = Item(qkv[o:].data_ptr(), qkv[o:, C:].data_ptr(), qkv[o:, 2 * C:].data_ptr(), qp[o:].data_ptr(), qb[o:].data_ptr(), E[c].data_ptr(),
                    out[o:].data_ptr(), N, N)
lib = L.lib()
lib.geob200_attention_batched_workspace_bytes.restype = ctypes.c_size_t
ws = torch.empty(lib.geob200_attention_batched_workspace_bytes(items, NC, H) + 1024, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
res = {}
for tma in (0, 1):
    lib.geob200_set_attention_tma(tma)
    for _ in range(3):
        L.check(lib.geob200_attention_batched(items, NC, 3 * C, 3 * C, 3 * C, C, C, H, ws.data_ptr(), ws.numel(), st), 'att')
    ts = []
    for _ in range(REPS):
        flush.zero_()                           # E (NC x 105 MB) exceeds the L2 anyway; flush for the small operands too
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.geob200_attention_batched(items, NC, 3 * C, 3 * C, 3 * C, C, C, H, ws.data_ptr(), ws.numel(), st), 'att')
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gb = NC * N * N * C * 4 / 1e9
    res[tma] = out.clone()
    print(f'tma={tma}: {ms:.3f} ms per layer for {NC} clouds of {N} superpoints: E stream {gb:.2f} GB -> {gb / ms:.2f} TB/s (whole attention, 3 or 2 launches)')
print('max |tma - cp.async| =', float((res[0] - res[1]).abs().max()))
