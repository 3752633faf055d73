"""The C-ABI library loads and exports every symbol include/geob200.h declares, with matching argument counts
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from geotransformer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    hdr = open(os.path.join(ROOT, 'include', 'geob200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    hdr = re.sub(r'typedef struct \{.*?\} \w+;', '', hdr, flags=re.S)
    return re.findall(r'\b(?:int|void|size_t|uint64_t|int64_t|const char\*)\s+(geob200_\w+)\s*\(([^;]*?)\)\s*;', hdr, flags=re.S)


def test_library_exports_every_declared_symbol():
    decls = _header_decls()
    assert len(decls) >= 30
    lib = ctypes.CDLL(L.LIB_PATH)
    for name, args in decls:
        assert hasattr(lib, name), f'{name} declared in geob200.h but not exported'
        n = 0 if args.strip() == 'void' else len(args.split(','))
        assert name in L.SIGNATURES, f'{name} has no ctypes signature'
        assert len(L.SIGNATURES[name][1]) == n, f'{name}: header has {n} args, binding {len(L.SIGNATURES[name][1])}'
    assert set(L.SIGNATURES) == {d[0] for d in decls}


def test_no_torch_types_in_the_abi():
    """signatures use plain pointers and sizes only (comments may cite torch/ATen call sites of the reference)"""
    for name, args in _header_decls():
        for tok in ('at::', 'Tensor', 'torch', 'c10', 'std::'):
            assert tok not in args, f'{name}: {tok} in signature'
    assert 'extern "C"' in open(os.path.join(ROOT, 'include', 'geob200.h')).read()


def test_pure_host_queries_work_without_gpu():
    lib = L.lib()
    assert lib.geob200_grid_subsample_workspace_bytes(40000, 2) > 40000 * 40
    assert lib.geob200_radius_search_workspace_bytes(40000, 40000, 2) > 40000 * 20
    assert lib.geob200_lgr_workspace_bytes(256, 64, 3) > 0
    assert lib.geob200_launch_count() == 0


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under geotransformer_b200/ may reference it"""
    pkg = os.path.join(ROOT, 'geotransformer_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'liboracle' not in src and 'libref_ext' not in src, f
