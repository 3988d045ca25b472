"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/mulactseg_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mulactseg_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mulactseg_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mulactseg_b200 import build
        build.build(verbose=False)
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in mulactseg_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.mas_abi_version() == 1


def test_argument_validation_needs_no_gpu():
    from mulactseg_b200 import _lib
    lib = _lib.load()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.mas_bvsb_segment_stats_dev(None, 0, 0, None, 1, 20, 8, 8, 4, 1.0, None, None, None, None) == -1
    assert b"null" in lib.mas_last_error()
    assert lib.mas_topk_u64_dev(None, 0, 0, None, None, None, 0, None) == -1
    assert lib.mas_sort_capacity(100001) == 131072
    assert lib.mas_topk_workspace_bytes() > 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mulactseg_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MulActSegError):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import torch
    from mulactseg_b200 import ops
    with pytest.raises(RuntimeError):
        ops.minmax_nonzero(torch.zeros(4))
    with pytest.raises(RuntimeError):
        ops.region_scores(torch.zeros(2, 3), torch.zeros(2, 3, dtype=torch.int32))
