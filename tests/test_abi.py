"""CPU: the C-ABI shared library builds/loads here and exports exactly what include/gava_sm100.h declares.
No compute call is made (no GPU in the build container)."""
import ctypes
import re

from tests.util import ROOT


def declared_symbols():
    text = open(f"{ROOT}/include/gava_sm100.h").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gava_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from gava_clip_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gava_sm100.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in gava_clip_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.gava_version() == 4


def test_epilogue_struct_layout():
    from gava_clip_b200._lib import GemmEpilogue
    # int32 mode (+4 pad), 2 ptr, i64, ptr, i64, 4 x i32, 2 ptr, ptr, i64, ptr, i32 (+4 pad), then the LayerNorm-fold block:
    # ptr, i64, 2 ptr, i64, ptr, i32, f32 -> 168 bytes on LP64 (api.cu static_asserts the same offsets on the C side)
    assert ctypes.sizeof(GemmEpilogue) == 168 and GemmEpilogue.aux.offset == 80 and GemmEpilogue.aux_out.offset == 96
    assert GemmEpilogue.max_ctas.offset == 104 and GemmEpilogue.copy_out.offset == 112 and GemmEpilogue.stats_in.offset == 136
    assert GemmEpilogue.colsum.offset == 152 and GemmEpilogue.n_slots.offset == 160 and GemmEpilogue.eps.offset == 164
    assert GemmEpilogue.out.offset == 32 and GemmEpilogue.group.offset == 48 and GemmEpilogue.pos.offset == 64


def test_no_oracle_in_product():
    """The product package must never import the oracle (it is test infrastructure)."""
    import pathlib
    for p in pathlib.Path(f"{ROOT}/gava_clip_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_workspace_bytes_matches_design_numbers():
    """gava_workspace_bytes is pure host arithmetic: config 2 (B=64, T=8, ViT-B/16) needs ~1.9 GB for inference (+ clip and weights) and
    ~1.9 GB of tape per block for a fine-tune step."""
    from gava_clip_b200 import _lib
    lib = _lib.load()
    infer = lib.gava_workspace_bytes(64, 8, 197, 768, 3072, 8, 768, 512, 0)
    train = lib.gava_workspace_bytes(64, 8, 197, 768, 3072, 8, 768, 512, 12)
    assert 1.7e9 < infer < 2.1e9
    assert 20e9 < train - infer < 30e9
    assert lib.gava_workspace_bytes(0, 8, 197, 768, 3072, 8, 768, 512, 0) == -1
