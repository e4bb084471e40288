"""The drop-in boundary: libfemto_b200.so loads, exports every symbol include/femto_b200.h declares,
and refuses to answer queries without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import femto_b200 as fb
from femto_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "femto_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 25
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in femto_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in femto_b200/_lib.py"
    assert sorted(_lib.PROTOTYPES) == names


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_oracle_in_product_path():
    """Nothing under femto_b200/ or integration/ (the tools and shims a maintainer links) may reference oracle/:
    the oracle is test infrastructure."""
    trees = list(os.walk(os.path.join(ROOT, "femto_b200"))) + list(os.walk(os.path.join(ROOT, "integration")))
    for base, _, files in trees:
        if "build" in base.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".c", ".cc", ".cu", ".hpp", ".cuh", ".h")):
                src = open(os.path.join(base, f), errors="replace").read()
                assert "fm_oracle" not in src and "libfemto_ref" not in src, f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="this check is for hosts without a GPU")
def test_open_fails_loudly_without_gpu(built_indexes):
    with pytest.raises(fb.FemtoError) as e:
        fb.Index(built_indexes["two_docs"])
    assert e.value.code == 2                              # ERR_IO
    assert "no CPU query path" in str(e.value)


def test_error_codes_follow_reference_numbering():
    # src/utils/error.h:25-39
    assert fb.ERR_NAMES[1] == "MEM" and fb.ERR_NAMES[2] == "IO" and fb.ERR_NAMES[3] == "PARAM"
    assert fb.ERR_NAMES[4] == "FORMAT" and fb.ERR_NAMES[6] == "INVALID" and fb.ERR_NAMES[10] == "FULL"
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.fm_open(None, 0, C.byref(h)) == 3          # PARAM for a null path


def test_host_suffix_sort_matches_naive():
    import numpy as np
    rng = np.random.default_rng(0)
    for docs in ([b"banana"], [b"aaaa", b"aaa"], [bytes(rng.integers(97, 100, 300, dtype=np.uint8))], [b"", b"", b"x"]):
        text, ends = fb.prepare_text(docs)
        sa = fb.suffix_sort_host(text)
        t = text.tolist()
        naive = sorted(range(len(t)), key=lambda i: t[i:])
        assert sa.tolist() == naive
