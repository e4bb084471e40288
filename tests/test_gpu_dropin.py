"""Drop-in at the tool level: the reference's UNMODIFIED batch driver femto_multiquery
(src/main/query_tool.c) linked against libfemto_b200.so through integration/femto_b200_shim.c
must report the same totals as the stock femto_multiquery on the same index and pattern file
(Pizza&Chili format, query_tool.c:48-98)."""
import os
import re
import subprocess

import pytest

import corpus
from oracle.bindings import REF_SO

pytestmark = pytest.mark.gpu

REF_DIR = os.path.dirname(REF_SO)
STOCK = os.path.join(REF_DIR, "femto_multiquery")
DROPIN = os.path.join(REF_DIR, "femto_multiquery_b200")


def _run(tool, index, mode, pattern_file, extra=()):
    with open(pattern_file, "rb") as f:
        out = subprocess.run([tool, index, mode, *extra], stdin=f, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


@pytest.mark.skipif(not (os.path.exists(STOCK) and os.path.exists(DROPIN)),
                    reason="oracle/_ref tools did not travel (make -C oracle dropin)")
@pytest.mark.parametrize("name,length", [("english_100k", 4), ("acgt_64k", 8), ("bytes_200k", 2)])
def test_femto_multiquery_runs_on_the_gpu_engine(name, length, built_indexes, corpora, tmp_path):
    docs, _ = corpora[name]
    pats = corpus.sample_patterns(docs, 500, [length], seed=7, random_fraction=0.3)
    pf = tmp_path / "patterns.pc"
    with open(pf, "wb") as f:
        f.write(f"# number={len(pats)} length={length} file=synthetic forbidden=\n".encode())
        for p in pats:
            f.write(bytes((p - 5).astype("uint8")))
    index = built_indexes[name]

    stock = _run(STOCK, index, "-count", pf)
    ours = _run(DROPIN, index, "-count", pf)
    counted = lambda s: int(re.search(r"Counted (\d+) results", s).group(1))
    assert counted(stock) == counted(ours) > 0

    located = lambda s: float(re.search(r"Did ([\d.]+) parallel locate results", s).group(1))
    for max_occs in ("3", "100000"):
        assert located(_run(STOCK, index, "-locate", pf, (max_occs,))) == \
               located(_run(DROPIN, index, "-locate", pf, (max_occs,)))

    chunk = lambda s: float(re.search(r"Did ([\d.]+) chunk locate results", s).group(1))
    assert chunk(_run(STOCK, index, "-chunklocate", pf, ("50",))) == chunk(_run(DROPIN, index, "-chunklocate", pf, ("50",)))


REQUEST_TOOL = os.path.join(REF_DIR, "femto_request_b200")


@pytest.mark.skipif(not os.path.exists(REQUEST_TOOL), reason="oracle/_ref tools did not travel (make -C oracle dropin)")
def test_femto_handle_request_tool_on_the_gpu_engine():
    """integration/femto_request_b200.c = the reference's femto_handle_request (src/main/handle_request.c) over
    fm_generic_request: its Response section must be the reference's own answer (golden fixtures)."""
    import json
    from conftest import GOLDEN_DIR
    base = os.path.join(GOLDEN_DIR, "mixed_1500")
    exp = json.load(open(os.path.join(base, "expected.json")))
    for req, want in list(exp["generic_requests"].items())[:6]:
        out = subprocess.run([REQUEST_TOOL, os.path.join(base, "index"), req], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-2000:]
        head, resp = out.stdout.split("Response:\n", 1)
        assert head == f"Index:{os.path.join(base, 'index')}\nRequest:\n{req}\n"
        assert resp == want + "\n"
