"""femto_search's reports for literal patterns (integration/femto_search_format.h, the printing half of
integration/femto_search_b200.c) against what the REFERENCE tool prints.

On CPU the results come from the oracle (count, locate_range, resolve) and go through the same formatting
code by way of the harness tests/search_format_check.c; the expected bytes are the committed outputs of the
reference's femto_search (tests/golden/search_tool, generator make_search_golden.py) and, where the
reference is built here, its live output -- including a pattern with more than 1 Mi occurrences, of which one
query reports the first 1 Mi rows.  The tool's own main() is replayed too, linked with a stand-in engine that
answers from the oracle.  The GPU test of the real tool is tests/test_gpu_zz_search_tool.py.
"""
import json
import os
import subprocess

import numpy as np
import pytest

import corpus
import femto_b200 as fb
from conftest import GOLDEN_DIR
from oracle.bindings import REF_SO, Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = os.path.join(GOLDEN_DIR, "search_tool")
REF_TOOL = os.path.join(os.path.dirname(REF_SO), "femto_search")
ROWS_PER_QUERY = 1 << 20


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fmt") / "search_format_check")
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-o", exe,
                    os.path.join(ROOT, "tests", "search_format_check.c")], check=True)
    return exe


def oracle_results(index, infos, pat, want_offsets):
    """(first, last, [(info, [offsets])...]) of one index for the harness."""
    with Oracle(index) as o:
        f, l = o.count([pat])
        first, last = int(f[0]), int(l[0])
        docs = {}
        if last >= first:
            top = min(last, first + ROWS_PER_QUERY - 1)
            for off in o.locate_range(first, top):
                d, k = o.resolve(int(off))
                docs.setdefault(d, []).append(k)
    out = [(infos[d], sorted(v) if want_offsets else []) for d, v in sorted(docs.items())]
    return first, last, out


def run_harness(harness, opts, pat, per_index):
    count, offsets, js, sep = opts
    words = [count, offsets, js, sep, len(pat), *pat.tolist(), len(per_index)]
    for first, last, docs in per_index:
        words += [first, last, len(docs)]
        for info, offs in docs:
            words += [len(info), *info, len(offs), *offs]
    out = subprocess.run([harness], input=" ".join(str(int(w)) for w in words).encode(), capture_output=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return out.stdout


def options_of(case_options):
    return ("--count" in case_options, "--offsets" in case_options, "--json" in case_options,
            0 if "--null" in case_options and "--json" not in case_options else 10)


def test_reports_equal_the_reference_tools_golden_output(harness):
    exp = json.load(open(os.path.join(BASE, "search_expected.json")))
    infos = {k: [bytes.fromhex(h) for h in v["infos_hex"]] for k, v in exp["indexes"].items()}
    seen = set()
    for case in exp["cases"]:
        pat = corpus.to_alpha(bytes.fromhex(case["pattern_hex"]))
        opts = options_of(case["options"])
        per_index = [oracle_results(os.path.join(BASE, name), infos[name], pat, opts[1]) for name in case["indexes"]]
        got = run_harness(harness, opts, pat, per_index)
        assert got == bytes.fromhex(case["stdout_hex"]), (case["indexes"], case["pattern_hex"], case["options"])
        seen.add((len(case["indexes"]), tuple(case["options"])))
    assert len(exp["cases"]) >= 150 and len(seen) == 18


def test_the_tools_own_main_replays_the_golden_reports(tmp_path):
    """integration/femto_search_b200.c itself -- option parsing, the queries it makes, sorting and grouping, the
    report -- linked with a stand-in engine that answers its fm_* calls from the oracle
    (tests/search_tool_oracle_engine.c; test infrastructure): every golden command line, byte for byte."""
    exe = str(tmp_path / "femto_search_oracle_engine")
    subprocess.run(["gcc", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-fsanitize=address,undefined",
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "integration"), "-o", exe,
                    os.path.join(ROOT, "integration", "femto_search_b200.c"),
                    os.path.join(ROOT, "tests", "search_tool_oracle_engine.c"),
                    os.path.join(ROOT, "oracle", "fm_oracle.c")], check=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")   # a command-line tool: results are freed by exit
    exp = json.load(open(os.path.join(BASE, "search_expected.json")))
    pf = tmp_path / "pattern.bin"
    for case in exp["cases"]:
        pf.write_bytes(bytes.fromhex(case["pattern_hex"]))
        args = [os.path.join(BASE, n) for n in case["indexes"]] + ["--raw-pattern-from", str(pf)] + case["options"]
        out = subprocess.run([exe] + args, capture_output=True, timeout=60, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        assert out.stdout == bytes.fromhex(case["stdout_hex"]), (case["indexes"], case["pattern_hex"], case["options"])
    # the pattern on the command line, and the report into a file
    dest = tmp_path / "report.txt"
    subprocess.run([exe, os.path.join(BASE, "index1"), "--offsets", "--raw-pattern", "ana", "--output", str(dest)],
                   check=True, env=env, timeout=60)
    want = [c for c in exp["cases"] if c["indexes"] == ["index1"] and c["pattern_hex"] == b"ana".hex()
            and c["options"] == ["--offsets"]][0]
    assert dest.read_bytes() == bytes.fromhex(want["stdout_hex"])


def test_oracle_document_names_are_the_stored_info_strings():
    exp = json.load(open(os.path.join(BASE, "search_expected.json")))
    import ctypes as C
    for name, v in exp["indexes"].items():
        with Oracle(os.path.join(BASE, name)) as o:
            for d, h in enumerate(v["infos_hex"]):
                p, n = C.POINTER(C.c_ubyte)(), C.c_int64()
                assert o.lib.fmo_doc_name(o.h, C.c_int64(d), C.byref(p), C.byref(n)) == 0
                assert bytes(p[:n.value]) == bytes.fromhex(h)
            p, n = C.POINTER(C.c_ubyte)(), C.c_int64()
            assert o.lib.fmo_doc_name(o.h, C.c_int64(len(v["infos_hex"])), C.byref(p), C.byref(n)) != 0


def test_golden_indexes_are_what_the_documents_give(tmp_path):
    """The committed search_tool indexes are this emitter's bytes for the documents in search_expected.json."""
    exp = json.load(open(os.path.join(BASE, "search_expected.json")))
    for name, v in exp["indexes"].items():
        d = str(tmp_path / name)
        fb.build_index_host([bytes.fromhex(h) for h in v["docs_hex"]], d,
                            doc_infos=[bytes.fromhex(h) for h in v["infos_hex"]], **exp["params"])
        for f in os.listdir(os.path.join(BASE, name)):
            assert open(os.path.join(d, f), "rb").read() == open(os.path.join(BASE, name, f), "rb").read(), (name, f)


@pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="oracle/_ref/femto_search not built (no /root/reference here)")
def test_live_reference_tool_and_the_one_mi_row_limit(harness, tmp_path):
    """A pattern with 1.2 Mi occurrences: one femto_search query reports the first 2^20 rows of its range."""
    rng = np.random.default_rng(4)
    docs = [bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), 1 << 20)) for _ in range(2)] + [b"a" * 300000]
    infos = [b"d0", b"d1", b"d2"]
    idx = str(tmp_path / "big")
    fb.build_index_host(docs, idx, doc_infos=infos, mark_period=2)
    pat = corpus.to_alpha(b"a")
    for options in (["--offsets"], ["--count"], [], ["--offsets", "--json"]):
        opts = options_of(options)
        want = subprocess.run([REF_TOOL, idx, "--raw-pattern", "a", *options], capture_output=True, timeout=600)
        assert want.returncode == 0, want.stderr
        first, last, found = oracle_results(idx, infos, pat, opts[1])
        assert last - first + 1 > ROWS_PER_QUERY
        got = run_harness(harness, opts, pat, [(first, last, found)])
        assert got == want.stdout, options
        if options == ["--offsets"]:
            assert sum(len(o) for _, o in found) == ROWS_PER_QUERY


def test_tool_refuses_what_needs_the_parser_and_fails_loudly_without_a_gpu():
    tool = os.path.join(os.path.dirname(REF_SO), "femto_search_b200")
    if not os.path.exists(tool):
        pytest.skip("femto_search_b200 not built")
    idx = os.path.join(BASE, "index1")
    r = subprocess.run([tool, idx, "banana"], capture_output=True, text=True)
    assert r.returncode != 0 and "--raw-pattern" in r.stderr
    r = subprocess.run([tool, idx, "--matches", "--raw-pattern", "a"], capture_output=True, text=True)
    assert r.returncode != 0 and "parser" in r.stderr
    r = subprocess.run([tool, idx, "--bogus"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unknown option --bogus" in r.stdout
    r = subprocess.run([tool, "/nonexistent", "--raw-pattern", "a"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not open index at /nonexistent" in r.stdout
    import torch
    if not torch.cuda.is_available():     # no CPU query path behind the tool
        r = subprocess.run([tool, idx, "--raw-pattern", "a", "--count"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU query path" in r.stderr
