"""Regenerates tests/golden/search_tool/: what the REFERENCE's femto_search prints for literal patterns
(--raw-pattern) in its count / documents / --offsets reports, plain, --json and --null, on one and on two
indexes.  oracle/_ref/femto_search = src/main_cc/search_tool.cc compiled unmodified (make -C oracle femto_search).

    python tests/golden/make_search_golden.py

The two small indexes carry document info strings (names with a quote, a '|' and a non-ASCII byte) and
document chunks; they are written by this repository's emitter, whose files are byte-identical to the
reference builder's (tests/test_builder_format.py) -- the reference's in-memory test builder cannot set
info strings.  search_expected.json holds the documents, the parameters and every command line with the tool's
stdout, so the cases can be replayed anywhere.
"""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import femto_b200 as fb  # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "femto_search")
PARAMS = dict(block_size=64, bucket_size=32, chunk_size=8, mark_period=4)
INDEXES = {
    "index1": ([b"abracadabra banana bandana", b"the banana band played abracadabra", b"", b"nothing here",
                b"banana" * 5, b"ana"],
               [b"doc/one.txt", b"two", b"empty", b"four|glom", b"five \"q\"", b"six\xe9"]),
    "index2": ([b"second index banana", b"ana banana", b"a\"b\\c"], [b"s1", b"s2|s2b", b"s3"]),
}
PATTERNS = [b"ana", b"banana", b"a", b" ", b"zzz", b"abracadabra", b"a\"b\\", b"here", b"d p"]
MODES = [[], ["--offsets"], ["--count"]]
STYLES = [[], ["--json"], ["--null"]]


def main():
    base = os.path.join(HERE, "search_tool")
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base)
    for name, (docs, infos) in INDEXES.items():
        fb.build_index_host(docs, os.path.join(base, name), doc_infos=infos, **PARAMS)
        os.remove(os.path.join(base, name, "_femto_index"))
    cases = []
    for which in (["index1"], ["index1", "index2"]):
        for pat in PATTERNS:
            for mode in MODES:
                for style in STYLES:
                    args = ["{%s}" % w for w in which] + ["--raw-pattern-hex", pat.hex()] + mode + style
                    real = [os.path.join(base, w) for w in which] + ["--raw-pattern", pat.decode("latin-1")] + mode + style
                    out = subprocess.run([TOOL] + real, capture_output=True, timeout=60)
                    assert out.returncode == 0, (real, out.stderr)
                    cases.append({"indexes": which, "pattern_hex": pat.hex(), "options": mode + style,
                                  "stdout_hex": out.stdout.hex()})
    json.dump({"params": PARAMS,
               "indexes": {k: {"docs_hex": [d.hex() for d in v[0]], "infos_hex": [i.hex() for i in v[1]]}
                           for k, v in INDEXES.items()},
               "cases": cases}, open(os.path.join(base, "search_expected.json"), "w"), indent=0)
    print(f"search_tool: {len(cases)} cases")


if __name__ == "__main__":
    main()
