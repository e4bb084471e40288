"""Regenerates tests/golden/*: small indexes BUILT BY THE REFERENCE and the answers THE REFERENCE
gives on them (oracle/_ref/libfemto_ref.so = unmodified femto compiled from /root/reference).

Run in the container where /root/reference is mounted:
    python tests/golden/make_golden.py
The fixtures let the oracle be pinned on machines where the reference does not exist.
Each case directory holds:  index/{00,01,...}  (the reference builder's bytes) and expected.json.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import corpus  # noqa: E402
from oracle.bindings import Reference  # noqa: E402

CASES = {
    # the reference's own golden fixture (src/main/index_test.c:514-533)
    "two_docs": ([b"test_one;", b"test_two_fun;"], dict(mark_period=100)),
    # multi-block, tiny buckets (src/main/index_test_funcs.c:46-86 "small blocks")
    "gen60_small_blocks": ([corpus.generate_text(60)], dict(block_size=16, bucket_size=4, chunk_size=-1)),
    # several documents, all byte values, RLE-friendly and random regions
    "mixed_1500": ([b"a" * 200, corpus.all_bytes_doc(), corpus.random_acgt(500, 9), b"", corpus.generate_text(250)],
                   dict(block_size=512, bucket_size=128, chunk_size=-1, mark_period=6)),
}


def main():
    for name, (docs, params) in CASES.items():
        base = os.path.join(HERE, name)
        shutil.rmtree(base, ignore_errors=True)
        os.makedirs(base)
        idx = os.path.join(base, "index")
        with tempfile.TemporaryDirectory() as tmp:
            Reference.build_index(docs, idx, tmp, **params)
        os.remove(os.path.join(idx, "_femto_index")) if os.path.exists(os.path.join(idx, "_femto_index")) else None
        pats = corpus.sample_patterns(docs, 60, [1, 2, 3, 4, 6, 9], seed=77)
        pats += [np.zeros(0, dtype=np.uint16), np.array([2], dtype=np.uint16)]
        with Reference(idx) as r:
            n = r.header_info()["total_length"]
            f, l = r.count(pats)
            max_occs = 5
            loc = r.locate(pats, max_occs)
            steps = [list(r.back_step(row)) for row in range(n)]
            rng = np.random.default_rng(1)
            occ = [[int(c), int(row), r.occ(int(c), int(row))[0]]
                   for c, row in zip(rng.integers(0, 261, 400), rng.integers(0, n, 400))]
            sa = r.locate_range(0, n - 1).tolist()
            # generic requests (femto.h:75-149): the pattern as byte values
            reqs = {}
            for p in [q for q in pats if len(q) and (q >= 5).all()][:6]:
                body = " ".join(str(int(x) - 5) for x in p)
                for word in ("string_rows", "string_rows_left", "string_rows_right", "string_rows_all"):
                    reqs[f"{word} {body}"] = r.generic_request(idx, f"{word} {body}")
        json.dump({
            "docs_hex": [d.hex() for d in docs], "params": params, "total_length": n,
            "patterns": [p.tolist() for p in pats], "first": f.tolist(), "last": l.tolist(),
            "max_occs": max_occs, "locate": [x.tolist() for x in loc], "back_step": steps,
            "occ_samples": occ, "sa": sa, "generic_requests": reqs,
        }, open(os.path.join(base, "expected.json"), "w"))
        size = sum(os.path.getsize(os.path.join(idx, f)) for f in os.listdir(idx))
        print(f"{name}: n={n} index bytes={size}")


if __name__ == "__main__":
    main()
