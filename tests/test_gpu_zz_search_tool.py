"""integration/femto_search_b200.c on the GPU engine must print what the reference's femto_search prints for
literal patterns: the committed outputs of the reference tool (tests/golden/search_tool) are replayed, one and
two indexes, count / documents / --offsets, plain / --json / --null.  (Runs last among the GPU tests.)"""
import json
import os
import subprocess

import pytest

from conftest import GOLDEN_DIR
from oracle.bindings import REF_SO

pytestmark = pytest.mark.gpu

TOOL = os.path.join(os.path.dirname(REF_SO), "femto_search_b200")
BASE = os.path.join(GOLDEN_DIR, "search_tool")


@pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref tools did not travel (make -C oracle dropin)")
def test_femto_search_reports_on_the_gpu_engine(tmp_path):
    exp = json.load(open(os.path.join(BASE, "search_expected.json")))
    cases = exp["cases"]
    picked = cases[::4] + [c for c in cases if c["pattern_hex"] == b"ana".hex()]
    combos = set()
    for case in picked:
        pf = tmp_path / "pattern.bin"
        pf.write_bytes(bytes.fromhex(case["pattern_hex"]))
        args = [os.path.join(BASE, n) for n in case["indexes"]] + ["--raw-pattern-from", str(pf)] + case["options"]
        out = subprocess.run([TOOL] + args, capture_output=True, timeout=120)
        assert out.returncode == 0, out.stderr[-2000:]
        assert out.stdout == bytes.fromhex(case["stdout_hex"]), (case["indexes"], case["pattern_hex"], case["options"])
        combos.add((len(case["indexes"]), tuple(case["options"])))
    assert len(combos) == 18
    # --output writes the same report to a file
    first = picked[0]
    dest = tmp_path / "report.txt"
    subprocess.run([TOOL, os.path.join(BASE, first["indexes"][0]), "--raw-pattern",
                    bytes.fromhex(first["pattern_hex"]).decode("latin-1"), "--output", str(dest)] + first["options"],
                   check=True, timeout=120)
    assert dest.read_bytes() == bytes.fromhex(first["stdout_hex"])
