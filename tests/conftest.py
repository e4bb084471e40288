import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library and the checkers exist (no-ops when already built)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def corpora():
    import corpus
    return corpus.standard_corpora()


@pytest.fixture(scope="session")
def built_indexes(corpora, tmp_path_factory):
    """name -> path of an index written by OUR builder (host suffix sort + emitter)."""
    import femto_b200 as fb
    base = tmp_path_factory.mktemp("indexes")
    out = {}
    for name, (docs, params) in corpora.items():
        path = str(base / name)
        fb.build_index_host(docs, path, **params)
        out[name] = path
    return out


@pytest.fixture(scope="session")
def have_ref():
    from oracle import bindings
    return bindings.have_reference()


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
