import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


@pytest.fixture(scope="session")
def eng():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("jsa-rag_b200")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test session (built in-tree; no fallback)."""
    build = importlib.import_module("jsa-rag_b200.build")
    build.build()


def load_golden(name):
    """Loads a golden case; regenerates seeded inputs when they are not stored and checks their checksum."""
    from oracle import make_golden
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    if "embeddings" not in g:
        e16, q = make_golden.synth_inputs(int(g["n"]), int(g["d"]), int(g["b"]), int(g["seed"]), int(g["dup_rows"]))
        assert make_golden.checksum(e16, q) == str(g["input_sha256"]), "regenerated inputs differ from the golden run"
        g["embeddings"], g["queries"] = e16, q
    return g


GOLDEN_CASES = ["flat_n1003_d768_b8_k20", "flat_n300_d1024_b5_k10", "flat_dups_n640_d768_b6_k100",
                "flat_n20000_d768_b64_k100", "flat_n100000_d768_b256_k20"]
