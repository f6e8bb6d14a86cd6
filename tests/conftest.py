import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenCase:
    def __init__(self, name):
        import datasets
        import torch

        self.name = name
        self.dir = os.path.join(GOLDEN, name)
        with open(os.path.join(self.dir, "meta.json")) as fr:
            self.meta = json.load(fr)
        self.X = datasets.case_docs(name)
        assert datasets.sha256(self.X) == self.meta["x_sha256"], "synthetic docs did not regenerate bit-identically"
        self.Q = datasets.make_queries(self.meta["d"])
        assert datasets.sha256(self.Q) == self.meta["q_sha256"]
        self.codebook_param = torch.load(os.path.join(self.dir, "codebook.pt"), map_location="cpu", weights_only=False)
        self.codebook = self.codebook_param.detach().numpy().copy()
        self.codes = np.load(os.path.join(self.dir, "codes.npy"))
        self.codes_ip = np.load(os.path.join(self.dir, "codes_ip.npy"))
        self.last_preds = np.load(os.path.join(self.dir, "last_preds.npy"))
        self.M, self.K, self.d, self.n = self.meta["M"], 2 ** self.meta["bits"], self.meta["d"], self.meta["n"]

    def load(self, fname):
        return np.load(os.path.join(self.dir, fname))

    def pickle(self, fname):
        import pickle

        with open(os.path.join(self.dir, fname), "rb") as fr:
            return pickle.load(fr)


_cases = {}


def golden_case(name):
    if name not in _cases:
        _cases[name] = GoldenCase(name)
    return _cases[name]


CASE_NAMES = ["gauss768", "mix768", "small64"]


@pytest.fixture(params=CASE_NAMES)
def case(request):
    return golden_case(request.param)


@pytest.fixture
def gauss():
    return golden_case("gauss768")
