"""Loader for the committed golden fixtures (tests/golden/*.json)."""
import json
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(_DIR, name + ".json")) as f:
        return json.load(f)["tables"]


def tab(g, name):
    t = g[name]
    return np.array(t["values"]).reshape(t["dims"])
