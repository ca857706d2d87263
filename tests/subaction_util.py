"""Shared by the CPU and GPU tests of the sub-actions: golden cases of tests/golden/subactions (reference outputs made
by tests/golden/make_golden_subactions.py) and the frame generator they were made from."""
import ast
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_subactions import make_frames  # noqa: E402,F401

GOLDEN = os.path.join(HERE, "golden", "subactions")


def cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, h, w, dtype, seed = ast.literal_eval(str(d["recipe"]))
    return d, ast.literal_eval(str(d["kwargs"])), make_frames(n, h, w, dtype, seed)


class Proc:
    """What a sub-action sees of its CombinedActions host (reference stack_framework.py:246-302)."""

    def __init__(self, folder, names, ref_idx=0):
        self.input_full_path, self.filenames, self.ref_idx, self.counts = folder, names, ref_idx, len(names)
        self.msgs = []

    def sub_message_r(self, m):
        self.msgs.append(m)

    def sub_message(self, m, level=None):
        self.msgs.append(m)

    def print_message(self, m, level=None):
        self.msgs.append(m)
