"""Shared helpers for the tests: load golden fixtures and build, from the same
constructor arguments, (a) the oracle's ProblemSpec and (b) the zfista_b200 Problem."""
from __future__ import annotations

import ast
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

AB_GRID = [
    (0.0, 0.0), (0.0, 1 / 8), (0.0, 1 / 4),
    (1 / 6, 1 / 144), (1 / 6, 37 / 288), (1 / 6, 1 / 4),
    (1 / 4, 1 / 64), (1 / 4, 17 / 128), (1 / 4, 1 / 4),
    (1 / 2, 1 / 16), (1 / 2, 5 / 32), (1 / 2, 1 / 4),
    (3 / 4, 9 / 64), (3 / 4, 25 / 128), (3 / 4, 1 / 4),
]


def golden_cases():
    """Names of the per-problem full-solve fixtures present (<Problem>__<algo>)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*__*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def case_kwargs(d):
    """Constructor kwargs stored in a fixture (kw_* arrays)."""
    kw = {}
    for k in d.files:
        if not k.startswith("kw_"):
            continue
        v = d[k]
        key = k[3:]
        if key == "bounds":
            kw["bounds"] = (float(v[0]), float(v[1]))
        elif key in ("n_features", "n_objectives"):
            kw[key] = int(v)
        else:
            kw[key] = np.array(v, dtype=np.float64)
    return kw


def case_options(d):
    """Solver options a fixture was generated with."""
    return {str(k): ast.literal_eval(str(v)) for k, v in zip(d["opt_keys"], d["opt_vals"])}


def oracle_spec(cls, kw):
    from oracle import zfista_oracle as zo

    kw = dict(kw)
    bounds = kw.pop("bounds", None)
    return zo.make_spec(cls, bounds=bounds, **kw)


def device_problem(cls, kw):
    import zfista_b200.problems as zp

    return getattr(zp, cls)(**kw)


# name -> (class, kwargs) for the subproblem / problem_eval fixtures (same table as
# tests/golden/make_golden.py:problem_specs, restated so tests do not import the generator)
def fixture_problems():
    def l1(n, m):
        return dict(l1_ratios=(np.arange(m) + 1) / n, l1_shifts=np.arange(m))

    return {
        "JOS1_n5": ("JOS1", dict(n_features=5)),
        "JOS1_n50": ("JOS1", dict(n_features=50)),
        "JOS1_n200": ("JOS1", dict(n_features=200)),
        "JOS1_n5_l1": ("JOS1", dict(n_features=5, **l1(5, 2))),
        "JOS1_n50_l1": ("JOS1", dict(n_features=50, **l1(50, 2))),
        "JOS1_n50_l1nb": ("JOS1", dict(n_features=50, l1_ratios=(1 / 50, 1 / 100),
                                       l1_shifts=(0, 1))),
        "SD": ("SD", dict()),
        "ZDT1_n50": ("ZDT1", dict(n_features=50)),
        "TOI4": ("TOI4", dict()),
        "TOI4_l1": ("TOI4", l1(4, 2)),
        "TRIDIA": ("TRIDIA", dict()),
        "TRIDIA_l1": ("TRIDIA", l1(3, 3)),
        "LFR1_n30": ("LinearFunctionRank1", dict(n_features=30)),
        "FDS_n5": ("FDS", dict(n_features=5)),
        "FDS_n10": ("FDS", dict(n_features=10)),
        "FDS_n10_l1": ("FDS", dict(n_features=10, **l1(10, 3))),
        "FDS_n10_box": ("FDS", dict(n_features=10, bounds=(0, np.inf))),
        "FDS_n100_l1": ("FDS", dict(n_features=100, **l1(100, 3))),
    }
