"""Margins of the rounding-sensitive two-objective cases (L1 terms, TOI4): device against the
reference's stored result, next to the reference's own 1-ulp envelope (8 seeds)."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers  # noqa: E402

FIX = helpers.fixture_problems()
for case in helpers.golden_cases():
    pname = case.split("__")[0]
    if FIX.get(pname, ("",))[0] not in ("JOS1", "SD", "ZDT1", "TOI4"):
        continue
    if not ("_l1" in pname or pname.startswith("TOI4")):
        continue
    d = helpers.load(case)
    cls, kw = str(d["problem"]), helpers.case_kwargs(d)
    prob = helpers.device_problem(cls, kw)
    opts = helpers.case_options(d)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        br = prob.minimize_proximal_gradient_batched(d["x0"], **opts)
    env = helpers.oracle_noise_envelope(helpers.oracle_spec(cls, kw), d["x0"], d["x"], d["fun"],
                                        d["nit"], opts, n_starts=16 if prob.n_features <= 10 else 4,
                                        seeds=tuple(range(8)))
    dnit = int(np.abs(br.nit - d["nit"]).max())
    dx = float(np.max(np.abs(br.x - d["x"])))
    dF = float(np.max(np.abs(br.fun - d["fun"]) / np.maximum(1.0, np.abs(d["fun"]))))
    print(f"{case:28s} nit_max {int(d['nit'].max()):5d} | dnit {dnit:4d} env {env['dnit']:4d} | "
          f"dx {dx:.2e} env {env['dx']:.2e} | dF {dF:.2e} env {env['dF']:.2e}", flush=True)
