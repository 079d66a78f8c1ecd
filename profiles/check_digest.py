"""Compare the batched kernel's results with the stored digest (tests/golden/device_digest.json):
prints the workloads whose nit / x / fun bytes differ.  python profiles/check_digest.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from device_digest import ROOT, digest  # noqa: E402

ref = json.load(open(os.path.join(ROOT, "tests", "golden", "device_digest.json")))
got = digest()
bad = 0
for k in sorted(ref):
    same = all(got[k][f] == ref[k][f] for f in ("nit_sha", "x_sha", "fun_sha"))
    bad += not same
    print(f"{k:24s} {'identical' if same else 'DIFFERENT'}  nit_sum {got[k]['nit_sum']} (ref "
          f"{ref[k]['nit_sum']})  n_dual {got[k]['n_dual_sum']} (ref {ref[k]['n_dual_sum']})")
if len(sys.argv) > 1:
    json.dump(got, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print("ALL IDENTICAL" if bad == 0 else f"{bad} workloads differ")
