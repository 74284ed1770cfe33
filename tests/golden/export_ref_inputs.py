"""Writes tests/golden/ref_inputs.json: the inputs of every golden case in a form the Julia generator (oracle/make_ref_fixtures.jl)
reads without extra packages.  Run:  python tests/golden/export_ref_inputs.py"""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
cases = []
for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
    z = np.load(f)
    c = {"name": os.path.basename(f)[:-4], "kernel": str(z["kernel"]), "mean": str(z["mean"]), "D": int(z["D"]),
         "X": z["X"].tolist(), "y": z["y"].tolist(), "Xs": z["Xs"].tolist(), "theta": z["theta"].tolist(), "theta2": z["theta2"].tolist()}
    for k in ("EI", "PI", "UCB", "MI"):
        c[k + "_params"] = z[k + "_params"].tolist()
    cases.append(c)
json.dump({"format": "rows of X / Xs are input dimensions (D x N, D x M); theta = [logNoise, (beta), ll..., lsigma]", "cases": cases},
          open(os.path.join(HERE, "ref_inputs.json"), "w"))
print("wrote ref_inputs.json with", len(cases), "cases")
