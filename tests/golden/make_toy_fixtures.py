"""Generates tests/golden/toy_inputs.npz from the reference's own PLS demo data
(/root/reference/lib/PLS/{toyX,toyY,nir,octane}.csv — the inputs lib/PLS/src/main.cpp:21-22 reads).
/root/reference does not exist on the GPU box, so the arrays are committed as a fixture.
Run in the authoring container:  python tests/golden/make_toy_fixtures.py
"""
import os
import numpy as np

REF = "/root/reference/lib/PLS"
out = {}
for name in ("toyX", "toyY", "nir", "octane"):
    out[name] = np.loadtxt(os.path.join(REF, name + ".csv"), delimiter=",", ndmin=2)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "toy_inputs.npz"), **out)
print({k: v.shape for k, v in out.items()})
