import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from flowgnn_b200.capi import Context
from flowgnn_b200.dataset import load_npz
from flowgnn_b200.weights import load_weights
b = load_npz("/root/repo/tests/golden/molhiv.npz").slice(0, 64)
w = load_weights("gin", "/root/repo/tests/golden/weights/GIN")
g = np.load("/root/repo/tests/golden/golden_molhiv.npz")["gin"][:64]
with Context(0) as c:
    y = c.run("gin", b, w)
    print("max err", np.abs(y-g).max())
