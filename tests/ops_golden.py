"""Shared reader of tests/golden/ref_ops.json (flux-limiter interpolators and convolutions evaluated by the unmodified reference,
generator oracle/ref_drivers/ref_ops.cpp): mesh, fields' initial functions and, per case, the signature / operand locations."""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PI = 3.141592653589793  # OpFlow::PI (src/Core/Constants.hpp:37)


def load():
    return json.load(open(os.path.join(HERE, "golden", "ref_ops.json")))


def coords(nx, ny):
    """the mesh lambdas of ref_ops.cpp, evaluated in the same order with libm's sin"""
    cx = [2.0 * ((i / (nx - 1)) + 0.15 * math.sin(2 * PI * (i / (nx - 1))) / (2 * PI)) for i in range(nx)]
    cy = [1.0 * ((i / (ny - 1)) + 0.15 * math.sin(2 * PI * (i / (ny - 1))) / (2 * PI)) for i in range(ny)]
    return np.array(cx), np.array(cy)


def fe(x):
    return math.sin(2.3 * x[0]) * math.cos(1.7 * x[1]) + 0.3 * x[0] * x[1]


def fu(x):
    return math.cos(3.1 * x[0] + 0.4) * math.sin(2.2 * x[1] + 0.3)


def kernel_of(node):
    if node == "Conv33":  # UniLS.cpp:107-108
        o, c = 1. / 24., 16. / 24.
        return np.array([[o, o, o], [o, c, o], [o, o, o]])
    return np.array([[0.1 * (i + 1) - 0.07 * (j + 1) * (i - 2) for j in range(3)] for i in range(5)])


def describe(case):
    """-> (signature, [(loc, init function) per field leaf], scalars)"""
    ax = case["axis"]
    if case["node"].startswith("Conv"):
        k = kernel_of(case["node"])
        return f"Conv<{k.shape[0]},{k.shape[1]},1,0,F<0>>", [([1, 1], fe)], list(k.reshape(-1, order="F")), k
    le, lu = [1, 1], [1, 1]
    if case["dir"] == "C2N":
        lu[ax] = 0
    else:
        le[ax] = 0
    return f"Fl{case['node']}{case['dir']}<{ax},F<0>,F<1>>", [(lu, fu), (le, fe)], [], None


def reference_values(case):
    acc = case["acc"]
    shape = [acc[1][0] - acc[0][0], acc[1][1] - acc[0][1]]
    return np.array([float.fromhex(v) for v in case["val"]]).reshape(shape, order="F")
