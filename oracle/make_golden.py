#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE.

Regenerates the fixtures under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/bin/*, built by
oracle/build_ref.sh from /root/reference).  The fixtures are committed; /root/reference is not needed to run the tests.

    python oracle/make_golden.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(HERE, "_ref", "bin")
ENV = dict(os.environ, LD_LIBRARY_PATH=os.path.join(HERE, "_ref", "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))

EXPLICIT = [  # (fixture name, ref_explicit arguments)
    ("ftcs2d_n65_s100", ["--case", "ftcs2d", "--n", 65, "--steps", 100]),
    ("ftcs2d_n33_sin_s25", ["--case", "ftcs2d", "--n", 33, "--steps", 25, "--init", "sin"]),
    ("ftcs3d_n17_sin_s20", ["--case", "ftcs3d", "--n", 17, "--steps", 20, "--init", "sin"]),
    ("ftcs3d_n21_zero_s15", ["--case", "ftcs3d", "--n", 21, "--steps", 15]),
    ("weno_down_n257_s40", ["--case", "weno_down", "--n", 257, "--steps", 40, "--ghosts", 1]),
    ("weno_up_n257_s40", ["--case", "weno_up", "--n", 257, "--steps", 40, "--ghosts", 1]),
    ("upwind1_n101_s100", ["--case", "upwind1", "--n", 101, "--steps", 100, "--ghosts", 1]),
    ("weno_down_n129_sin_s30", ["--case", "weno_down", "--n", 129, "--steps", 30, "--ghosts", 1, "--init", "sin"]),
    ("ftcs2d_mpi_n65_sin_s200", ["--case", "ftcs2d_mpi", "--n", 65, "--steps", 200, "--init", "sin", "--ghosts", 1]),
    ("ftcs2d_fbc_n41_sin_s60", ["--case", "ftcs2d_fbc", "--n", 41, "--steps", 60, "--init", "sin", "--ghosts", 1]),
]


def main():
    os.makedirs(GOLD, exist_ok=True)
    manifest = {}
    for name, args in EXPLICIT:
        path = os.path.join(GOLD, name + ".opfd")
        out = subprocess.run([os.path.join(BIN, "ref_explicit"), *map(str, args), "--dump", path], env=ENV, capture_output=True, text=True, check=True)
        manifest[name] = {"args": [str(a) for a in args], "stdout": json.loads(out.stdout.strip().splitlines()[-1])}
        manifest[name]["stdout"].pop("seconds", None), manifest[name]["stdout"].pop("mlups", None)
    out = subprocess.run([os.path.join(BIN, "ref_fields")], env=ENV, capture_output=True, text=True, check=True)
    json.loads(out.stdout)  # validate
    open(os.path.join(GOLD, "ref_fields.json"), "w").write(out.stdout)
    for drv in ("ref_implicit", "ref_csr"):
        exe = os.path.join(BIN, drv)
        if os.path.exists(exe):
            out = subprocess.run([exe], env=ENV, capture_output=True, text=True, check=True)
            json.loads(out.stdout)
            open(os.path.join(GOLD, drv + ".json"), "w").write(out.stdout)
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    sys.exit(main())
