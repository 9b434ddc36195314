#!/bin/bash
mkdir -p /tmp/lid && cd /tmp/lid
OPF_SOLVER_DEBUG=1 timeout 300 /root/repo/tests/frontend/_bin/ref_LidDriven2D 2>&1 | grep -v "Current step" | head -12
OPF_SOLVER_DEBUG=1 timeout 300 /root/repo/tests/frontend/_bin/ref_LidDriven2D 2>&1 | grep -v "Current step" | tail -4
python - <<'PY'
import re, numpy as np
for name in "uvp":
    txt=open(f"/tmp/lid/{name}.tec").read()
    lines=txt.rsplit("ZONE\n",1)[1].splitlines()
    ext=[int(x) for x in re.findall(r"= (\d+)", lines[1])]; n=int(np.prod(ext))
    body=lines[3:]
    d=np.array([float(x) for x in body[2*n:3*n]])
    print(name, ext, d.min(), d.max(), np.isnan(d).sum())
PY
