#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time ZFVM_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "device_stencil_search" ) > gpurun_out/r02_pytest_stsearch.log 2>&1; grep "zfvm stencils\|passed\|failed\|Error" gpurun_out/r02_pytest_stsearch.log | tail -40
ZFVM_VERBOSE=1 timeout 900 python - <<'PY' 2>&1 | grep -v "zfvm grid" | tee gpurun_out/r02_stsearch_n118.log
import time, os, hashlib, numpy as np
from zisafvm_b200 import cases
t=time.perf_counter(); case = cases.blast_3d(n=118, order=3, kind="blast"); print("case", round(time.perf_counter()-t,2), flush=True)
t=time.perf_counter(); st = case.ensure_stencils(); print("stencils (device)", round(time.perf_counter()-t,2), flush=True)
h=hashlib.md5()
for a in st.export_arrays(): h.update(np.ascontiguousarray(a).tobytes())
d=h.hexdigest()
case.stencils=None; os.environ["ZFVM_STENCILS"]="host"
t=time.perf_counter(); st = case.ensure_stencils(); print("stencils (host)", round(time.perf_counter()-t,2), flush=True)
h=hashlib.md5()
for a in st.export_arrays(): h.update(np.ascontiguousarray(a).tobytes())
print("identical" if h.hexdigest()==d else "DIFFERENT")
PY
