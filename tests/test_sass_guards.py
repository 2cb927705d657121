"""Static guard on the shipped library's SASS: a kernel that can be launched with programmatic stream serialization (the launch
chain of include/ppg.h ppg_set_pdl_chain) must not touch global memory before its `griddepcontrol.wait` (SASS `ACQBULK`).  At
the end of round 2 the compiler had hoisted one load of the action kernels above it (a `const __restrict__` pointer), which
made the chain read a stale row count (tests/test_gpu_rollout.py).  The observation kernel is exempt: it waits at its END by
design and is synchronised with the step kernel through the completion queue."""
import os
import re
import shutil
import subprocess

import pytest

from predpreygrass_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")
def test_no_global_access_before_griddepcontrol_wait():
    sass = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    access = re.compile(r"\b(LDG|LD\.E|LDGSTS|ATOM|ATOMG|RED|STG|ST\.E)\b")
    fn, seen, n, checked, bad = None, False, 0, 0, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, seen, n = m.group(1), False, 0
            continue
        if fn is None or seen:
            continue
        if "ACQBULK" in line:
            seen = True
            if "ppg_obs_kernel" not in fn:
                checked += 1
                if n:
                    bad.append((fn, n))
        elif access.search(line):
            n += 1
    assert checked >= 30, checked  # every step kernel instantiation and both action kernels carry the wait
    assert not bad, bad
    # the observation kernel triggers its dependents at its start and must itself wait for the step kernel before it exits
    # (round-1 advisor finding): both instructions are in every instantiation
    obs = re.split(r"Function : ", sass)
    obs = [f for f in obs if f.startswith("_ZN3ppg14ppg_obs_kernel")]
    assert len(obs) >= 6 and all("PREEXIT" in f and "ACQBULK" in f for f in obs), len(obs)
