#!/usr/bin/env python
"""Run the default-planner large-lattice cases one by one, each in a subprocess with a short timeout (debug aid)."""
import subprocess, sys, os
CASES = ["(40,41,42)", "(24,25,26,27)", "(300,300)", "(8,)*6", "(50,3000)", "(50,)*4", "(3,11,11,11,11,11,11)"]
code = """
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from conftest import random_triple
from mrmustard_b200 import strategies as S
import oracle
shape = %s
A, b, c = random_triple(len(shape), (), seed=3)
got = S.vanilla_numba(shape, A, b, complex(c)); got2 = S.vanilla_numba(shape, A, b, complex(c))
want = oracle.vanilla(shape, A, b, complex(c))
print(shape, 'OK' if np.array_equal(got, want) and np.array_equal(got2, want) else 'MISMATCH', flush=True)
"""
for c in CASES if len(sys.argv) < 2 else sys.argv[1:]:
    try:
        r = subprocess.run([sys.executable, "-c", code % c], capture_output=True, text=True, timeout=40)
        print((r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
    except subprocess.TimeoutExpired:
        print(c, "TIMEOUT", flush=True)
