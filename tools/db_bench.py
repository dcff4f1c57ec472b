"""Builds an AbcSmc job database of one SMC set at a BASELINE.json shape (the reference's schema, src/AbcSmc.cpp:819-834; values at the 6
significant digits its streams print) and runs tools/db_bench.cpp on it: the library's bulk load / batched rank write-back against the
reference's access pattern, same SQLite engine, host only (no GPU). usage: python tools/db_bench.py [N] [P] [K]"""
import os
import sqlite3
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from abcsmc_b200 import _capi, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
_capi.build()
exe = os.path.join(ROOT, "tools", "db_bench")
libdir = os.path.join(ROOT, "abcsmc_b200")
subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tools", "db_bench.cpp"), f"-L{libdir}", "-labcsmc_b200", f"-Wl,-rpath,{libdir}", "-ldl", "-o", exe])
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "abc.sqlite")
    par, met, _ = synth.make_set(N, P, K, 7)
    con = sqlite3.connect(path)
    cur = con.cursor()
    cur.execute("create table job ( serial int primary key asc, smcSet int, particleIdx int, startTime int, duration real, status text, posterior int, attempts int );")
    cur.execute("create index idx1 on job (status, attempts);")
    cur.execute("create table par ( serial int primary key, seed blob, " + ", ".join(f"p{j} real" for j in range(P)) + ");")
    cur.execute("create table met ( serial int primary key, " + ", ".join(f"m{j} real" for j in range(K)) + ");")
    t0 = time.time()
    cur.executemany("insert into job values (?, 0, ?, 0, NULL, 'D', -1, 0);", ((i, i) for i in range(N)))
    cur.executemany("insert into par values (" + ", ".join("?" * (P + 2)) + ");", ([i, str(1000 + i)] + [float(f"{v:.6g}") for v in par[i]] for i in range(N)))
    cur.executemany("insert into met values (" + ", ".join("?" * (K + 1)) + ");", ([i] + [float(f"{v:.6g}") for v in met[i]] for i in range(N)))
    con.commit(); con.close()
    print(f"# database of {N} particles x ({P} parameters, {K} metrics) built in {time.time() - t0:.1f} s, {os.path.getsize(path) / 1e6:.1f} MB", file=sys.stderr)
    for rep in range(3):
        print(subprocess.run([exe, path, "0"], capture_output=True, text=True, check=True).stdout.strip())
