"""The storage boundary: bulk load of one SMC set from an AbcSmc SQLite database and the batched rank write-back.
The database is built here with the reference's schema and value formatting (src/AbcSmc.cpp:819-834; parameters and metrics are
streamed into the SQL text at the default 6 significant digits, :537, :1018). Host code only: runs without a GPU; the one-call
`--process` of a set (load -> chain -> write-back) is the gpu-marked test at the end."""
import ctypes as C
import sqlite3

import numpy as np
import pytest

from abcsmc_b200 import _capi, synth


def _make_db(path, sets, P, K, seed=7):
    """sets: list of N. Returns per set (par, met) as stored (values rounded to 6 significant digits like the reference's streams)."""
    con = sqlite3.connect(path)
    cur = con.cursor()
    cur.execute("create table job ( serial int primary key asc, smcSet int, particleIdx int, startTime int, duration real, status text, posterior int, attempts int );")
    cur.execute("create index idx1 on job (status, attempts);")
    cur.execute("create table par ( serial int primary key, seed blob, " + ", ".join(f"p{j} real" for j in range(P)) + ");")
    cur.execute("create table met ( serial int primary key, " + ", ".join(f"m{j} real" for j in range(K)) + ");")
    out, serial = [], 0
    for t, N in enumerate(sets):
        par, met, target = synth.make_set(N, P, K, seed + t)
        par = np.asfortranarray(np.array([[float(f"{v:.6g}") for v in row] for row in par]))
        met = np.asfortranarray(np.array([[float(f"{v:.6g}") for v in row] for row in met]))
        for i in range(N):
            cur.execute(f"insert into job values ( {serial}, {t}, {i}, 0, NULL, 'D', -1, 0 );")
            cur.execute(f"insert into par values ( {serial}, '{1000 + serial}', " + ", ".join(f"{v:.6g}" for v in par[i]) + " );")
            cur.execute(f"insert into met values ( {serial}, " + ", ".join(f"{v:.6g}" for v in met[i]) + " );")
            serial += 1
        out.append((par, met, target))
    con.commit(); con.close()
    return out


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def test_bulk_load_and_rank_write_back(tmp_path):
    _capi.build()
    lib = _capi.lib()
    db = str(tmp_path / "abc.sqlite").encode()
    P, K, sets = 3, 5, [40, 64]
    stored = _make_db(db.decode(), sets, P, K)
    for t, N in enumerate(sets):
        n, p, k = C.c_int64(0), C.c_int(0), C.c_int(0)
        assert lib.abcb200_db_set_shape(db, t, C.byref(n), C.byref(p), C.byref(k)) == 0, lib.abcb200_db_last_error()
        assert (n.value, p.value, k.value) == (N, P, K)
        ld = N + 8                                          # a leading dimension larger than N
        par = np.zeros((ld, P), order="F"); met = np.zeros((ld, K), order="F")
        serial = np.zeros(N, dtype=np.int64); post = np.zeros(N, dtype=np.int32)
        assert lib.abcb200_db_load_set(db, t, N, P, K, _ptr(par), ld, _ptr(met), ld, _ptr(serial), _ptr(post)) == 0, lib.abcb200_db_last_error()
        assert np.array_equal(par[:N], stored[t][0]) and np.array_equal(met[:N], stored[t][1])
        assert np.array_equal(serial, np.arange(N) + sum(sets[:t])) and np.all(post == -1)
    # ranks of set 1: any permutation of half of its particles
    N = sets[1]
    order = np.random.default_rng(3).permutation(N)[: N // 2]
    by_rank = (order + sets[0]).astype(np.int64)
    assert lib.abcb200_db_write_ranks(db, _ptr(by_rank), len(by_rank)) == 0, lib.abcb200_db_last_error()
    con = sqlite3.connect(db.decode())
    rows = dict(con.execute("select particleIdx, posterior from job where smcSet = 1;").fetchall())
    con.close()
    for i in range(N):
        want = int(np.where(order == i)[0][0]) if i in order else -1
        assert rows[i] == want
    post = np.zeros(N, dtype=np.int32)
    par = np.zeros((N, P), order="F"); met = np.zeros((N, K), order="F")
    assert lib.abcb200_db_load_set(db, 1, N, P, K, _ptr(par), N, _ptr(met), N, None, _ptr(post)) == 0
    assert np.array_equal(np.argsort(np.where(post >= 0, post, 10 ** 6))[: N // 2], order)


def test_load_errors(tmp_path):
    _capi.build()
    lib = _capi.lib()
    db = str(tmp_path / "abc.sqlite")
    _make_db(db, [10], 2, 2)
    par = np.zeros((10, 2), order="F"); met = np.zeros((10, 2), order="F")
    assert lib.abcb200_db_load_set(db.encode(), 0, 10, 3, 2, _ptr(par), 10, _ptr(met), 10, None, None) == -1      # wrong parameter count
    assert b"2 parameters" in lib.abcb200_db_last_error()
    par12 = np.zeros((12, 2), order="F"); met12 = np.zeros((12, 2), order="F")
    assert lib.abcb200_db_load_set(db.encode(), 0, 12, 2, 2, _ptr(par12), 12, _ptr(met12), 12, None, None) == -1  # wrong set size
    con = sqlite3.connect(db); con.execute("update met set m1 = NULL where serial = 4;"); con.commit(); con.close()
    assert lib.abcb200_db_load_set(db.encode(), 0, 10, 2, 2, _ptr(par), 10, _ptr(met), 10, None, None) == -1      # unfinished simulation
    assert b"NULL" in lib.abcb200_db_last_error()
    assert lib.abcb200_db_set_shape(str(tmp_path / "missing.sqlite").encode(), 0, None, None, None) == -1


@pytest.mark.parametrize("stride,shuffle", [(1, False), (2, True), (50, True)])
def test_load_is_independent_of_serial_layout_and_row_order(tmp_path, stride, shuffle):
    """The load merges three scans by serial: interleaved sets (serials of another set inside the range), widely spaced serials (the
    sorted lookup instead of the dense table), rows inserted in another order than particleIdx, a metric that is exactly 0.0 (not
    NULL), a missing par row."""
    _capi.build()
    lib = _capi.lib()
    db = str(tmp_path / "abc.sqlite")
    P, K, N = 3, 4, 60
    rng = np.random.default_rng(stride)
    con = sqlite3.connect(db); cur = con.cursor()
    cur.execute("create table job ( serial int primary key asc, smcSet int, particleIdx int, startTime int, duration real, status text, posterior int, attempts int );")
    cur.execute("create table par ( serial int primary key, seed blob, " + ", ".join(f"p{j} real" for j in range(P)) + ");")
    cur.execute("create table met ( serial int primary key, " + ", ".join(f"m{j} real" for j in range(K)) + ");")
    data = {t: (rng.random((N, P)), rng.random((N, K))) for t in (0, 1)}
    data[1][1][7, 2] = 0.0
    rows = [(t, i) for t in (0, 1) for i in range(N)]
    if shuffle:
        rows = [rows[j] for j in rng.permutation(len(rows))]
    serial_of = {}
    for t, i in rows:
        sv = 1000 + stride * (2 * i + t)                                   # the two sets interleave
        serial_of[(t, i)] = sv
        cur.execute("insert into job values (?, ?, ?, 0, NULL, 'D', -1, 0);", (sv, t, i))
        cur.execute("insert into par values (?, ?, " + ", ".join("?" * P) + ");", [sv, str(sv)] + [float(v) for v in data[t][0][i]])
        cur.execute("insert into met values (?, " + ", ".join("?" * K) + ");", [sv] + [float(v) for v in data[t][1][i]])
    con.commit(); con.close()
    for t in (0, 1):
        par = np.zeros((N, P), order="F"); met = np.zeros((N, K), order="F"); serial = np.zeros(N, dtype=np.int64)
        assert lib.abcb200_db_load_set(db.encode(), t, N, P, K, _ptr(par), N, _ptr(met), N, _ptr(serial), None) == 0, lib.abcb200_db_last_error()
        assert np.array_equal(par, data[t][0]) and np.array_equal(met, data[t][1])
        assert np.array_equal(serial, np.array([serial_of[(t, i)] for i in range(N)]))
    con = sqlite3.connect(db); con.execute("delete from par where serial = ?;", (serial_of[(1, 5)],)); con.commit(); con.close()
    par = np.zeros((N, P), order="F"); met = np.zeros((N, K), order="F")
    assert lib.abcb200_db_load_set(db.encode(), 1, N, P, K, _ptr(par), N, _ptr(met), N, None, None) == -1
    assert b"table par holds 59 of the 60" in lib.abcb200_db_last_error()
    assert lib.abcb200_db_load_set(db.encode(), 0, N, P, K, _ptr(par), N, _ptr(met), N, None, None) == 0


@pytest.mark.gpu
def test_process_db_sets_in_one_call_each(tmp_path, oracle):
    """Two sets of a database processed the way `abc --process` would: load, filter, weights, ranks written back — one call per set."""
    from abcsmc_b200 import api
    lib = _capi.lib()
    db = str(tmp_path / "abc.sqlite")
    P, K, sets = 4, 6, [600, 800]
    stored = _make_db(db, sets, P, K, seed=21)
    ctx = api.get_context(0)
    chain = api.SmcChain(P, ctx)
    pt = np.zeros(P, dtype=np.int32); pa = np.zeros(P); pb = np.ones(P)
    prev = None
    try:
        for t, N in enumerate(sets):
            n_pp = N // 4
            par, met, target = stored[t]
            order = np.zeros(n_pp, dtype=np.uint64); w = np.zeros(n_pp); dv = np.zeros(P); rep = np.zeros(1 + 2 * (P + K)); used = C.c_int(0)
            rc = lib.abcb200_chain_process_db_set(chain._h, db.encode(), t, _ptr(np.ascontiguousarray(target)), 0, 0.5, 0, n_pp, _ptr(pt), _ptr(pa), _ptr(pb),
                                                  _ptr(order), _ptr(w), _ptr(dv), _ptr(rep), C.cast(C.byref(used), C.c_void_p))
            assert rc == 0, (lib.abcb200_db_last_error(), ctx._lib.abcb200_last_error(ctx._h))
            ref = oracle.particle_ranking_PLS(met, par, target, 0.5)
            assert np.array_equal(order.astype(np.int64), ref["order"][:n_pp].astype(np.int64))
            sel = par[order.astype(np.int64)]
            np.testing.assert_allclose(dv, oracle.calculate_doubled_variance(sel), rtol=1e-10)
            if prev is None:
                assert np.all(w == 1.0 / n_pp)
            else:
                numer = np.array([np.prod([oracle.prior_likelihood(0, 0.0, 1.0, float(v)) for v in row]) for row in sel])
                np.testing.assert_allclose(w, oracle.weight_predictive_prior(numer, sel, prev[0], prev[1], prev[2]), rtol=1e-10)
            prev = (sel, w.copy(), dv.copy())
            con = sqlite3.connect(db)
            got = con.execute(f"select particleIdx from job where smcSet = {t} and posterior > -1 order by posterior;").fetchall()
            con.close()
            assert [g[0] for g in got] == [int(v) for v in order]
        # a ranked set is not filtered twice
        rc = lib.abcb200_chain_process_db_set(chain._h, db.encode(), 1, _ptr(np.ascontiguousarray(stored[1][2])), 0, 0.5, 0, 10, None, None, None, None, None, None, None, None)
        assert rc == -1 and b"already ranked" in lib.abcb200_db_last_error()
        # a chain created for another number of parameters is refused before anything is read
        assert lib.abcb200_chain_nparams(chain._h) == P
        other = api.SmcChain(P + 3, ctx)
        try:
            rc = lib.abcb200_chain_process_db_set(other._h, db.encode(), 0, _ptr(np.ascontiguousarray(stored[0][2])), 0, 0.5, 0, 10, None, None, None, None, None, None, None, None)
            assert rc == -1 and b"the chain was created for" in lib.abcb200_db_last_error()
        finally:
            other.close()
    finally:
        chain.close()
