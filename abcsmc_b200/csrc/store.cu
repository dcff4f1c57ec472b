// store.cu — the storage boundary of the hot path: host code only, no kernels.
//
// Reference: AbcSmc::read_SMC_sets_from_database joins the three tables of its SQLite job database per set and copies the result
// field by field into Eigen matrices through sqdb (src/AbcSmc.cpp:596-621), and writes the ranks back with one UPDATE *string* per
// particle (:653-661). Schema (:819-834): job(serial int pk, smcSet, particleIdx, startTime, duration real, status text, posterior
// int, attempts int), par(serial int pk, seed blob, <short_name> real ...), met(serial int pk, <short_name> real ...). With the numerics
// at milliseconds these two steps are the wall time of `--process`. Here: three scans (job filtered once, par and met each read over the
// set's serial range) merged by serial and stored straight into column-major (pinned) host buffers, row = particleIdx — the layout every
// entry point of this library takes — and one prepared UPDATE re-bound per particle inside a single transaction. Measured against the
// reference's pattern on the same engine by tools/db_bench.py (profiles/r02_db_bench.txt): the engine's record decoding bounds both.
// The database engine is the one AbcSmc already uses: libsqlite3 is loaded with dlopen (no link-time dependency, no header needed:
// the dozen C entry points used are declared below as in sqlite3.h, whose ABI is stable across 3.x).
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/abcsmc_b200.h"

namespace {

struct sqlite3;
struct sqlite3_stmt;
constexpr int SQLITE_OK_ = 0, SQLITE_ROW_ = 100, SQLITE_DONE_ = 101, SQLITE_OPEN_READONLY_ = 1, SQLITE_OPEN_READWRITE_ = 2, SQLITE_NULL_ = 5;

struct SqliteApi {
    void* h = nullptr;
    int (*open_v2)(const char*, sqlite3**, int, const char*) = nullptr;
    int (*close)(sqlite3*) = nullptr;
    int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**) = nullptr;
    int (*step)(sqlite3_stmt*) = nullptr;
    int (*finalize)(sqlite3_stmt*) = nullptr;
    int (*reset)(sqlite3_stmt*) = nullptr;
    int (*bind_int64)(sqlite3_stmt*, int, long long) = nullptr;
    int (*column_count)(sqlite3_stmt*) = nullptr;
    int (*column_type)(sqlite3_stmt*, int) = nullptr;
    double (*column_double)(sqlite3_stmt*, int) = nullptr;
    long long (*column_int64)(sqlite3_stmt*, int) = nullptr;
    int (*exec)(sqlite3*, const char*, int (*)(void*, int, char**, char**), void*, char**) = nullptr;
    const char* (*errmsg)(sqlite3*) = nullptr;
    int (*busy_timeout)(sqlite3*, int) = nullptr;
};

thread_local char g_err[512] = {0};

SqliteApi load_sqlite() {
    SqliteApi a;
    const char* names[] = {"libsqlite3.so.0", "libsqlite3.so"};
    for (const char* n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (a.h) break; }
    if (!a.h) return a;
#define SQ_SYM(field, sym) do { *(void**)(&a.field) = dlsym(a.h, sym); if (!a.field) { a.h = nullptr; return a; } } while (0)
    SQ_SYM(open_v2, "sqlite3_open_v2"); SQ_SYM(close, "sqlite3_close"); SQ_SYM(prepare_v2, "sqlite3_prepare_v2"); SQ_SYM(step, "sqlite3_step");
    SQ_SYM(finalize, "sqlite3_finalize"); SQ_SYM(reset, "sqlite3_reset"); SQ_SYM(bind_int64, "sqlite3_bind_int64"); SQ_SYM(column_count, "sqlite3_column_count");
    SQ_SYM(column_type, "sqlite3_column_type"); SQ_SYM(column_double, "sqlite3_column_double"); SQ_SYM(column_int64, "sqlite3_column_int64");
    SQ_SYM(exec, "sqlite3_exec"); SQ_SYM(errmsg, "sqlite3_errmsg"); SQ_SYM(busy_timeout, "sqlite3_busy_timeout");
#undef SQ_SYM
    return a;
}
SqliteApi& sq() {
    static SqliteApi a = load_sqlite();      // initialised once, thread-safe (C++11 static local)
    return a;
}

struct Db {
    sqlite3* db = nullptr;
    ~Db() { if (db) sq().close(db); }
    int open(const char* path, bool write) {
        if (!sq().h) { snprintf(g_err, sizeof(g_err), "libsqlite3.so.0 could not be loaded"); return ABCB200_ENODEV; }
        if (sq().open_v2(path, &db, write ? SQLITE_OPEN_READWRITE_ : SQLITE_OPEN_READONLY_, nullptr) != SQLITE_OK_) {
            snprintf(g_err, sizeof(g_err), "cannot open %s: %s", path, db ? sq().errmsg(db) : "out of memory");
            return ABCB200_EINVAL;
        }
        sq().busy_timeout(db, 60000);
        return ABCB200_OK;
    }
    int fail(const char* what) { snprintf(g_err, sizeof(g_err), "%s: %s", what, sq().errmsg(db)); return ABCB200_EINVAL; }
};

// number of columns of a table (0 when it does not exist)
int table_columns(Db& d, const char* table) {
    sqlite3_stmt* st = nullptr;
    const std::string q = std::string("select * from ") + table + " limit 0;";
    if (sq().prepare_v2(d.db, q.c_str(), -1, &st, nullptr) != SQLITE_OK_) return 0;
    const int n = sq().column_count(st);
    sq().finalize(st);
    return n;
}

}  // namespace

extern "C" const char* abcb200_db_last_error(void) { return g_err; }

// Shape of one SMC set in an AbcSmc database: particles with smcSet = set, parameter and metric counts from the table layouts.
extern "C" int abcb200_db_set_shape(const char* db_path, int set, int64_t* n_out, int* npar_out, int* nmet_out) {
    if (!db_path) return ABCB200_EINVAL;
    Db d;
    int rc = d.open(db_path, false);
    if (rc != ABCB200_OK) return rc;
    const int pc = table_columns(d, "par"), mc = table_columns(d, "met");
    if (pc < 3 || mc < 2 || table_columns(d, "job") < 8) { snprintf(g_err, sizeof(g_err), "%s: not an AbcSmc database (tables job / par / met, src/AbcSmc.cpp:819-834)", db_path); return ABCB200_EINVAL; }
    if (npar_out) *npar_out = pc - 2;          // serial, seed, then the parameters
    if (nmet_out) *nmet_out = mc - 1;          // serial, then the metrics
    if (n_out) {
        sqlite3_stmt* st = nullptr;
        if (sq().prepare_v2(d.db, "select count(*) from job where smcSet = ?1;", -1, &st, nullptr) != SQLITE_OK_) return d.fail("count");
        sq().bind_int64(st, 1, set);
        *n_out = (sq().step(st) == SQLITE_ROW_) ? (int64_t)sq().column_int64(st, 0) : 0;
        sq().finalize(st);
    }
    return ABCB200_OK;
}

// What the join of src/AbcSmc.cpp:596-621 returns for one set, stored column-major: par (N x P, ld_par), met (N x K, ld_met), row = particleIdx.
// serial_out (N, nullable): job.serial per particle (what the rank write-back needs); posterior_out (N, nullable): job.posterior
// (-1 = not ranked yet). Returns EINVAL when the set's particle indices are not exactly 0 .. N-1, each once (the reference asserts, :613) or a
// metric is still NULL (simulations not finished).
extern "C" int abcb200_db_load_set(const char* db_path, int set, int64_t N, int P, int K, double* par, int64_t ld_par, double* met, int64_t ld_met,
                                   int64_t* serial_out, int32_t* posterior_out) {
    if (!db_path || !par || !met || N < 1 || P < 1 || K < 1 || ld_par < N || ld_met < N) { snprintf(g_err, sizeof(g_err), "db_load_set: bad argument"); return ABCB200_EINVAL; }
    Db d;
    int rc = d.open(db_path, false);
    if (rc != ABCB200_OK) return rc;
    if (table_columns(d, "par") != P + 2 || table_columns(d, "met") != K + 1) {
        snprintf(g_err, sizeof(g_err), "db_load_set: the database holds %d parameters and %d metrics, not %d and %d", table_columns(d, "par") - 2, table_columns(d, "met") - 1, P, K);
        return ABCB200_EINVAL;
    }
    // Three scans merged by serial instead of the reference's three-table join (:597-599): the join costs two index seeks from the
    // root per particle (met and par by serial); here `job` is filtered once (serial -> particleIdx), and `par` / `met` are each read
    // in ONE range scan over the serials of the set, which AbcSmc assigns consecutively per set (:520-551), so the scans touch only
    // this set's rows. Rows are stored at their particleIdx whatever order they come in; nothing is sorted (tools/db_bench.py).
    sqlite3_stmt* st = nullptr;
    if (sq().prepare_v2(d.db, "select serial, particleIdx, posterior from job where smcSet = ?1;", -1, &st, nullptr) != SQLITE_OK_) return d.fail("prepare job");
    sq().bind_int64(st, 1, set);
    std::vector<int64_t> ser((size_t)N, -1);
    std::vector<char> seen((size_t)N, 0);
    int64_t count = 0, smin = INT64_MAX, smax = INT64_MIN;
    int step;
    while ((step = sq().step(st)) == SQLITE_ROW_) {
        const int64_t row = (int64_t)sq().column_int64(st, 1);
        if (row < 0 || row >= N || seen[(size_t)row]) { sq().finalize(st); snprintf(g_err, sizeof(g_err), "db_load_set: particleIdx %lld of set %d is outside 0 .. %lld or appears twice", (long long)row, set, (long long)N - 1); return ABCB200_EINVAL; }
        seen[(size_t)row] = 1;
        const int64_t sv = (int64_t)sq().column_int64(st, 0);
        ser[(size_t)row] = sv; smin = std::min(smin, sv); smax = std::max(smax, sv);
        if (serial_out) serial_out[row] = sv;
        if (posterior_out) posterior_out[row] = (sq().column_type(st, 2) == SQLITE_NULL_) ? -1 : (int32_t)sq().column_int64(st, 2);
        count++;
    }
    sq().finalize(st);
    if (step != SQLITE_DONE_) return d.fail("step job");
    if (count != N) { snprintf(g_err, sizeof(g_err), "db_load_set: set %d has %lld particles, not %lld", set, (long long)count, (long long)N); return ABCB200_EINVAL; }
    // serial -> particleIdx: a dense table over [smin, smax] (consecutive serials), a sorted list when the range is sparse
    const bool dense = (smax - smin) < 8 * N;
    std::vector<int64_t> to_row;
    std::vector<std::pair<int64_t, int64_t>> sorted;
    if (dense) { to_row.assign((size_t)(smax - smin + 1), -1); for (int64_t i = 0; i < N; i++) to_row[(size_t)(ser[(size_t)i] - smin)] = i; }
    else { sorted.resize((size_t)N); for (int64_t i = 0; i < N; i++) sorted[(size_t)i] = {ser[(size_t)i], i}; std::sort(sorted.begin(), sorted.end()); }
    auto row_of = [&](int64_t sv) -> int64_t {
        if (sv < smin || sv > smax) return -1;
        if (dense) return to_row[(size_t)(sv - smin)];
        auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(sv, (int64_t)INT64_MIN));
        return (it != sorted.end() && it->first == sv) ? it->second : -1;
    };
    struct Side { const char* table; int first, ncol; double* dst; int64_t ld; bool null_is_error; };
    const Side sides[2] = {{"par", 2, P, par, ld_par, false}, {"met", 1, K, met, ld_met, true}};     // past (serial, seed) / past serial
    for (const Side& sd : sides) {
        const std::string q = std::string("select * from ") + sd.table + " where serial >= ?1 and serial <= ?2;";
        if (sq().prepare_v2(d.db, q.c_str(), -1, &st, nullptr) != SQLITE_OK_) return d.fail("prepare");
        sq().bind_int64(st, 1, smin); sq().bind_int64(st, 2, smax);
        int64_t got = 0;
        while ((step = sq().step(st)) == SQLITE_ROW_) {
            const int64_t row = row_of((int64_t)sq().column_int64(st, 0));
            if (row < 0) continue;                               // a serial of another set inside the range
            for (int c = 0; c < sd.ncol; c++) {
                const double v = sq().column_double(st, sd.first + c);
                if (v == 0.0 && sd.null_is_error && sq().column_type(st, sd.first + c) == SQLITE_NULL_) {      // NULL reads as 0.0
                    sq().finalize(st);
                    snprintf(g_err, sizeof(g_err), "db_load_set: metric %d of particle %lld of set %d is NULL (simulation not finished)", c, (long long)row, set);
                    return ABCB200_EINVAL;
                }
                sd.dst[(int64_t)c * sd.ld + row] = v;
            }
            got++;
        }
        sq().finalize(st);
        if (step != SQLITE_DONE_) return d.fail("step");
        if (got != N) { snprintf(g_err, sizeof(g_err), "db_load_set: table %s holds %lld of the %lld particles of set %d", sd.table, (long long)got, (long long)N, set); return ABCB200_EINVAL; }
    }
    return ABCB200_OK;
}

// job.posterior = rank for the n particles whose serials are given in rank order (src/AbcSmc.cpp:653-661): one prepared statement,
// one transaction.
extern "C" int abcb200_db_write_ranks(const char* db_path, const int64_t* serial_by_rank, int64_t n) {
    if (!db_path || !serial_by_rank || n < 0) { snprintf(g_err, sizeof(g_err), "db_write_ranks: bad argument"); return ABCB200_EINVAL; }
    Db d;
    int rc = d.open(db_path, true);
    if (rc != ABCB200_OK) return rc;
    if (sq().exec(d.db, "BEGIN EXCLUSIVE;", nullptr, nullptr, nullptr) != SQLITE_OK_) return d.fail("begin");
    sqlite3_stmt* st = nullptr;
    if (sq().prepare_v2(d.db, "update job set posterior = ?1 where serial = ?2;", -1, &st, nullptr) != SQLITE_OK_) { sq().exec(d.db, "ROLLBACK;", nullptr, nullptr, nullptr); return d.fail("prepare"); }
    for (int64_t i = 0; i < n; i++) {
        sq().bind_int64(st, 1, (long long)i);
        sq().bind_int64(st, 2, (long long)serial_by_rank[i]);
        if (sq().step(st) != SQLITE_DONE_) { sq().finalize(st); sq().exec(d.db, "ROLLBACK;", nullptr, nullptr, nullptr); return d.fail("update"); }
        sq().reset(st);
    }
    sq().finalize(st);
    if (sq().exec(d.db, "COMMIT;", nullptr, nullptr, nullptr) != SQLITE_OK_) return d.fail("commit");
    return ABCB200_OK;
}

// `--process` for one set in one call: bulk load (pinned buffers) -> abcb200_chain_process_set -> batched rank write-back.
// A set that already carries ranks (job.posterior > -1, :623-631) is not filtered again: its stored predictive prior is re-used and only
// the statistics / doubled variance / weights are recomputed... by the caller through abcb200_chain_restore; this entry point returns
// EINVAL for such sets. Arguments after `top_n` as abcb200_chain_process_set; numer_all is not available here (flat priors or none).
extern "C" int abcb200_chain_process_db_set(abcb200_chain* ch, const char* db_path, int set, const double* target, int filter, double training_fraction,
                                            int method, int64_t top_n, const int32_t* prior_type, const double* prior_a, const double* prior_b,
                                            uint64_t* order_out, double* weights_out, double* dv_out, double* report_out, int* n_comp_used_out) {
    if (!ch || !db_path || !target) return ABCB200_EINVAL;
    int64_t N = 0; int P = 0, K = 0;
    int rc = abcb200_db_set_shape(db_path, set, &N, &P, &K);
    if (rc != ABCB200_OK) return rc;
    if (N < 1) { snprintf(g_err, sizeof(g_err), "set %d is empty", set); return ABCB200_EINVAL; }
    if (P != abcb200_chain_nparams(ch)) {     // the chain reads abcb200_chain_nparams(ch) parameter columns from the loaded buffer
        snprintf(g_err, sizeof(g_err), "%s holds %d parameters, the chain was created for %d", db_path, P, abcb200_chain_nparams(ch));
        return ABCB200_EINVAL;
    }
    if (top_n <= 0 || top_n > N) top_n = N;
    void *hp = nullptr, *hm = nullptr;
    std::vector<int64_t> serial((size_t)N);
    std::vector<int32_t> post((size_t)N);
    std::vector<uint64_t> order_local;
    if (!order_out) { order_local.resize((size_t)top_n); order_out = order_local.data(); }
    if (abcb200_host_alloc((size_t)N * P * 8, &hp) != ABCB200_OK || abcb200_host_alloc((size_t)N * K * 8, &hm) != ABCB200_OK) {
        if (hp) abcb200_host_free(hp);
        snprintf(g_err, sizeof(g_err), "pinned buffers for %lld x (%d + %d) doubles", (long long)N, P, K);
        return ABCB200_ENOMEM;
    }
    rc = abcb200_db_load_set(db_path, set, N, P, K, (double*)hp, N, (double*)hm, N, serial.data(), post.data());
    if (rc == ABCB200_OK) {
        for (int64_t i = 0; i < N; i++) if (post[(size_t)i] > -1) { snprintf(g_err, sizeof(g_err), "set %d is already ranked (job.posterior > -1)", set); rc = ABCB200_EINVAL; break; }
    }
    if (rc == ABCB200_OK) {
        rc = abcb200_chain_process_set(ch, (const double*)hm, N, (const double*)hp, N, N, K, target, filter, training_fraction, method, top_n, prior_type, prior_a, prior_b,
                                       nullptr, order_out, weights_out, dv_out, report_out, n_comp_used_out);
        if (rc != ABCB200_OK) snprintf(g_err, sizeof(g_err), "abcb200_chain_process_set failed (%d): see abcb200_last_error of the chain's context", rc);
    }
    if (rc == ABCB200_OK) {
        std::vector<int64_t> by_rank((size_t)top_n);
        for (int64_t i = 0; i < top_n; i++) by_rank[(size_t)i] = serial[(size_t)order_out[i]];
        rc = abcb200_db_write_ranks(db_path, by_rank.data(), top_n);
    }
    abcb200_host_free(hp); abcb200_host_free(hm);
    return rc;
}
