// db_bench.cpp — the storage boundary timed on the host: abcb200_db_load_set / abcb200_db_write_ranks against the
// reference's access pattern on the same AbcSmc database, with the same SQLite engine (dlopen, as store.cu does).
// Reference pattern (restated, src/AbcSmc.cpp:596-621, 653-661, 726-749): the three-table join WITHOUT an ORDER BY, every field fetched
// with its own column call and stored at (row, col) of a column-major matrix (a strided write per field), then one UPDATE *string* per
// ranked particle — built with a stringstream, prepared, stepped and finalised one by one — inside a single exclusive transaction.
// No GPU involved. usage: db_bench <abc.sqlite> <set> ; prints one JSON line.
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

#include "../include/abcsmc_b200.h"

struct sqlite3; struct sqlite3_stmt;
static int (*s_open)(const char*, sqlite3**, int, const char*);
static int (*s_close)(sqlite3*);
static int (*s_prepare)(sqlite3*, const char*, int, sqlite3_stmt**, const char**);
static int (*s_step)(sqlite3_stmt*);
static int (*s_finalize)(sqlite3_stmt*);
static double (*s_col_double)(sqlite3_stmt*, int);
static int (*s_col_int)(sqlite3_stmt*, int);
static int (*s_exec)(sqlite3*, const char*, int (*)(void*, int, char**, char**), void*, char**);

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: db_bench <abc.sqlite> <set>\n"); return 2; }
    const char* path = argv[1];
    const int set = atoi(argv[2]);
    void* h = dlopen("libsqlite3.so.0", RTLD_NOW);
    if (!h) { fprintf(stderr, "libsqlite3.so.0 not found\n"); return 3; }
    *(void**)&s_open = dlsym(h, "sqlite3_open_v2"); *(void**)&s_close = dlsym(h, "sqlite3_close"); *(void**)&s_prepare = dlsym(h, "sqlite3_prepare_v2");
    *(void**)&s_step = dlsym(h, "sqlite3_step"); *(void**)&s_finalize = dlsym(h, "sqlite3_finalize"); *(void**)&s_col_double = dlsym(h, "sqlite3_column_double");
    *(void**)&s_col_int = dlsym(h, "sqlite3_column_int"); *(void**)&s_exec = dlsym(h, "sqlite3_exec");

    int64_t N = 0; int P = 0, K = 0;
    if (abcb200_db_set_shape(path, set, &N, &P, &K) != 0) { fprintf(stderr, "%s\n", abcb200_db_last_error()); return 4; }
    std::vector<double> par((size_t)N * P), met((size_t)N * K), par_r((size_t)N * P), met_r((size_t)N * K);
    std::vector<int64_t> serial((size_t)N);
    std::vector<int32_t> post((size_t)N);

    // ---- load: the library
    double t0 = now();
    if (abcb200_db_load_set(path, set, N, P, K, par.data(), N, met.data(), N, serial.data(), post.data()) != 0) { fprintf(stderr, "%s\n", abcb200_db_last_error()); return 5; }
    const double t_load = now() - t0;

    // ---- load: the reference's pattern
    sqlite3* db = nullptr;
    if (s_open(path, &db, 2 /* READWRITE */, nullptr) != 0) return 6;
    std::vector<int> serial_r((size_t)N);
    t0 = now();
    {
        std::stringstream q;
        q << "select J.serial, J.particleIdx, J.posterior, ";
        for (int p = 0; p < P; p++) q << "P.p" << p << ", ";
        for (int k = 0; k < K; k++) q << "M.m" << k << (k + 1 < K ? ", " : " ");
        q << "from job J, met M, par P where J.serial = M.serial and J.serial = P.serial and J.smcSet = " << set << ";";
        sqlite3_stmt* st = nullptr;
        if (s_prepare(db, q.str().c_str(), -1, &st, nullptr) != 0) return 7;
        int64_t row = 0;
        while (s_step(st) == 100) {
            const int ser = s_col_int(st, 0); const int idx = s_col_int(st, 1); const int rank = s_col_int(st, 2);
            if (idx != row) { fprintf(stderr, "particleIdx out of order\n"); return 8; }
            (void)rank;
            serial_r[(size_t)row] = ser;
            for (int p = 0; p < P; p++) par_r[(size_t)p * N + row] = s_col_double(st, 3 + p);
            for (int k = 0; k < K; k++) met_r[(size_t)k * N + row] = s_col_double(st, 3 + P + k);
            row++;
        }
        s_finalize(st);
        if (row != N) return 9;
    }
    const double t_load_ref = now() - t0;
    if (par != par_r || met != met_r) { fprintf(stderr, "the two loads disagree\n"); return 10; }

    // ---- rank write-back of the first N / 20 particles (predictive prior 5 %), twice: once per method, ranks reset in between
    const int64_t n_pp = N / 20 > 0 ? N / 20 : 1;
    std::vector<int64_t> by_rank((size_t)n_pp);
    for (int64_t i = 0; i < n_pp; i++) by_rank[(size_t)i] = serial[(size_t)((i * 7919) % N)];
    t0 = now();
    if (abcb200_db_write_ranks(path, by_rank.data(), n_pp) != 0) { fprintf(stderr, "%s\n", abcb200_db_last_error()); return 11; }
    const double t_write = now() - t0;
    s_exec(db, "update job set posterior = -1;", nullptr, nullptr, nullptr);
    t0 = now();
    {
        std::vector<std::string> updates((size_t)n_pp);
        for (int64_t i = 0; i < n_pp; i++) { std::stringstream ss; ss << "update job set posterior = " << i << " where serial = " << by_rank[(size_t)i] << ";"; updates[(size_t)i] = ss.str(); }
        s_exec(db, "BEGIN EXCLUSIVE;", nullptr, nullptr, nullptr);
        for (const auto& u : updates) { sqlite3_stmt* st = nullptr; if (s_prepare(db, u.c_str(), -1, &st, nullptr) != 0) return 12; s_step(st); s_finalize(st); }
        s_exec(db, "COMMIT;", nullptr, nullptr, nullptr);
    }
    const double t_write_ref = now() - t0;
    s_exec(db, "update job set posterior = -1;", nullptr, nullptr, nullptr);
    s_close(db);
    printf("{\"N\": %lld, \"P\": %d, \"K\": %d, \"n_ranked\": %lld, \"load_ms\": %.2f, \"load_reference_pattern_ms\": %.2f, \"write_ranks_ms\": %.2f, "
           "\"write_ranks_reference_pattern_ms\": %.2f}\n", (long long)N, P, K, (long long)n_pp, t_load * 1e3, t_load_ref * 1e3, t_write * 1e3, t_write_ref * 1e3);
    return 0;
}
