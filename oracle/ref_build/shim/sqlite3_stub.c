/* Link-time stand-ins for the eight SQLite entry points the reference's BLAST-DB taxonomy reader
 * references (src/data/blastdb/blastdb.cpp:119-472).  That reader is outside the hot path and is never
 * reached by `makedb`/`blastp`/`blastx` on FASTA or .dmnd input; every stub reports failure. */
#include "sqlite3.h"
int sqlite3_open_v2(const char *f, sqlite3 **db, int fl, const char *v) { (void)f; (void)fl; (void)v; if (db) *db = 0; return 1; }
int sqlite3_close(sqlite3 *db) { (void)db; return 0; }
const char *sqlite3_errmsg(sqlite3 *db) { (void)db; return "sqlite3 not available in the oracle build"; }
int sqlite3_prepare_v2(sqlite3 *db, const char *s, int n, sqlite3_stmt **st, const char **t) { (void)db; (void)s; (void)n; (void)t; if (st) *st = 0; return 1; }
int sqlite3_step(sqlite3_stmt *s) { (void)s; return 1; }
int sqlite3_finalize(sqlite3_stmt *s) { (void)s; return 0; }
int sqlite3_column_int(sqlite3_stmt *s, int i) { (void)s; (void)i; return 0; }
int sqlite3_bind_int(sqlite3_stmt *s, int i, int v) { (void)s; (void)i; (void)v; return 1; }
