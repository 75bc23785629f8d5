/* Declarations-only stand-in for <sqlite3.h> (header package absent from this image; libsqlite3.so.0 is
 * present).  Only used to compile the reference's BLAST-DB reader (src/data/blastdb/blastdb.cpp), which is
 * outside the hot path and never executed by the oracle runs.  Written from the public SQLite C API docs. */
#ifndef DMND_ORACLE_SQLITE3_SHIM_H
#define DMND_ORACLE_SQLITE3_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
#define SQLITE_OK 0
#define SQLITE_ROW 100
#define SQLITE_DONE 101
#define SQLITE_OPEN_READONLY 0x00000001
int sqlite3_open_v2(const char *filename, sqlite3 **ppDb, int flags, const char *zVfs);
int sqlite3_close(sqlite3 *);
const char *sqlite3_errmsg(sqlite3 *);
int sqlite3_prepare_v2(sqlite3 *db, const char *zSql, int nByte, sqlite3_stmt **ppStmt, const char **pzTail);
int sqlite3_step(sqlite3_stmt *);
int sqlite3_finalize(sqlite3_stmt *);
int sqlite3_column_int(sqlite3_stmt *, int iCol);
int sqlite3_bind_int(sqlite3_stmt *, int, int);
#ifdef __cplusplus
}
#endif
#endif
