/* mpi.h -- single-rank stand-in for the reference-source oracle build (test infrastructure).
 * Reductions copy, barriers return, file IO goes through stdio. */
#ifndef SHIM_MPI_H
#define SHIM_MPI_H
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef int MPI_Info;
typedef struct { int count; } MPI_Status;
typedef struct shim_mpi_file* MPI_File;
typedef int MPI_Request;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_MAX_ERROR_STRING 256
#define MPI_MODE_RDONLY 1
#define MPI_MODE_CREATE 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_EXCL 8
#define MPI_BOTTOM ((void*)0)
#define MPI_FILE_NULL ((MPI_File)0)
enum { MPI_BYTE = 1, MPI_CHAR = 1, MPI_INT = 4, MPI_LONG = 8, MPI_LONG_LONG = 8, MPI_LONG_LONG_INT = 8, MPI_DOUBLE = 8, MPI_UNSIGNED_LONG = 8, MPI_UNSIGNED_LONG_LONG = 8, MPI_UNSIGNED = 4, MPI_FLOAT = 4, MPI_UINT64_T = 8, MPI_INT64_T = 8 };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN };
int MPI_Init(int* argc, char*** argv);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm c, int* rank);
int MPI_Comm_size(MPI_Comm c, int* size);
int MPI_Barrier(MPI_Comm c);
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Send(const void* b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Error_string(int code, char* s, int* len);
int MPI_Type_create_hindexed(int n, const int* lens, const MPI_Aint* disp, MPI_Datatype old, MPI_Datatype* newt);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);
int MPI_File_open(MPI_Comm c, const char* name, int mode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status* st);
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status* st);
int MPI_File_read_all(MPI_File fh, void* buf, int n, MPI_Datatype t, MPI_Status* st);
int MPI_File_get_size(MPI_File fh, MPI_Offset* sz);
double MPI_Wtime(void);
#ifdef __cplusplus
}
#endif
#endif
