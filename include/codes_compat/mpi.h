/* mpi.h -- single-rank MPI declarations for builds of a CODES driver against the B200 engine without an MPI
 * installation (one process drives the GPU; multi-GPU runs use the engine's own group sharding, dfdally_b200.h).
 * Reductions copy their input, rank is 0, size is 1. */
#ifndef CODES_COMPAT_MPI_H
#define CODES_COMPAT_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef int MPI_Info;
typedef int MPI_Request;
typedef struct MPI_Status {
    int count;
} MPI_Status;
typedef struct codes_compat_mpi_file* MPI_File;
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
/* a datatype handle is the element size in bytes */
enum { MPI_BYTE = 1, MPI_CHAR = 1, MPI_INT = 4, MPI_UNSIGNED = 4, MPI_FLOAT = 4, MPI_LONG = 8, MPI_LONG_LONG = 8, MPI_LONG_LONG_INT = 8,
       MPI_DOUBLE = 8, MPI_UNSIGNED_LONG = 8, MPI_UNSIGNED_LONG_LONG = 8, MPI_UINT64_T = 8, MPI_INT64_T = 8 };
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
int MPI_Abort(MPI_Comm c, int code);
double MPI_Wtime(void);
#ifdef __cplusplus
}
#endif
#endif
