/* mpi_shim.c -- single-rank MPI stand-in (TEST INFRASTRUCTURE, see mpi.h) */
#include <mpi.h>
#include <stdlib.h>
#include <time.h>
struct shim_mpi_file { FILE* f; };
int MPI_Init(int* a, char*** b) { (void)a; (void)b; return 0; }
int MPI_Initialized(int* flag) { *flag = 1; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; memcpy(r, s, (size_t)n * t); return 0; }
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; memcpy(r, s, (size_t)n * t); return 0; }
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op; (void)c; memcpy(r, s, (size_t)n * t); return 0; }
int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; return 0; }
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code ? code : 1); }
int MPI_Error_string(int code, char* s, int* len) { *len = sprintf(s, "shim MPI error %d", code); return 0; }
/* hindexed types: the only use is lp-io's gather of its buffers (lp-io.c:304-350); keep the pieces */
static int n_pieces; static int* piece_len; static MPI_Aint* piece_disp;
int MPI_Type_create_hindexed(int n, const int* lens, const MPI_Aint* disp, MPI_Datatype old, MPI_Datatype* newt) {
    (void)old;
    n_pieces = n;
    piece_len = (int*)malloc(sizeof(int) * n);
    piece_disp = (MPI_Aint*)malloc(sizeof(MPI_Aint) * n);
    memcpy(piece_len, lens, sizeof(int) * n);
    memcpy(piece_disp, disp, sizeof(MPI_Aint) * n);
    *newt = -1;
    return 0;
}
int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype* t) { (void)t; free(piece_len); free(piece_disp); piece_len = NULL; piece_disp = NULL; n_pieces = 0; return 0; }
int MPI_File_open(MPI_Comm c, const char* name, int mode, MPI_Info info, MPI_File* fh) {
    (void)c; (void)info;
    FILE* f = fopen(name, (mode & MPI_MODE_WRONLY) ? "wb" : "rb");
    if (!f) { *fh = NULL; return 1; }
    *fh = (MPI_File)malloc(sizeof(struct shim_mpi_file));
    (*fh)->f = f;
    return 0;
}
int MPI_File_close(MPI_File* fh) { if (*fh) { fclose((*fh)->f); free(*fh); *fh = NULL; } return 0; }
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status* st) {
    (void)st;
    fseek(fh->f, (long)off, SEEK_SET);
    if (t == -1) { for (int i = 0; i < n_pieces; i++) fwrite((const char*)buf + piece_disp[i], 1, (size_t)piece_len[i], fh->f); }
    else fwrite(buf, (size_t)t, (size_t)n, fh->f);
    return 0;
}
int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* buf, int n, MPI_Datatype t, MPI_Status* st) { return MPI_File_write_at(fh, off, buf, n, t, st); }
int MPI_File_read_all(MPI_File fh, void* buf, int n, MPI_Datatype t, MPI_Status* st) { (void)st; return fread(buf, (size_t)t, (size_t)n, fh->f) == (size_t)n ? 0 : 1; }
int MPI_File_get_size(MPI_File fh, MPI_Offset* sz) { long cur = ftell(fh->f); fseek(fh->f, 0, SEEK_END); *sz = ftell(fh->f); fseek(fh->f, cur, SEEK_SET); return 0; }
double MPI_Wtime(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
