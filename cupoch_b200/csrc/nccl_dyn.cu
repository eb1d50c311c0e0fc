// nccl_dyn.cu -- NCCL bound at run time (dlopen) so that the library loads on
// hosts without NCCL and shares the process's libnccl.so.2 (torch's) when one
// is already loaded.  Only the 32-double all-reduce of the ICP sums uses it.
#include <dlfcn.h>
#include <string.h>

#include "cphb_internal.cuh"

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm_t);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char *(*fn_errstr)(int);

static struct {
    void *h;
    fn_get_uid get_uid;
    fn_init_rank init_rank;
    fn_destroy destroy;
    fn_allreduce allreduce;
    fn_errstr errstr;
} g_nccl;

static int nccl_load() {
    if (g_nccl.h) return CPHB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        cphb_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
        return CPHB_ERR_NCCL;
    }
    g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allreduce) {
        cphb_set_error("NCCL symbols missing");
        return CPHB_ERR_NCCL;
    }
    g_nccl.h = h;
    return CPHB_OK;
}
#define NCCL_TRY(x)                                                                          \
    do {                                                                                     \
        int r__ = (x);                                                                       \
        if (r__ != 0) {                                                                      \
            cphb_set_error("%s: %s", #x, g_nccl.errstr ? g_nccl.errstr(r__) : "nccl error"); \
            return CPHB_ERR_NCCL;                                                            \
        }                                                                                    \
    } while (0)

extern "C" int cphb_nccl_unique_id(char h_id[128]) {
    int rc = nccl_load();
    if (rc) return rc;
    nccl_uid u;
    NCCL_TRY(g_nccl.get_uid(&u));
    memcpy(h_id, u.internal, 128);
    return CPHB_OK;
}
extern "C" int cphb_nccl_comm_init(const char h_id[128], int world_size, int rank, void **out_comm) {
    int rc = nccl_load();
    if (rc) return rc;
    nccl_uid u;
    memcpy(u.internal, h_id, 128);
    nccl_comm_t c = nullptr;
    NCCL_TRY(g_nccl.init_rank(&c, world_size, u, rank));
    *out_comm = c;
    return CPHB_OK;
}
extern "C" int cphb_nccl_comm_destroy(void *comm) {
    if (!comm) return CPHB_OK;
    int rc = nccl_load();
    if (rc) return rc;
    NCCL_TRY(g_nccl.destroy((nccl_comm_t)comm));
    return CPHB_OK;
}
int cphb_nccl_allreduce_f64(void *comm, const double *send, double *recv, size_t count, cudaStream_t s) {
    int rc = nccl_load();
    if (rc) return rc;
    // ncclFloat64 = 8, ncclSum = 0 (nccl.h)
    NCCL_TRY(g_nccl.allreduce(send, recv, count, 8, 0, (nccl_comm_t)comm, s));
    return CPHB_OK;
}
