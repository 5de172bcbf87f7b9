// Genome-partition mode (BASELINE.json configs[4]; fm_index.erl:113-124 writes the _pJ partitions, worker_bwt.erl:48,53 loads
// one): every GPU aligns ALL reads against its partition, and the one exchange step of the path is the gather of the
// per-(read, partition) hit records to the root -- NCCL point-to-point over NVLink / NVSwitch, straight out of the batch's
// device-resident compact result (no host bounce, no torch).  The reference defines no reduction across partitions
// (master.erl:151-155 prints every cigar it receives), so the root gets the partitions' records side by side.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already loaded, if any), so the library
// itself has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <vector>

#include "common.cuh"

// batch.cu
int batch_result_view(bwt_batch *b, int *device, cudaStream_t *stream, int64_t *n, const uint8_t **d_status8, const uint32_t **d_hit_off,
                      const bwt_hit16 **d_hits16, const char **d_cigars, int64_t *n_hits, int64_t *cigar_bytes);

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { bwt_set_error("partition gather: libnccl.so.2 not found (%s)", dlerror()); return nullptr; }
    bool ok = true;
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) ok = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { bwt_set_error("partition gather: libnccl.so.2 lacks a needed symbol"); return nullptr; }
    api.lib = h;
    return &api;
}

#define NCCL_TRY(api, expr)                                                                           \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            bwt_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r));   \
            return BWT_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)
}  // namespace

struct bwt_comm {
    NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long *d_sizes = nullptr;          // world x {n_hits, cigar_bytes}
    long long *h_sizes = nullptr;          // pinned
    uint8_t *d_stage = nullptr;            // root: the other ranks' blocks
    size_t stage_bytes = 0;
};

extern "C" int bwt_comm_unique_id(uint8_t id[128]) {
    if (!id) { bwt_set_error("bwt_comm_unique_id: null argument"); return BWT_ERR_ARG; }
    NcclApi *api = nccl_api();
    if (!api) return BWT_ERR_IO;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    NCCL_TRY(api, api->GetUniqueId(&u));
    memcpy(id, &u, 128);
    return BWT_OK;
}

extern "C" int bwt_comm_create(const uint8_t id[128], int world, int rank, int device, bwt_comm **out) {
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id)) { bwt_set_error("bwt_comm_create: bad argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    bwt_comm *c = new bwt_comm();
    c->world = world; c->rank = rank; c->device = device;
    int st = BWT_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
            cudaEventCreate(&c->ev1) != cudaSuccess || cudaMalloc(&c->d_sizes, (size_t)world * 2 * sizeof(long long)) != cudaSuccess ||
            cudaMallocHost(&c->h_sizes, (size_t)world * 2 * sizeof(long long)) != cudaSuccess) {
            bwt_set_error("bwt_comm_create: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_CUDA; break;
        }
        if (world > 1) {
            c->api = nccl_api();
            if (!c->api) { st = BWT_ERR_IO; break; }
            ncclUniqueId u;
            memcpy(&u, id, 128);
            ncclResult_t r = c->api->CommInitRank(&c->comm, world, u, rank);
            if (r != ncclSuccess) { bwt_set_error("ncclCommInitRank: %s", c->api->GetErrorString(r)); st = BWT_ERR_CUDA; break; }
        }
    } while (0);
    if (st != BWT_OK) { bwt_comm_destroy(c); return st; }
    *out = c;
    return BWT_OK;
}

extern "C" int bwt_comm_destroy(bwt_comm *c) {
    if (!c) return BWT_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->api) c->api->CommDestroy(c->comm);
    cudaFree(c->d_sizes); cudaFree(c->d_stage);
    if (c->h_sizes) cudaFreeHost(c->h_sizes);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BWT_OK;
}

// Every rank calls this with its batch (compact result of the SAME n reads against its own partition).  Exchange:
// an all-gather of the two sizes, then every other rank sends status | offsets | hits | CIGAR bytes to the root, which
// receives them into a device staging area and copies everything to the host outputs, rank after rank:
//   status_out[p*n ..], hit_off_out[p*(n+1) ..] (rank p's own offsets, starting at 0),
//   hits_out / cig_out: rank p's records at the running sums of part_hits / part_cig.
// part_hits / part_cig (world entries) come back on EVERY rank; the other outputs only on the root (may be NULL elsewhere).
// ms_gather: CUDA-event time of the exchange on the communicator's stream; link_bytes: bytes that crossed NVLink into the root.
extern "C" int bwt_partition_gather(bwt_comm *c, bwt_batch *b, int root, uint8_t *status_out, uint32_t *hit_off_out,
                                    bwt_hit16 *hits_out, int64_t hits_cap, char *cig_out, int64_t cig_cap, int64_t *part_hits,
                                    int64_t *part_cig, float *ms_gather, int64_t *link_bytes) {
    if (!c || !b || root < 0 || root >= c->world || !part_hits || !part_cig) { bwt_set_error("bwt_partition_gather: bad argument"); return BWT_ERR_ARG; }
    int device;
    cudaStream_t bstream;
    int64_t n, nh, nc;
    const uint8_t *d_st; const uint32_t *d_off; const bwt_hit16 *d_hits; const char *d_cig;
    BWT_TRY(batch_result_view(b, &device, &bstream, &n, &d_st, &d_off, &d_hits, &d_cig, &nh, &nc));
    if (device != c->device) { bwt_set_error("bwt_partition_gather: batch and communicator are on different devices"); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(c->device));
    const int W = c->world;
    const bool is_root = c->rank == root;
    cudaStream_t s = c->stream;
    CU_TRY(cudaEventRecord(c->ev0, s));
    // 1. sizes
    c->h_sizes[2 * c->rank] = (long long)nh; c->h_sizes[2 * c->rank + 1] = (long long)nc;
    CU_TRY(cudaMemcpyAsync(c->d_sizes + 2 * c->rank, c->h_sizes + 2 * c->rank, 2 * sizeof(long long), cudaMemcpyHostToDevice, s));
    if (W > 1) NCCL_TRY(c->api, c->api->AllGather(c->d_sizes + 2 * c->rank, c->d_sizes, 2 * sizeof(long long), ncclChar, c->comm, s));
    CU_TRY(cudaMemcpyAsync(c->h_sizes, c->d_sizes, (size_t)W * 2 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    int64_t tot_h = 0, tot_c = 0;
    for (int p = 0; p < W; p++) { part_hits[p] = c->h_sizes[2 * p]; part_cig[p] = c->h_sizes[2 * p + 1]; tot_h += part_hits[p]; tot_c += part_cig[p]; }
    // 2. records: rank p's block = status (n) | offsets (4(n+1)) | hits (16 nh_p) | cigars (nc_p), each part 16-byte aligned
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    auto block_bytes = [&](int p) { return up16((size_t)n) + up16(4 * (size_t)(n + 1)) + up16(16 * (size_t)part_hits[p]) + up16((size_t)part_cig[p]); };
    int64_t moved = 0;
    if (W > 1) {
        if (is_root) {
            size_t need = 0;
            for (int p = 0; p < W; p++) if (p != root) need += block_bytes(p);
            if (need > c->stage_bytes) {
                cudaFree(c->d_stage); c->d_stage = nullptr;
                c->stage_bytes = need + need / 4 + 4096;
                CU_TRY(cudaMalloc(&c->d_stage, c->stage_bytes));
            }
        }
        NCCL_TRY(c->api, c->api->GroupStart());
        if (is_root) {
            size_t o = 0;
            for (int p = 0; p < W; p++) {
                if (p == root) continue;
                uint8_t *base = c->d_stage + o;
                NCCL_TRY(c->api, c->api->Recv(base, (size_t)n, ncclChar, p, c->comm, s)); base += up16((size_t)n);
                NCCL_TRY(c->api, c->api->Recv(base, 4 * (size_t)(n + 1), ncclChar, p, c->comm, s)); base += up16(4 * (size_t)(n + 1));
                if (part_hits[p]) NCCL_TRY(c->api, c->api->Recv(base, 16 * (size_t)part_hits[p], ncclChar, p, c->comm, s));
                base += up16(16 * (size_t)part_hits[p]);
                if (part_cig[p]) NCCL_TRY(c->api, c->api->Recv(base, (size_t)part_cig[p], ncclChar, p, c->comm, s));
                o += block_bytes(p);
                moved += n + 4 * (n + 1) + 16 * part_hits[p] + part_cig[p];
            }
        } else {
            NCCL_TRY(c->api, c->api->Send(d_st, (size_t)n, ncclChar, root, c->comm, s));
            NCCL_TRY(c->api, c->api->Send(d_off, 4 * (size_t)(n + 1), ncclChar, root, c->comm, s));
            if (nh) NCCL_TRY(c->api, c->api->Send(d_hits, 16 * (size_t)nh, ncclChar, root, c->comm, s));
            if (nc) NCCL_TRY(c->api, c->api->Send(d_cig, (size_t)nc, ncclChar, root, c->comm, s));
        }
        NCCL_TRY(c->api, c->api->GroupEnd());
    }
    CU_TRY(cudaEventRecord(c->ev1, s));
    // 3. root: everything to the host outputs
    int st = BWT_OK;
    if (is_root) {
        if (!status_out || !hit_off_out || (tot_h > 0 && !hits_out) || (tot_c > 0 && !cig_out) || tot_h > hits_cap || tot_c > cig_cap) {
            bwt_set_error("bwt_partition_gather: output capacity too small: need %lld hits and %lld CIGAR bytes", (long long)tot_h, (long long)tot_c);
            st = BWT_ERR_NOSPACE;
        } else {
            size_t o = 0;
            int64_t hb = 0, cb = 0;
            for (int p = 0; p < W; p++) {
                const uint8_t *src_st; const void *src_off, *src_h, *src_c;
                if (p == root) { src_st = d_st; src_off = d_off; src_h = d_hits; src_c = d_cig; }
                else {
                    const uint8_t *base = c->d_stage + o;
                    src_st = base; base += up16((size_t)n);
                    src_off = base; base += up16(4 * (size_t)(n + 1));
                    src_h = base; base += up16(16 * (size_t)part_hits[p]);
                    src_c = base;
                    o += block_bytes(p);
                }
                if (n) CU_TRY(cudaMemcpyAsync(status_out + (size_t)p * n, src_st, (size_t)n, cudaMemcpyDeviceToHost, s));
                CU_TRY(cudaMemcpyAsync(hit_off_out + (size_t)p * (n + 1), src_off, 4 * (size_t)(n + 1), cudaMemcpyDeviceToHost, s));
                if (part_hits[p]) CU_TRY(cudaMemcpyAsync(hits_out + hb, src_h, 16 * (size_t)part_hits[p], cudaMemcpyDeviceToHost, s));
                if (part_cig[p]) CU_TRY(cudaMemcpyAsync(cig_out + cb, src_c, (size_t)part_cig[p], cudaMemcpyDeviceToHost, s));
                hb += part_hits[p]; cb += part_cig[p];
            }
        }
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (ms_gather) { float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); *ms_gather = ms; }
    if (link_bytes) *link_bytes = moved;
    return st;
}
