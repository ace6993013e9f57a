// engine.cu -- the C ABI of libfhestr_engine.so (include/fhestr_engine.h): device-resident key store,
// ciphertext arena, LUT registry and the batched-PBS launch path.  No CPU fallback anywhere: if CUDA
// is not usable every entry point fails with an error code.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace fhestr;

static thread_local std::string g_create_error;

struct fhestr_engine {
    fhestr_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    u64* arena = nullptr;
    uint64_t arena_blocks = 0;
    bool own_arena = false;
    cplx* bsk_f = nullptr;
    u64* ksk = nullptr;
    u64* ksk_corr = nullptr;
    unsigned char* ksk8 = nullptr;     // limb matrix for the tensor-core keyswitch
    unsigned char* ks_digits = nullptr;
    u64* ks_body = nullptr;
    int ks_path = 0;                   // 0 = tensor cores (tcgen05 kind::i8), 1 = CUDA cores (u64 IMAD)
    cplx* tf = nullptr;
    cplx* bsk_w = nullptr;     // the same key in the spectrum order of the latency kernel (br_wide.cuh)
    WideConsts* wide_tab = nullptr;   // [kWT] per-thread transform constants of the latency kernel
    int n_sms = 148;
    bool keys_loaded = false;
    u64* luts = nullptr;
    u64* lut_post = nullptr;           // [cap_luts]: delta/2 for half-step tables, else 0
    int n_luts = 0, cap_luts = 256;
    std::vector<std::vector<uint8_t>> lut_tables;   // host copy: identical tables share one id
    // per-batch scratch (grown on demand)
    fhestr_job* d_jobs = nullptr;
    size_t jobs_cap = 0;
    u64* ks_out = nullptr;
    size_t ks_cap = 0;
    uint8_t* d_bytes = nullptr;
    size_t bytes_cap = 0;
    u64* gather_buf = nullptr;         // staging for fhestr_ct_download_slots
    size_t gather_cap = 0;
    int br_mode = 0;                   // 0 = by level size, 1 = throughput kernel only, 2 = latency kernel only
    int wide_max_jobs = 0;             // levels of at most this many PBS jobs run on the latency kernel (0 = 3 x SMs)
    int wide_shape = 0;                // 0 = by level size, 1 = one PBS per SM, 2 = the pair form (two PBS per SM)
    uint64_t launches = 0;
    bool timing = false;
    struct Timed { cudaEvent_t a, b, c; uint32_t pbs; };   // a..b keyswitch, b..c blind rotation
    std::vector<Timed> timed;
    // multi-GPU, preferred path: peer arenas mapped with cudaIpc; the blind-rotation epilogue stores every result
    // into all arenas and a flag barrier over peer memory closes the level (no collective on the data path)
    u64* peer_arena[8] = {};          // by rank; [rank] = own arena
    uint32_t* peer_flags[8] = {};     // by rank; [rank] = own flag array
    uint32_t* my_flags = nullptr;     // [8] epochs written by the peers + [8] = status word
    bool peers_attached = false;
    uint32_t epoch = 0;
    // multi-GPU (one process per GPU): NCCL communicator, resolved at run time from libnccl.so.2
    void* comm = nullptr;
    uint32_t rank = 0, world = 1;
    std::string err;
};

// ---- NCCL, bound with dlopen so that the library has no link-time dependency on it: inside a torch
// process this resolves to the libnccl.so.2 torch already loaded, otherwise to the system one.
namespace {
struct NcclUid { char b[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
    return api;
}
constexpr int kNcclUint64 = 5;  // ncclUint64 in nccl.h
}  // namespace

struct fhestr_program {
    fhestr_engine* eng = nullptr;
    fhestr_job* d_jobs = nullptr;          // all jobs, PBS jobs of a level first, then its leveled jobs
    std::vector<uint32_t> level_off;       // n_levels + 1
    std::vector<uint32_t> level_pbs;       // PBS jobs in each level (the rest are leveled)
    std::vector<uint32_t> level_first_dst; // dst of the level's first PBS job
    std::vector<uint8_t> level_contiguous; // PBS job i of the level writes first_dst + i (needed to shard it)
    uint32_t max_level_pbs = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            char _b[512];                                                                          \
            snprintf(_b, sizeof _b, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                     cudaGetErrorString(_e));                                                      \
            e->err = _b;                                                                           \
            return FHESTR_E_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

static int fail(fhestr_engine* e, int code, const std::string& msg) {
    e->err = msg;
    return code;
}

// After a synchronising call: a peer flag barrier that timed out (a peer died or fell behind by more than the spin
// bound) leaves a non-zero status word; the results of that run are incomplete, so the call fails instead of
// returning them.  The word stays set until fhestr_peer_detach.
static int check_peer_status(fhestr_engine* e) {
    if (!e->peers_attached || !e->my_flags) return FHESTR_OK;
    uint32_t st = 0;
    CK(cudaMemcpyAsync(&st, e->my_flags + 8, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (st) return fail(e, FHESTR_E_STATE, "a multi-GPU level barrier timed out: peer results are incomplete (fhestr_peer_status)");
    return FHESTR_OK;
}

// Grow the per-batch scratch.  A failed allocation leaves the capacity at 0 and the pointers null, so the next call
// allocates again instead of running on freed memory.
static int ensure_scratch(fhestr_engine* e, size_t n_jobs) {
    if (n_jobs > e->jobs_cap) {
        e->jobs_cap = 0;
        if (e->d_jobs) { CK(cudaFree(e->d_jobs)); e->d_jobs = nullptr; }
        const size_t cap = n_jobs * 2 + 64;
        CK(cudaMalloc(&e->d_jobs, cap * sizeof(fhestr_job)));
        e->jobs_cap = cap;
    }
    if (n_jobs > e->ks_cap) {
        e->ks_cap = 0;
        if (e->ks_out) { CK(cudaFree(e->ks_out)); e->ks_out = nullptr; }
        if (e->ks_digits) { CK(cudaFree(e->ks_digits)); e->ks_digits = nullptr; }
        if (e->ks_body) { CK(cudaFree(e->ks_body)); e->ks_body = nullptr; }
        const size_t cap = n_jobs * 2 + 64;
        CK(cudaMalloc(&e->ks_out, cap * (size_t)(e->prm.n + 1) * sizeof(u64)));
        CK(cudaMalloc(&e->ks_digits, ks_digit_rows(cap) * (size_t)kN * e->prm.ks_level));
        // on the engine stream: the stream is non-blocking, so a legacy-stream memset is NOT ordered before its kernels
        CK(cudaMemsetAsync(e->ks_digits, 0, ks_digit_rows(cap) * (size_t)kN * e->prm.ks_level, e->stream));
        CK(cudaMalloc(&e->ks_body, cap * sizeof(u64)));
        e->ks_cap = cap;
    }
    return FHESTR_OK;
}

static int validate_jobs(fhestr_engine* e, const fhestr_job* jobs, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const fhestr_job& j = jobs[i];
        if (j.n_terms > FHESTR_MAX_TERMS) return fail(e, FHESTR_E_INVALID, "job has too many terms");
        if (j.dst >= e->arena_blocks) return fail(e, FHESTR_E_INVALID, "job dst outside the arena");
        if (j.lut >= e->n_luts) return fail(e, FHESTR_E_INVALID, "job uses an unregistered LUT");
        for (uint32_t t = 0; t < j.n_terms; t++)
            if (j.src[t] >= e->arena_blocks) return fail(e, FHESTR_E_INVALID, "job src outside the arena");
    }
    return FHESTR_OK;
}

#pragma GCC visibility push(default)
extern "C" {

const char* fhestr_last_error(const fhestr_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int fhestr_engine_create(const fhestr_params* p, int device, uint64_t arena_blocks, void* external_arena,
                         fhestr_engine** out) {
    if (!p || !out) { g_create_error = "null argument"; return FHESTR_E_INVALID; }
    *out = nullptr;
    if (p->N != kN || p->k != 1 || p->pbs_level != 1 || p->pbs_base_log != kPbsBaseLog || p->n < 1 ||
        p->n > 767 || p->ks_level < 1 || p->ks_level > 8 || p->ks_base_log < 1 ||
        (p->ks_base_log + 1) * p->ks_level > 32 || p->delta_log < 48 || p->delta_log > 62) {
        g_create_error = "unsupported parameter set (need N=2048, k=1, pbs 1x23 bits, n<=767, (ks_base_log+1)*ks_level<=32)";
        return FHESTR_E_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        g_create_error = "no usable CUDA device (this engine has no CPU fallback)";
        return FHESTR_E_NOGPU;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        g_create_error = "device is not sm_100-class; the kernels are built for sm_100a only";
        return FHESTR_E_NOGPU;
    }
    fhestr_engine* e = new (std::nothrow) fhestr_engine();
    if (!e) { g_create_error = "out of host memory"; return FHESTR_E_STATE; }
    e->prm = *p;
    e->device = device;
    e->arena_blocks = arena_blocks;
    auto bail = [&](int code) { g_create_error = e->err; fhestr_engine_destroy(e); return code; };
#define CKC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string(#call) + ": " + cudaGetErrorString(_e); return bail(FHESTR_E_CUDA); } } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    if (external_arena) e->arena = static_cast<u64*>(external_arena);
    else {
        CKC(cudaMalloc(&e->arena, arena_blocks * (size_t)(kN + 1) * sizeof(u64)));
        e->own_arena = true;
        CKC(cudaMemsetAsync(e->arena, 0, arena_blocks * (size_t)(kN + 1) * sizeof(u64), e->stream));
    }
    CKC(cudaMalloc(&e->tf, 1024 * sizeof(cplx)));
    {
        std::vector<cplx> tf(1024);
        make_twiddles(tf.data());
        CKC(cudaMemcpyAsync(e->tf, tf.data(), 1024 * sizeof(cplx), cudaMemcpyHostToDevice, e->stream));
        CKC(cudaStreamSynchronize(e->stream));
    }
    e->n_sms = prop.multiProcessorCount;
    CKC(cudaMalloc(&e->wide_tab, kWT * sizeof(WideConsts)));
    {
        std::vector<WideConsts> t(kWT);
        make_wide_consts(t.data());
        CKC(cudaMemcpyAsync(e->wide_tab, t.data(), t.size() * sizeof(WideConsts), cudaMemcpyHostToDevice, e->stream));
        CKC(cudaStreamSynchronize(e->stream));
    }
    CKC(cudaMalloc(&e->bsk_w, (size_t)p->n * kWKeyTile * sizeof(cplx)));
    CKC(blind_rotate_wide_configure());
    CKC(cudaMalloc(&e->luts, (size_t)e->cap_luts * kN * sizeof(u64)));
    CKC(cudaMalloc(&e->lut_post, (size_t)e->cap_luts * sizeof(u64)));
    CKC(cudaMemsetAsync(e->lut_post, 0, (size_t)e->cap_luts * sizeof(u64), e->stream));
    CKC(cudaMalloc(&e->bsk_f, (size_t)p->n * kBskStepElems * sizeof(cplx)));
    CKC(cudaMalloc(&e->ksk, (size_t)kN * p->ks_level * (p->n + 1) * sizeof(u64)));
    CKC(cudaMalloc(&e->ksk_corr, (size_t)(p->n + 1) * sizeof(u64)));
    CKC(cudaMalloc(&e->ksk8, (size_t)ks_cols_padded(p->n) * 8 * kN * p->ks_level));
    CKC(keyswitch_tc_configure());
    if (p->ks_base_log > 7) e->ks_path = 1;   // unsigned digits must fit a byte
    CKC(blind_rotate_configure());
    CKC(keyswitch_configure());
    CKC(cudaStreamSynchronize(e->stream));
#undef CKC
    *out = e;
    return FHESTR_OK;
}

void fhestr_engine_destroy(fhestr_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->own_stream) cudaStreamSynchronize(e->own_stream);
    if (e->comm) fhestr_comm_destroy(e);
    if (e->peers_attached) fhestr_peer_detach(e);
    cudaFree(e->my_flags);
    if (e->own_arena && e->arena) cudaFree(e->arena);
    cudaFree(e->bsk_f); cudaFree(e->ksk); cudaFree(e->ksk_corr); cudaFree(e->tf);
    cudaFree(e->bsk_w); cudaFree(e->wide_tab); cudaFree(e->ksk8); cudaFree(e->ks_digits); cudaFree(e->ks_body);
    cudaFree(e->gather_buf); cudaFree(e->luts); cudaFree(e->lut_post); cudaFree(e->d_jobs); cudaFree(e->ks_out); cudaFree(e->d_bytes);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

int fhestr_set_stream(fhestr_engine* e, void* cuda_stream) {
    if (!e) return FHESTR_E_INVALID;
    e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
    return FHESTR_OK;
}

int fhestr_sync(fhestr_engine* e) {
    if (!e) return FHESTR_E_INVALID;
    CK(cudaStreamSynchronize(e->stream));
    return check_peer_status(e);
}

void* fhestr_arena_ptr(fhestr_engine* e) { return e ? e->arena : nullptr; }

int fhestr_load_keys(fhestr_engine* e, const uint64_t* bsk_std, const uint64_t* ksk) {
    if (!e || !bsk_std || !ksk) return e ? fail(e, FHESTR_E_INVALID, "null key pointer") : FHESTR_E_INVALID;
    CK(cudaSetDevice(e->device));
    const fhestr_params& p = e->prm;
    const size_t bsk_words = (size_t)p.n * 4 * kN;
    const size_t ksk_words = (size_t)kN * p.ks_level * (p.n + 1);
    u64* d_std = nullptr;
    CK(cudaMalloc(&d_std, bsk_words * sizeof(u64)));
    CK(cudaMemcpyAsync(d_std, bsk_std, bsk_words * sizeof(u64), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->ksk, ksk, ksk_words * sizeof(u64), cudaMemcpyHostToDevice, e->stream));
    e->launches += launch_bsk_convert(d_std, p.n, e->tf, e->bsk_f, e->stream);
    e->launches += launch_bsk_convert_wide(d_std, p.n, e->wide_tab, e->bsk_w, e->stream);
    e->launches += launch_ksk_correction(e->ksk, kN * p.ks_level, p.n, p.ks_base_log, e->ksk_corr, e->stream);
    e->launches += launch_ksk_limbs(e->ksk, kN * p.ks_level, p.n, e->ksk8, e->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaFree(d_std));
    e->keys_loaded = true;
    return FHESTR_OK;
}

static int stage_bytes(fhestr_engine* e, const uint8_t* host, size_t n) {
    if (n > e->bytes_cap) {
        if (e->d_bytes) CK(cudaFree(e->d_bytes));
        e->bytes_cap = n * 2 + 256;
        CK(cudaMalloc(&e->d_bytes, e->bytes_cap));
    }
    CK(cudaMemcpyAsync(e->d_bytes, host, n, cudaMemcpyHostToDevice, e->stream));
    return FHESTR_OK;
}

int fhestr_lut_register(fhestr_engine* e, const uint8_t* table, int32_t* lut_id) {
    if (!e || !table || !lut_id) return FHESTR_E_INVALID;
    CK(cudaSetDevice(e->device));
    const int entries = 1 << (63 - e->prm.delta_log);
    const std::vector<uint8_t> key(table, table + entries);
    int flagged = 0;
    for (int i = 0; i < entries; i++) {
        flagged += (table[i] & 0x80) ? 1 : 0;
        if ((table[i] & 0x7f) >= entries) return fail(e, FHESTR_E_INVALID, "LUT entry outside the block's value range");
    }
    if (flagged && flagged != entries) return fail(e, FHESTR_E_INVALID, "a half-step LUT must flag every entry with 0x80");
    for (int i = 0; i < e->n_luts; i++)
        if (e->lut_tables[i] == key) { *lut_id = i; return FHESTR_OK; }
    if (e->n_luts >= e->cap_luts) {  // grow the registry (ids stay valid)
        u64* bigger = nullptr;
        CK(cudaMalloc(&bigger, (size_t)e->cap_luts * 2 * kN * sizeof(u64)));
        CK(cudaMemcpyAsync(bigger, e->luts, (size_t)e->n_luts * kN * sizeof(u64), cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaFree(e->luts));
        e->luts = bigger;
        u64* bigger_post = nullptr;
        CK(cudaMalloc(&bigger_post, (size_t)e->cap_luts * 2 * sizeof(u64)));
        CK(cudaMemsetAsync(bigger_post, 0, (size_t)e->cap_luts * 2 * sizeof(u64), e->stream));
        CK(cudaMemcpyAsync(bigger_post, e->lut_post, (size_t)e->n_luts * sizeof(u64), cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaFree(e->lut_post));
        e->lut_post = bigger_post;
        e->cap_luts *= 2;
    }
    int rc = stage_bytes(e, table, entries);
    if (rc) return rc;
    e->launches += launch_lut_poly(e->d_bytes, entries, e->prm.delta_log, e->luts + (size_t)e->n_luts * kN, e->stream);
    CK(cudaGetLastError());
    const u64 post = flagged ? (u64)1 << (e->prm.delta_log - 1) : 0;
    CK(cudaMemcpyAsync(e->lut_post + e->n_luts, &post, sizeof post, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // d_bytes is reused by the next call
    e->lut_tables.push_back(key);
    *lut_id = e->n_luts++;
    return FHESTR_OK;
}

int fhestr_lut_download(fhestr_engine* e, int32_t lut_id, uint64_t* out_poly) {
    if (!e || !out_poly || lut_id < 0 || lut_id >= e->n_luts) return e ? fail(e, FHESTR_E_INVALID, "bad lut id") : FHESTR_E_INVALID;
    CK(cudaMemcpyAsync(out_poly, e->luts + (size_t)lut_id * kN, kN * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return FHESTR_OK;
}

int fhestr_ct_upload(fhestr_engine* e, uint32_t first, uint32_t count, const uint64_t* host) {
    if (!e || !host) return FHESTR_E_INVALID;
    if ((uint64_t)first + count > e->arena_blocks) return fail(e, FHESTR_E_STATE, "upload outside the arena");
    CK(cudaMemcpyAsync(e->arena + (size_t)first * (kN + 1), host, (size_t)count * (kN + 1) * sizeof(u64),
                       cudaMemcpyHostToDevice, e->stream));
    return FHESTR_OK;
}

int fhestr_ct_download(fhestr_engine* e, uint32_t first, uint32_t count, uint64_t* host) {
    if (!e || !host) return FHESTR_E_INVALID;
    if ((uint64_t)first + count > e->arena_blocks) return fail(e, FHESTR_E_STATE, "download outside the arena");
    CK(cudaMemcpyAsync(host, e->arena + (size_t)first * (kN + 1), (size_t)count * (kN + 1) * sizeof(u64),
                       cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return check_peer_status(e);
}

int fhestr_ct_download_slots(fhestr_engine* e, const uint32_t* slots, uint32_t count, uint64_t* host) {
    if (!e || !slots || !host) return FHESTR_E_INVALID;
    if (!count) return FHESTR_OK;
    for (uint32_t i = 0; i < count; i++)
        if (slots[i] >= e->arena_blocks) return fail(e, FHESTR_E_STATE, "download outside the arena");
    CK(cudaSetDevice(e->device));
    const size_t bytes = (size_t)count * (kN + 1) * sizeof(u64);
    if (bytes > e->gather_cap) {
        if (e->gather_buf) { CK(cudaFree(e->gather_buf)); e->gather_buf = nullptr; e->gather_cap = 0; }
        CK(cudaMalloc(&e->gather_buf, bytes * 2));
        e->gather_cap = bytes * 2;
    }
    int rc = stage_bytes(e, reinterpret_cast<const uint8_t*>(slots), (size_t)count * sizeof(uint32_t));
    if (rc) return rc;
    e->launches += launch_gather_blocks(e->arena, reinterpret_cast<const uint32_t*>(e->d_bytes), count, e->gather_buf, e->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, e->gather_buf, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return check_peer_status(e);
}

int fhestr_ct_download_async(fhestr_engine* e, uint32_t first, uint32_t count, uint64_t* pinned_host) {
    if (!e || !pinned_host) return FHESTR_E_INVALID;
    if ((uint64_t)first + count > e->arena_blocks) return fail(e, FHESTR_E_STATE, "download outside the arena");
    CK(cudaMemcpyAsync(pinned_host, e->arena + (size_t)first * (kN + 1), (size_t)count * (kN + 1) * sizeof(u64),
                       cudaMemcpyDeviceToHost, e->stream));
    return FHESTR_OK;
}

int fhestr_ct_trivial(fhestr_engine* e, uint32_t first, uint32_t count, const uint8_t* values) {
    if (!e || !values) return FHESTR_E_INVALID;
    if ((uint64_t)first + count > e->arena_blocks) return fail(e, FHESTR_E_STATE, "trivial outside the arena");
    int rc = stage_bytes(e, values, count);
    if (rc) return rc;
    e->launches += launch_trivial(e->arena, first, count, e->d_bytes, e->prm.delta_log, e->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    return FHESTR_OK;
}

// Which blind-rotation kernel runs a level: a level with few jobs is bound by the serial chain of 742 CMUX steps of ONE
// PBS, so it runs on the latency kernel (one PBS per SM, 128 threads, br_wide.cuh); a level with many jobs is bound by
// FP64 issue and runs on the throughput kernel (four PBS per SM, br_core.cuh).
static bool use_wide(const fhestr_engine* e, uint32_t n_pbs) {
    if (e->br_mode == 1) return false;
    if (e->br_mode == 2) return true;
    const uint32_t lim = e->wide_max_jobs > 0 ? (uint32_t)e->wide_max_jobs : 3u * (uint32_t)e->n_sms;
    return n_pbs <= lim;
}

// the latency kernel's two forms: ONE PBS per SM (128 threads, two-tile key ring) for levels of at most one job per SM,
// TWO PBS per SM (256 threads, one shared key tile) up to two jobs per SM; a level between two and three jobs per SM
// runs as one full wave of pairs followed by one wave of singles (3.55 + 2.14 ms against 6.0-6.8 ms on the
// throughput kernel, profiles/r2_level_latency.md)
static BrBatchArgs br_slice(const BrBatchArgs& br, uint32_t first, uint32_t count) {
    BrBatchArgs s = br;
    s.ks = br.ks + (size_t)first * (br.n + 1);
    if (br.jobs) s.jobs = br.jobs + first;
    if (br.lut_ids) s.lut_ids = br.lut_ids + first;
    if (br.init_acc) s.init_acc = br.init_acc + (size_t)first * 2 * kN;
    if (br.out_acc) s.out_acc = br.out_acc + (size_t)first * 2 * kN;
    s.B = (int)count;
    return s;
}
static int launch_wide(const fhestr_engine* e, const BrBatchArgs& br, uint32_t n_pbs) {
    const uint32_t sms = (uint32_t)e->n_sms;
    if (e->wide_shape == 1 || (e->wide_shape == 0 && n_pbs <= sms)) return launch_blind_rotate_wide(br, e->stream);
    if (e->wide_shape == 2 || n_pbs <= 2 * sms) return launch_blind_rotate_wide2(br, e->stream);
    return launch_blind_rotate_wide2(br_slice(br, 0, 2 * sms), e->stream) +
           launch_blind_rotate_wide(br_slice(br, 2 * sms, n_pbs - 2 * sms), e->stream);
}

// Blind rotation of one level.  Above three jobs per SM the throughput kernel runs it in waves of four PBS per SM; a last
// wave that is mostly empty still costs a whole PBS chain on that kernel (626 jobs = one wave + 34: 9.95 ms), so a
// small remainder goes to the latency kernel instead, right behind the full waves (8.85 ms).  Measured on one B200
// (profiles/r2b_tail_remainder.md): a remainder of at most one job per SM pays up to three full waves (1332 jobs: 16.1
// -> 15.1 ms; from four waves on the CTAs of a launch have drifted apart enough to hide their own tail), a remainder of
// up to three jobs per SM only behind a single full wave (950 jobs: 12.7 -> 12.4 ms).
static int launch_level_br(const fhestr_engine* e, const BrBatchArgs& br, uint32_t n_pbs) {
    if (use_wide(e, n_pbs)) return launch_wide(e, br, n_pbs);
    if (e->br_mode == 0) {
        const uint32_t sms = (uint32_t)e->n_sms, wave = 4 * sms;
        const uint32_t full = n_pbs / wave, r = n_pbs % wave;
        if (r > 0 && ((full >= 1 && full <= 3 && r <= sms) || (full == 1 && r <= 3 * sms)))
            return launch_blind_rotate(br_slice(br, 0, n_pbs - r), e->stream) + launch_wide(e, br_slice(br, n_pbs - r, r), r);
    }
    return launch_blind_rotate(br, e->stream);
}

// launch one level: jobs[0, n_pbs) are PBS jobs, jobs[n_pbs, n_all) leveled-only
static int run_level(fhestr_engine* e, const fhestr_job* d_jobs, uint32_t n_pbs, uint32_t n_all, bool peer_stores = false) {
    if (n_pbs) {
        fhestr_engine::Timed t{};
        if (e->timing) {
            CK(cudaEventCreate(&t.a)); CK(cudaEventCreate(&t.b)); CK(cudaEventCreate(&t.c));
            t.pbs = n_pbs;
            CK(cudaEventRecord(t.a, e->stream));
        }
        KsBatchArgs ks{d_jobs, e->arena, e->ksk, e->ksk_corr, e->ks_out, e->prm.n, (int)n_pbs,
                       e->prm.ks_base_log, e->prm.ks_level, e->ksk8, e->ks_digits, e->ks_body};
        const int ks_launches = e->ks_path == 0 ? launch_keyswitch_tc(ks, e->stream) : launch_keyswitch(ks, e->stream);
        if (ks_launches == 0 && n_pbs > 0) return fail(e, FHESTR_E_CUDA, "keyswitch: the TMA tensor maps could not be encoded");
        e->launches += ks_launches;
        if (e->timing) CK(cudaEventRecord(t.b, e->stream));
        BrBatchArgs br{};
        br.ks = e->ks_out; br.luts = e->luts; br.lut_post = e->lut_post; br.lut_ids = nullptr; br.jobs = d_jobs; br.arena = e->arena;
        br.bsk = e->bsk_f; br.tf = e->tf; br.n = e->prm.n; br.B = (int)n_pbs;
        br.bsk_w = e->bsk_w; br.wide_tab = e->wide_tab;
        br.n_peers = 0;
        if (e->peers_attached && peer_stores) {
            for (uint32_t r = 0; r < e->world; r++) if (r != e->rank) br.peer_arena[br.n_peers++] = e->peer_arena[r];
        }
        e->launches += launch_level_br(e, br, n_pbs);
        if (e->timing) { CK(cudaEventRecord(t.c, e->stream)); e->timed.push_back(t); }
    }
    if (n_all > n_pbs) e->launches += launch_linear(d_jobs + n_pbs, (int)(n_all - n_pbs), e->arena, e->stream);
    CK(cudaGetLastError());
    return FHESTR_OK;
}

// stable partition: PBS jobs first, leveled jobs after; returns the number of PBS jobs
static uint32_t partition_jobs(const fhestr_job* in, uint32_t n, fhestr_job* out) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; i++) if (in[i].lut >= 0) out[k++] = in[i];
    const uint32_t n_pbs = k;
    for (uint32_t i = 0; i < n; i++) if (in[i].lut < 0) out[k++] = in[i];
    return n_pbs;
}

int fhestr_pbs_batch(fhestr_engine* e, const fhestr_job* jobs, uint32_t n_jobs) {
    if (!e || (!jobs && n_jobs)) return FHESTR_E_INVALID;
    if (!e->keys_loaded) return fail(e, FHESTR_E_STATE, "keys not loaded");
    if (!n_jobs) return FHESTR_OK;
    int rc = validate_jobs(e, jobs, n_jobs);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    rc = ensure_scratch(e, n_jobs);
    if (rc) return rc;
    std::vector<fhestr_job> sorted(n_jobs);
    const uint32_t n_pbs = partition_jobs(jobs, n_jobs, sorted.data());
    CK(cudaMemcpyAsync(e->d_jobs, sorted.data(), (size_t)n_jobs * sizeof(fhestr_job), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // `sorted` is pageable host memory
    return run_level(e, e->d_jobs, n_pbs, n_jobs);
}

int fhestr_program_create(fhestr_engine* e, const fhestr_job* jobs, const uint32_t* level_offsets,
                          uint32_t n_levels, fhestr_program** out) {
    if (!e || !level_offsets || !out) return FHESTR_E_INVALID;
    const uint32_t total = level_offsets[n_levels];
    if (!jobs && total) return fail(e, FHESTR_E_INVALID, "null job list");
    int rc = validate_jobs(e, jobs, total);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    fhestr_program* p = new (std::nothrow) fhestr_program();
    if (!p) return fail(e, FHESTR_E_STATE, "out of host memory");
    p->eng = e;
    p->level_off.assign(level_offsets, level_offsets + n_levels + 1);
    p->level_pbs.resize(n_levels);
    std::vector<fhestr_job> sorted(total ? total : 1);
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t a = level_offsets[l], b = level_offsets[l + 1];
        p->level_pbs[l] = partition_jobs(jobs + a, b - a, sorted.data() + a);
        if (p->level_pbs[l] > p->max_level_pbs) p->max_level_pbs = p->level_pbs[l];
        uint8_t contiguous = 1;
        for (uint32_t i = 1; i < p->level_pbs[l]; i++)
            if (sorted[a + i].dst != sorted[a].dst + i) { contiguous = 0; break; }
        p->level_first_dst.push_back(p->level_pbs[l] ? sorted[a].dst : 0u);
        p->level_contiguous.push_back(contiguous);
    }
    cudaError_t ce = cudaMalloc(&p->d_jobs, (size_t)(total ? total : 1) * sizeof(fhestr_job));
    // The upload goes on the ENGINE stream and is waited for.  A legacy-stream cudaMemcpy from pageable memory
    // returns once the data are staged (the DMA may still be in flight), and the engine stream is non-blocking, so the
    // first level's kernels could read whatever a recycled allocation still held: a stale job list.
    if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(p->d_jobs, sorted.data(), (size_t)total * sizeof(fhestr_job), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) {
        cudaFree(p->d_jobs);
        delete p;
        return fail(e, FHESTR_E_CUDA, std::string("program upload: ") + cudaGetErrorString(ce));
    }
    *out = p;
    return FHESTR_OK;
}

int fhestr_program_level_jobs(const fhestr_program* p, uint32_t level, uint32_t* n_jobs) {
    if (!p || !n_jobs || level + 1 >= p->level_off.size()) return FHESTR_E_INVALID;
    *n_jobs = p->level_off[level + 1] - p->level_off[level];
    return FHESTR_OK;
}

int fhestr_program_run(fhestr_engine* e, fhestr_program* p, uint32_t first_level, uint32_t last_level,
                       uint32_t rank, uint32_t world) {
    if (!e || !p || p->eng != e || world == 0 || rank >= world) return FHESTR_E_INVALID;
    if (!e->keys_loaded) return fail(e, FHESTR_E_STATE, "keys not loaded");
    if (last_level > p->level_pbs.size() || first_level > last_level) return fail(e, FHESTR_E_INVALID, "bad level range");
    CK(cudaSetDevice(e->device));
    int rc = ensure_scratch(e, p->max_level_pbs);
    if (rc) return rc;
    if (world > 1 && ((!e->comm && !e->peers_attached) || e->world != world || e->rank != rank))
        return fail(e, FHESTR_E_STATE, "multi-rank run needs fhestr_peer_attach or fhestr_comm_init with the same rank/world");
    if (world > 1 && e->peers_attached) {
        // Nobody stores into a peer's arena before that peer has ENTERED this run: in its stream order that is after
        // everything it did with the previous run's results (downloads included), whose slots this run may reuse.
        e->launches += launch_peer_barrier(e->peer_flags, e->my_flags, e->my_flags + 8, (int)rank, (int)world, ++e->epoch, e->stream);
        CK(cudaGetLastError());
    }
    for (uint32_t l = first_level; l < last_level; l++) {
        const uint32_t a = p->level_off[l], n_all = p->level_off[l + 1] - a, n_pbs = p->level_pbs[l];
        if (world == 1) {
            rc = run_level(e, p->d_jobs + a, n_pbs, n_all);
            if (rc) return rc;
            continue;
        }
        // shard the PBS jobs of the level across ranks (equal contiguous slices, the last ones may be
        // short or empty), all-gather the result blocks in place, then run the cheap leveled jobs
        // replicated on every rank so that no exchange is needed for them
        uint32_t lo, hi, per;
        fhestr_shard_range(n_pbs, rank, world, &lo, &hi, &per);
        if (n_pbs) {
            const uint32_t first = p->level_first_dst[l];
            if (!p->level_contiguous[l] || (uint64_t)first + (uint64_t)per * world > e->arena_blocks)
                return fail(e, FHESTR_E_INVALID, "level cannot be sharded: PBS results must be contiguous and padded to a multiple of world");
            rc = run_level(e, p->d_jobs + a + lo, hi - lo, hi - lo, e->peers_attached);
            if (rc) return rc;
            if (e->peers_attached) {   // results are already in every arena: close the level with the flag barrier
                e->launches += launch_peer_barrier(e->peer_flags, e->my_flags, e->my_flags + 8, (int)rank, (int)world, ++e->epoch, e->stream);
                CK(cudaGetLastError());
                if (n_all > n_pbs) {
                    e->launches += launch_linear(p->d_jobs + a + n_pbs, (int)(n_all - n_pbs), e->arena, e->stream);
                    CK(cudaGetLastError());
                }
                continue;
            }
            u64* base = e->arena + (size_t)first * (kN + 1);
            const size_t cnt = (size_t)per * (kN + 1);
            const int nr = nccl().AllGather(base + (size_t)rank * cnt, base, cnt, kNcclUint64, e->comm, e->stream);
            if (nr != 0) return fail(e, FHESTR_E_CUDA, std::string("ncclAllGather: ") + (nccl().GetErrorString ? nccl().GetErrorString(nr) : "error"));
        }
        if (n_all > n_pbs) {
            e->launches += launch_linear(p->d_jobs + a + n_pbs, (int)(n_all - n_pbs), e->arena, e->stream);
            CK(cudaGetLastError());
        }
    }
    return FHESTR_OK;
}

void fhestr_program_destroy(fhestr_program* p) {
    if (!p) return;
    cudaFree(p->d_jobs);
    delete p;
}

void fhestr_shard_range(uint32_t n_jobs, uint32_t rank, uint32_t world, uint32_t* lo, uint32_t* hi, uint32_t* per) {
    const uint32_t w = world ? world : 1;
    const uint32_t p = (n_jobs + w - 1) / w;
    const uint32_t l = (uint64_t)p * rank < n_jobs ? p * rank : n_jobs;
    const uint32_t h = l + p < n_jobs ? l + p : n_jobs;
    if (lo) *lo = l;
    if (hi) *hi = h;
    if (per) *per = p;
}

int fhestr_peer_export(fhestr_engine* e, void* arena_handle_64, void* flags_handle_64) {
    if (!e || !arena_handle_64 || !flags_handle_64) return FHESTR_E_INVALID;
    if (!e->own_arena) return fail(e, FHESTR_E_STATE, "peer export needs an engine-owned arena");
    CK(cudaSetDevice(e->device));
    if (!e->my_flags) {
        CK(cudaMalloc(&e->my_flags, 16 * sizeof(uint32_t)));
        CK(cudaMemsetAsync(e->my_flags, 0, 16 * sizeof(uint32_t), e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(arena_handle_64), e->arena));
    CK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(flags_handle_64), e->my_flags));
    return FHESTR_OK;
}

int fhestr_peer_attach(fhestr_engine* e, uint32_t rank, uint32_t world, const void* arena_handles, const void* flags_handles) {
    if (!e || !arena_handles || !flags_handles || world < 2 || world > 8 || rank >= world) return FHESTR_E_INVALID;
    if (!e->my_flags) return fail(e, FHESTR_E_STATE, "call fhestr_peer_export first");
    if (e->peers_attached) return fail(e, FHESTR_E_STATE, "peers already attached");
    CK(cudaSetDevice(e->device));
    const cudaIpcMemHandle_t* ah = static_cast<const cudaIpcMemHandle_t*>(arena_handles);
    const cudaIpcMemHandle_t* fh = static_cast<const cudaIpcMemHandle_t*>(flags_handles);
    for (uint32_t r = 0; r < world; r++) {
        if (r == rank) { e->peer_arena[r] = e->arena; e->peer_flags[r] = e->my_flags; continue; }
        void* pa = nullptr; void* pf = nullptr;
        CK(cudaIpcOpenMemHandle(&pa, ah[r], cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(&pf, fh[r], cudaIpcMemLazyEnablePeerAccess));
        e->peer_arena[r] = static_cast<u64*>(pa);
        e->peer_flags[r] = static_cast<uint32_t*>(pf);
    }
    // epochs restart at 0: stale flags of an earlier attachment must not satisfy a new barrier.  Every rank zeroes its
    // own array here; the caller synchronises the ranks (any host barrier) between attach and the first run.
    CK(cudaMemsetAsync(e->my_flags, 0, 16 * sizeof(uint32_t), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->rank = rank; e->world = world; e->peers_attached = true; e->epoch = 0;
    return FHESTR_OK;
}

int fhestr_peer_detach(fhestr_engine* e) {
    if (!e) return FHESTR_E_INVALID;
    if (e->peers_attached) {
        cudaStreamSynchronize(e->stream);
        for (uint32_t r = 0; r < e->world; r++) {
            if (r == e->rank) continue;
            if (e->peer_arena[r]) cudaIpcCloseMemHandle(e->peer_arena[r]);
            if (e->peer_flags[r]) cudaIpcCloseMemHandle(e->peer_flags[r]);
            e->peer_arena[r] = nullptr; e->peer_flags[r] = nullptr;
        }
        e->peers_attached = false;
        if (!e->comm) { e->world = 1; e->rank = 0; }
    }
    return FHESTR_OK;
}

int fhestr_peer_status(fhestr_engine* e, uint32_t* timed_out) {
    if (!e || !timed_out) return FHESTR_E_INVALID;
    *timed_out = 0;
    if (!e->my_flags) return FHESTR_OK;
    CK(cudaMemcpyAsync(timed_out, e->my_flags + 8, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return FHESTR_OK;
}

int fhestr_comm_unique_id(void* unique_id_128) {
    if (!unique_id_128) return FHESTR_E_INVALID;
    if (!nccl().ok) return FHESTR_E_STATE;
    return nccl().GetUniqueId(static_cast<NcclUid*>(unique_id_128)) == 0 ? FHESTR_OK : FHESTR_E_CUDA;
}

int fhestr_comm_init(fhestr_engine* e, uint32_t rank, uint32_t world, const void* unique_id_128) {
    if (!e || !unique_id_128 || world == 0 || rank >= world) return FHESTR_E_INVALID;
    if (!nccl().ok) return fail(e, FHESTR_E_STATE, "libnccl.so.2 not found (dlopen)");
    if (e->comm) return fail(e, FHESTR_E_STATE, "communicator already initialised");
    CK(cudaSetDevice(e->device));
    NcclUid uid;
    memcpy(&uid, unique_id_128, sizeof uid);
    void* comm = nullptr;
    const int nr = nccl().CommInitRank(&comm, (int)world, uid, (int)rank);
    if (nr != 0) return fail(e, FHESTR_E_CUDA, std::string("ncclCommInitRank: ") + (nccl().GetErrorString ? nccl().GetErrorString(nr) : "error"));
    e->comm = comm;
    e->rank = rank;
    e->world = world;
    return FHESTR_OK;
}

int fhestr_comm_destroy(fhestr_engine* e) {
    if (!e) return FHESTR_E_INVALID;
    if (e->comm) {
        cudaStreamSynchronize(e->stream);
        nccl().CommDestroy(e->comm);
        e->comm = nullptr;
        e->world = 1;
        e->rank = 0;
    }
    return FHESTR_OK;
}

int fhestr_debug_keyswitch(fhestr_engine* e, const fhestr_job* jobs, uint32_t n_jobs, uint64_t* host_out) {
    if (!e || !jobs || !host_out) return FHESTR_E_INVALID;
    if (!e->keys_loaded) return fail(e, FHESTR_E_STATE, "keys not loaded");
    int rc = validate_jobs(e, jobs, n_jobs);
    if (rc) return rc;
    CK(cudaSetDevice(e->device));
    rc = ensure_scratch(e, n_jobs);
    if (rc) return rc;
    CK(cudaMemcpyAsync(e->d_jobs, jobs, (size_t)n_jobs * sizeof(fhestr_job), cudaMemcpyHostToDevice, e->stream));
    KsBatchArgs ks{e->d_jobs, e->arena, e->ksk, e->ksk_corr, e->ks_out, e->prm.n, (int)n_jobs,
                   e->prm.ks_base_log, e->prm.ks_level, e->ksk8, e->ks_digits, e->ks_body};
    const int ks_launches = e->ks_path == 0 ? launch_keyswitch_tc(ks, e->stream) : launch_keyswitch(ks, e->stream);
    if (ks_launches == 0 && n_jobs > 0) return fail(e, FHESTR_E_CUDA, "keyswitch: the TMA tensor maps could not be encoded");
    e->launches += ks_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_out, e->ks_out, (size_t)n_jobs * (e->prm.n + 1) * sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return FHESTR_OK;
}

int fhestr_debug_blind_rotate(fhestr_engine* e, const uint64_t* ks_host, const int32_t* lut_ids,
                              const uint64_t* init_acc_host, uint32_t count, uint64_t* acc_out_host) {
    if (!e || !ks_host || !acc_out_host || (!lut_ids && !init_acc_host)) return FHESTR_E_INVALID;
    if (!e->keys_loaded) return fail(e, FHESTR_E_STATE, "keys not loaded");
    CK(cudaSetDevice(e->device));
    int rc = ensure_scratch(e, count);
    if (rc) return rc;
    const size_t acc_bytes = (size_t)count * 2 * kN * sizeof(u64);
    u64 *d_init = nullptr, *d_out = nullptr;
    int32_t* d_ids = nullptr;
    CK(cudaMalloc(&d_out, acc_bytes));
    CK(cudaMalloc(&d_ids, count * sizeof(int32_t)));
    std::vector<int32_t> ids(count, 0);
    if (lut_ids) ids.assign(lut_ids, lut_ids + count);
    for (auto v : ids) if (v < 0 || (v >= e->n_luts && !init_acc_host)) { cudaFree(d_out); cudaFree(d_ids); return fail(e, FHESTR_E_INVALID, "bad lut id"); }
    CK(cudaMemcpyAsync(d_ids, ids.data(), count * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (init_acc_host) {
        CK(cudaMalloc(&d_init, acc_bytes));
        CK(cudaMemcpyAsync(d_init, init_acc_host, acc_bytes, cudaMemcpyHostToDevice, e->stream));
    }
    CK(cudaMemcpyAsync(e->ks_out, ks_host, (size_t)count * (e->prm.n + 1) * sizeof(u64), cudaMemcpyHostToDevice, e->stream));
    BrBatchArgs br{};
    br.ks = e->ks_out; br.luts = e->luts; br.lut_ids = d_ids; br.jobs = nullptr; br.arena = e->arena;
    br.bsk = e->bsk_f; br.tf = e->tf; br.init_acc = d_init; br.out_acc = d_out;
    br.n = e->prm.n; br.B = (int)count;
    br.bsk_w = e->bsk_w; br.wide_tab = e->wide_tab;
    e->launches += launch_level_br(e, br, count);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(acc_out_host, d_out, acc_bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_init); cudaFree(d_out); cudaFree(d_ids);
    return FHESTR_OK;
}

int fhestr_measure_fp64_peak(fhestr_engine* e, double* tflops, double* sm_clock_mhz_hint) {
    if (!e || !tflops) return FHESTR_E_INVALID;
    CK(cudaSetDevice(e->device));
    double* sink = nullptr;
    CK(cudaMalloc(&sink, 148 * 8 * 256 * sizeof(double)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    unsigned long long fmas = 0;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a, e->stream));
        e->launches += launch_dfma_peak(sink, 4096, &fmas, e->stream);
        CK(cudaEventRecord(b, e->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double tf = 2.0 * (double)fmas / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    if (sm_clock_mhz_hint) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
        *sm_clock_mhz_hint = khz / 1000.0;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    return FHESTR_OK;
}

uint64_t fhestr_kernel_launches(const fhestr_engine* e) { return e ? e->launches : 0; }

static void drop_timed(fhestr_engine* e) {
    for (auto& t : e->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); cudaEventDestroy(t.c); }
    e->timed.clear();
}

int fhestr_set_timing(fhestr_engine* e, int enable) {
    if (!e) return FHESTR_E_INVALID;
    CK(cudaStreamSynchronize(e->stream));
    drop_timed(e);
    e->timing = enable != 0;
    return FHESTR_OK;
}

int fhestr_get_timing(fhestr_engine* e, double* ks_ms, double* br_ms, uint64_t* br_launches, uint64_t* br_pbs) {
    if (!e) return FHESTR_E_INVALID;
    CK(cudaStreamSynchronize(e->stream));
    double ks = 0, br = 0;
    uint64_t pbs = 0;
    for (auto& t : e->timed) {
        float m1 = 0, m2 = 0;
        CK(cudaEventElapsedTime(&m1, t.a, t.b));
        CK(cudaEventElapsedTime(&m2, t.b, t.c));
        ks += m1; br += m2; pbs += t.pbs;
    }
    if (ks_ms) *ks_ms = ks;
    if (br_ms) *br_ms = br;
    if (br_launches) *br_launches = e->timed.size();
    if (br_pbs) *br_pbs = pbs;
    drop_timed(e);
    return FHESTR_OK;
}

int fhestr_set_keyswitch_path(fhestr_engine* e, int path) {
    if (!e || (path != 0 && path != 1)) return FHESTR_E_INVALID;
    if (path == 0 && e->prm.ks_base_log > 7) return fail(e, FHESTR_E_INVALID, "tensor-core keyswitch needs ks_base_log <= 7");
    e->ks_path = path;
    return FHESTR_OK;
}

int fhestr_set_br_mode(fhestr_engine* e, int mode, int wide_max_jobs) {
    if (!e || mode < 0 || mode > 4 || wide_max_jobs < 0) return FHESTR_E_INVALID;
    e->br_mode = mode > 2 ? 2 : mode;
    e->wide_shape = mode > 2 ? mode - 2 : 0;     // 3 / 4: the latency kernel forced to its single / pair form
    e->wide_max_jobs = wide_max_jobs;
    return FHESTR_OK;
}

}  // extern "C"
#pragma GCC visibility pop
