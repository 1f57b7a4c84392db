// oon_capi.cu — the C ABI of include/oon_gpu.h: world handle, device memory, step orchestration,
// NCCL exchange.  Host-side only; the kernels are in oon_kernels.cu.
//
// Reference interface replaced: OON::Model::World (/root/reference/src/app/model/World.hpp:82-158) and the
// model part of OONApp::updates_for_next_frame (/root/reference/src/app/OON.cpp:1083-1095).
#include "oon_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include <algorithm>
#include <functional>
#include <utility>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace oon;

static_assert(sizeof(oon_body_f32) == 60, "oon_body_f32 must match the reference's float-build Entity");
static_assert(sizeof(oon_body_f64) == 96, "oon_body_f64 must match the reference's double-build Entity");
static_assert(sizeof(oon_props) == 40 && sizeof(oon_events) == 64, "ABI struct sizes");

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: a single-GPU world never needs the library.
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) {
            error = "libnccl.so.2 lacks a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
std::string g_create_error;
}  // namespace

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct SoAStorage {
    BodySoA d{};
    uint32_t cap = 0;
};

struct oon_world {
    // A handle is one of three things:
    //   * one world on one GPU (nranks == 1),
    //   * one shard of a world sharded over nranks PROCESSES (oon_create_sharded: comm != nullptr, collectives through NCCL),
    //   * a GROUP (oon_create_multi): one process, one host thread, kids.size() GPUs.  The group handle itself owns no
    //     device state; its children are ordinary shards (rank g of kids.size()) whose collectives are done in-process:
    //     peer-access pointers instead of CUDA IPC, cross-stream events + peer copies instead of NCCL.
    // Every entry point works on the TEAM of shards the calling thread drives: {this} or kids.
    std::vector<oon_world*> kids;      // group handle only
    oon_world* parent = nullptr;       // child of a group handle
    std::vector<oon_world*> team;      // {this} for a plain handle, kids for a group handle
    cudaEvent_t ev_grp = nullptr, ev_grp_done = nullptr;   // in-process collectives of a group (record / consumed)
    int device = 0;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_a = nullptr, ev_b = nullptr;

    oon_props props{};
    int math_mode = OON_MATH_FAST;

    // partition of the world: rank g owns the contiguous index block [g*block, min(n_total,(g+1)*block))
    uint64_t n_total = 0;     // bodies in the whole world (host view; refreshed from the device after sweeps)
    uint32_t block = 0;       // slots per rank (multiple of TS); slots == block
    uint32_t cblock = 0;      // bodies per canonical block (block == cblock * 8 / nranks)
    uint32_t n_local = 0;     // host view of the live local count
    bool counts_stale = false; // a sweep ran on the device since the host last learnt the counts
    // What the host knows about expiry (same on every shard): while `life_known`, no body is expired and none can expire
    // before `life_budget` more seconds of dt have passed — such frames run the immortal schedule (no sweep, fused launches,
    // Hilbert order).  Learnt from the device at uploads and whenever the counts are fetched after a sweep; kept by edits.
    bool life_known = false, pending_expired = false;
    double life_budget = 0.0;
    bool last_frame_swept = false; // the most recent frame ended with a sweep (oon_debug_fetch_removed: else nothing was removed)
    uint32_t* d_minfo = nullptr;   // [4] lifetime summary written by the unpack / sweep kernels (see note_lifetime)
    uint32_t* d_work_ctr = nullptr; // [2] work counter of the persistent FAST interact kernel (zero between launches)
    bool persistent_interact = false;   // OON_K1_PERSISTENT=1: ~592 persistent CTAs draw the work items dynamically (measured slower, see profiles/README.md)
    bool has_mortals = false; // any body with lifetime > 0 or a non-unlimited non-positive lifetime
    uint32_t player_ndx = 0;  // player_entity_ndx(): the sweep inside oon_step tests the bodies after it (OON.cpp:1089)

    SoAStorage soa[2];        // [cur] is live; the other one is the sweep target
    int cur = 0;
    float* tiles = nullptr;   // exchange buffer in use: nranks*block/TS tiles of TILE_FLOATS floats (x,y,m,r,id arrays + header)
    float* tiles_buf[2] = {nullptr, nullptr};            // [1] only with the peer-to-peer exchange (epoch parity picks the buffer)
    float4* partial = nullptr; size_t partial_cap = 0;   // [block]: result of the EXACT / player-only kernels
    unsigned long long* acc = nullptr;                   // [2][FX_LIMBS][block] exact kick sums of the FAST all-pairs kernel
    uint32_t* touch = nullptr;                           // [block] touch visits of the FAST all-pairs kernel
    uint32_t acc_cap = 0; bool acc_dirty = false;
    int* fx_scale = nullptr;                             // exponent of the fixed-point sums of the last pass
    uint32_t* rank_stats = nullptr;                      // pack kernels: {max |coord|, min |mass|, ticket}

    // peer-to-peer exchange (sharded worlds on one NVLink node): the pack kernel stores this rank's slice straight into
    // every peer's exchange buffer and raises a flag there; the interact kernel waits for the flag of a slice before
    // the TMA of its first tile.  No collective call, no extra launch on the critical path.
    bool p2p_wanted = true, p2p = false;
    uint32_t epoch = 0;                                  // packs so far; slice g of buffer epoch&1 is valid once flags[g] >= epoch
    uint32_t waited_epoch = 0;                           // last epoch a kernel of this stream has waited for
    uint32_t* flags = nullptr;                           // [nranks], written by the peers
    uint32_t* timeouts = nullptr;                        // waits that gave up (device alias of timeouts_host)
    uint32_t* timeouts_host = nullptr;                   // mapped pinned word: read after every synchronisation (sync_main)
    void* ipc_stage = nullptr;                           // device staging for the handle exchange
    float* peer_tiles[2][8] = {};                        // peers' exchange buffers (imported), indexed by rank
    uint32_t* peer_flags[8] = {};                        // peers' flag arrays (imported)
    uint32_t group_tiles_override = 0, debug_gpc = 0, debug_spc = 0, tail_waves_x2 = 4;
    int small_tiles_override = -1;
    uint32_t dbg_fraction = 0, dbg_part = 0;             // oon_debug_time_interact: launch a part of the target blocks
    unsigned long long* dbg_trace = nullptr; uint32_t dbg_trace_ctas = 0;   // oon_debug_trace_interact: per-CTA timeline of that launch
    float4* scratch4 = nullptr;                          // player-only reaction buffer, one float4 per slot
    size_t tiles_cap_slots = 0;                          // slots the exchange buffer (and scratch4) can hold
    float2* kick = nullptr; uint32_t kick_cap = 0; bool kick_valid = false;
    uint32_t* d_n = nullptr;  // live local bodies (device)
    uint32_t* d_removed_count = nullptr;
    int* runinfo = nullptr;                              // sharded sweep: [nranks][4] {bodies, trailing expired run, whole block one run, -}
    uint32_t* d_counts = nullptr;                        // sharded: live bodies of every rank (device), counts[] = host copy
    std::vector<uint32_t> counts;                        // sharded: bodies of every rank as the host last learnt them (authoritative unless counts_stale)
    EmitBase* emit_rec = nullptr;                        // [nranks] records of the body-creating loops (emitter / parent / result)
    unsigned long long* counters = nullptr;   // C_COUNT
    void* sweep_temp = nullptr; size_t sweep_temp_bytes = 0;
    int* sweep_scratch = nullptr; uint32_t sweep_scratch_cap = 0;
    uint32_t* removed_idx = nullptr; uint32_t removed_cap = 0;
    oon_body_f32* aos_dev = nullptr; size_t aos_cap = 0;   // staging for upload/download
    float* xyr_dev = nullptr; size_t xyr_cap = 0;
    cudaStream_t copy_stream = nullptr;                     // render read-back beside the next frames (oon_download_render_begin/_end)
    cudaEvent_t ev_render_ready = nullptr, ev_render_done = nullptr;
    bool render_pending = false; uint64_t render_count = 0;
    CaptureBuf cap{};

    // slot order of the exchange buffer: inv[slot] = local body stored there
    uint32_t* inv = nullptr; uint32_t inv_cap = 0;
    uint32_t* inv_next = nullptr;   // order being computed ahead on the aux stream (from the positions of the current frame)
    uint32_t* slot_of = nullptr;    // [body] slot it was packed into (inverse of inv as of the last pack)
    bool fuse_reorder = true;       // integrate(k) + pack(k+1) in one launch even when the slot order changes in between
    cudaStream_t aux = nullptr;     // sort-ahead stream: the O(N) ordering of frame k+1 hides behind the O(N^2) pass of frame k
    cudaEvent_t ev_main = nullptr, ev_sorted = nullptr;
    bool next_order_pending = false;
    bool async_sort = true;
    int order_kind = 0;             // 0 = none yet, 1 = identity, 2 = Hilbert curve (the code says "morton" for historical reasons)
    bool order_dirty = true;        // body table changed since the order was computed
    uint32_t order_age = 0;         // frames since the last curve sort
    uint32_t resort_every = 1;
    bool morton_enabled = true;
    bool order_one_cta = true;      // canonical blocks that fit one CTA are ordered by one launch (k_order_block)
    uint64_t morton_min_bodies = 32768;
    void* order_temp = nullptr; size_t order_temp_bytes = 0;
    unsigned long long* order_keys = nullptr; uint32_t* order_vals = nullptr; uint32_t* order_bounds = nullptr; uint32_t order_cap = 0;

    bool tiles_fresh = false;       // exchange buffer matches the SoA state
    uint64_t steps = 0;             // since last drain
    uint64_t launches = 0;
    int sm_count = 148;

    int sticky = OON_OK;
    std::string error;
};

namespace {

int fail(oon_world* w, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (w) {
        w->error = buf;
        if (code == OON_ERR_CUDA || code == OON_ERR_NCCL) w->sticky = code;
        if (w->parent) { w->parent->error = buf; if (w->sticky != OON_OK) w->parent->sticky = w->sticky; }   // the caller holds the group handle
    } else g_create_error = buf;
    return code;
}

#define CU(w, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(w, OON_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define NC(w, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(w, OON_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)
#define CHECK_W(w) do { if (!(w)) return OON_ERR_INVALID; if ((w)->sticky != OON_OK) return (w)->sticky; if ((w)->kids.empty() && cudaSetDevice((w)->device) != cudaSuccess) return fail(w, OON_ERR_CUDA, "cudaSetDevice(%d) failed", (w)->device); } while (0)
// make shard m's device current (a team of a group handle spans several devices; a plain handle's device was set by CHECK_W)
#define USE(m) do { if ((m)->parent && cudaSetDevice((m)->device) != cudaSuccess) return fail(m, OON_ERR_CUDA, "cudaSetDevice(%d) failed", (m)->device); } while (0)
inline bool in_group(const oon_world* m) { return m->parent != nullptr; }

// Wait for the main stream, then look at the peer-to-peer exchange's time-out word (mapped host memory the kernels bump
// when a wait for a peer's slice gives up): a kernel that carried on without a slice has produced garbage and has broken
// the epoch protocol, so the failure is sticky — every later call on the handle refuses to run.
int peers_ok(oon_world* w) {
    if (w->timeouts_host) {
        const uint32_t t = *static_cast<volatile uint32_t*>(w->timeouts_host);
        if (t) return fail(w, OON_ERR_NCCL, "peer-to-peer exchange: %u waits for a peer's slice timed out (a rank fell out of step); the world is no longer valid", t);
    }
    return OON_OK;
}
int sync_main(oon_world* w) {
    cudaError_t e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) return fail(w, OON_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
    return peers_ok(w);
}
#define SYNC(w) do { int s_ = sync_main(w); if (s_) return s_; } while (0)

template <class T>
int dev_alloc(oon_world* w, T** p, size_t count) {
    *p = nullptr;
    if (!count) return OON_OK;
    CU(w, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return OON_OK;
}

void soa_free(SoAStorage& s) {
    cudaFree(s.d.p); cudaFree(s.d.v); cudaFree(s.d.T); cudaFree(s.d.life); cudaFree(s.d.mass); cudaFree(s.d.r);
    cudaFree(s.d.density); cudaFree(s.d.color); cudaFree(s.d.flags); cudaFree(s.d.thrust);
    s = SoAStorage{};
}

int soa_alloc(oon_world* w, SoAStorage& s, uint32_t cap) {
    soa_free(s);
    int rc;
    if ((rc = dev_alloc(w, &s.d.p, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.v, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.T, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.life, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.mass, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.r, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.density, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.color, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.flags, cap))) return rc;
    if ((rc = dev_alloc(w, &s.d.thrust, cap))) return rc;
    s.cap = cap;
    return OON_OK;
}

#define COPY_FIELD(w, f, n) CU(w, cudaMemcpyAsync(dst.d.f, src.d.f, size_t(n) * sizeof(*dst.d.f), cudaMemcpyDeviceToDevice, (w)->stream))
int soa_copy(oon_world* w, SoAStorage& dst, const SoAStorage& src, uint32_t n) {
    if (!n) return OON_OK;
    COPY_FIELD(w, p, n); COPY_FIELD(w, v, n); COPY_FIELD(w, T, n); COPY_FIELD(w, life, n); COPY_FIELD(w, mass, n);
    COPY_FIELD(w, r, n); COPY_FIELD(w, density, n); COPY_FIELD(w, color, n); COPY_FIELD(w, flags, n); COPY_FIELD(w, thrust, n);
    return OON_OK;
}

uint32_t round_up(uint64_t x, uint32_t m) { return uint32_t((x + m - 1) / m * m); }

// ---- peer-to-peer exchange plumbing ------------------------------------------------------------------------------
typedef std::vector<oon_world*> Team;

// Host-level rendezvous of the ranks of a multi-process world: a one-word all-gather, then wait for it.  (The shards of a
// group are driven by one host thread: synchronising their streams IS the rendezvous.)
int barrier(oon_world* w) {
    if (w->nranks == 1 || !w->comm) return OON_OK;
    NC(w, g_nccl.AllGather(static_cast<char*>(w->ipc_stage) + size_t(w->rank) * 4, w->ipc_stage, 4, ncclChar, w->comm, w->stream));
    SYNC(w);
    return OON_OK;
}

void p2p_close_tiles(oon_world* w) {
    for (int b = 0; b < 2; ++b)
        for (int g = 0; g < 8; ++g)
            if (w->peer_tiles[b][g]) { if (!in_group(w)) cudaIpcCloseMemHandle(w->peer_tiles[b][g]); w->peer_tiles[b][g] = nullptr; }
}

struct IpcMsg {
    cudaIpcMemHandle_t tiles[2];
    cudaIpcMemHandle_t flags;
    int ok;
    int pad[15];
};
static_assert(sizeof(IpcMsg) == 256, "IpcMsg is one 256-byte row of the staging buffer");

// Multi-process world: export this rank's two exchange buffers (and, once, its flag array), import everybody else's.  Every
// rank reaches the same verdict: if any export or import fails anywhere (no peer access, different nodes, two ranks in one
// process...) all ranks fall back to the NCCL all-gather.
int p2p_open(oon_world* w) {
    w->p2p = false;
    if (w->nranks < 2 || w->nranks > 8 || !w->p2p_wanted) return OON_OK;
    IpcMsg mine{};
    mine.ok = 1;
    if (cudaIpcGetMemHandle(&mine.tiles[0], w->tiles_buf[0]) != cudaSuccess || cudaIpcGetMemHandle(&mine.tiles[1], w->tiles_buf[1]) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.flags, w->flags) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;
    }
    std::vector<IpcMsg> all(size_t(w->nranks));
    auto gather = [&]() -> int {
        IpcMsg* stage = static_cast<IpcMsg*>(w->ipc_stage);
        CU(w, cudaMemcpyAsync(stage + w->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, w->stream));
        NC(w, g_nccl.AllGather(stage + w->rank, stage, sizeof(IpcMsg), ncclChar, w->comm, w->stream));
        CU(w, cudaMemcpyAsync(all.data(), stage, sizeof(IpcMsg) * size_t(w->nranks), cudaMemcpyDeviceToHost, w->stream));
        SYNC(w);
        return OON_OK;
    };
    int rc = gather();
    if (rc) return rc;
    bool ok = true;
    for (const IpcMsg& m : all) ok = ok && m.ok;
    if (ok) {
        for (int g = 0; g < w->nranks && mine.ok; ++g) {
            if (g == w->rank) continue;
            for (int b = 0; b < 2 && mine.ok; ++b) {
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[size_t(g)].tiles[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
                else w->peer_tiles[b][g] = static_cast<float*>(ptr);
            }
            if (mine.ok && !w->peer_flags[g]) {
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[size_t(g)].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
                else w->peer_flags[g] = static_cast<uint32_t*>(ptr);
            }
        }
    }
    if ((rc = gather())) return rc;                        // second round: did every import work everywhere?
    for (const IpcMsg& m : all) ok = ok && m.ok;
    if (!ok) {
        p2p_close_tiles(w);
        if ((rc = barrier(w))) return rc;                  // nobody maps a buffer that is about to be freed
        cudaFree(w->tiles_buf[1]); w->tiles_buf[1] = nullptr;
        return OON_OK;
    }
    w->p2p = true;
    return OON_OK;
}

// (Re)allocate the exchange buffers of every shard of the team for `need_slots` slots and connect the peers again.
// Collective on multi-process worlds (every rank lays the same world out at the same time).
int team_realloc_exchange(Team& T, size_t need_slots) {
    int rc;
    for (oon_world* w : T) { USE(w); SYNC(w); }            // no kernel of this team still stores into a buffer about to go
    for (oon_world* w : T) {
        if (!w->p2p) continue;
        p2p_close_tiles(w);
        if (!in_group(w) && (rc = barrier(w))) return rc;  // multi-process: every peer has unmapped my buffers before I free them
        w->p2p = false;
    }
    for (oon_world* w : T) {
        USE(w);
        cudaFree(w->tiles_buf[0]); cudaFree(w->tiles_buf[1]); cudaFree(w->scratch4);
        w->tiles_buf[0] = w->tiles_buf[1] = w->tiles = nullptr; w->scratch4 = nullptr; w->tiles_cap_slots = 0;
        const bool two = in_group(w) || (w->nranks >= 2 && w->nranks <= 8 && w->p2p_wanted);
        if ((rc = dev_alloc(w, &w->tiles_buf[0], need_slots / TS * TILE_FLOATS))) return rc;
        if (two && (rc = dev_alloc(w, &w->tiles_buf[1], need_slots / TS * TILE_FLOATS))) return rc;
        if ((rc = dev_alloc(w, &w->scratch4, need_slots))) return rc;
        w->tiles = w->tiles_buf[0];
        w->tiles_cap_slots = need_slots;
    }
    for (oon_world* w : T) {
        if (in_group(w)) {                                 // group: the peers' buffers are plain peer-access pointers
            for (oon_world* g : T) {
                if (g == w) continue;
                w->peer_tiles[0][g->rank] = g->tiles_buf[0]; w->peer_tiles[1][g->rank] = g->tiles_buf[1];
                w->peer_flags[g->rank] = g->flags;
            }
            w->p2p = true;
        } else if (w->nranks >= 2 && w->nranks <= 8 && w->p2p_wanted) {
            USE(w);
            if ((rc = p2p_open(w))) return rc;
        }
        if (w->p2p) w->tiles = w->tiles_buf[w->epoch & 1u];
    }
    return OON_OK;
}

// Lay one shard out for a world with room for `n_cap` bodies of which `n_present` exist now (spread over the ranks by
// block boundaries): block size, SoA capacity, accumulators, order buffers.  Keeps the first `keep` local bodies of the
// current SoA.  *need_slots_out = slots the exchange buffer must be re-allocated for (0: it is large enough).
int layout_one(oon_world* w, uint64_t n_cap, uint64_t n_present, uint32_t keep, size_t* need_slots_out) {
    // Canonical blocks: 1/8 of the world each, whatever the number of GPUs; a rank owns 8/nranks of them.  (Other
    // rank counts fall back to one block per rank: still correct, but their sums differ in the last bits from 1/2/4/8.)
    const bool canon = (w->nranks == 1 || w->nranks == 2 || w->nranks == 4 || w->nranks == 8);
    const uint32_t nblocks = canon ? 8u : uint32_t(w->nranks);
    const uint32_t cblock = std::max<uint32_t>(TS, round_up((n_cap + nblocks - 1) / nblocks, TS));
    const uint32_t block = cblock * (nblocks / uint32_t(w->nranks));
    w->cblock = cblock;
    if (uint64_t(block) * w->nranks > 0x40000000ull)   // 2^22 tiles: the carry-save limbs take that many addends
        return fail(w, OON_ERR_INVALID, "world too large: %llu bodies", (unsigned long long)n_cap);
    if (block > w->soa[w->cur].cap) {
        // a single-GPU world gets head-room so that add_bodies rarely reallocates
        const uint32_t cap = w->nranks == 1 ? round_up(uint64_t(block) + block / 2, TS) : block;
        SoAStorage fresh;
        int rc = soa_alloc(w, fresh, cap);
        if (rc) return rc;
        if (keep) { rc = soa_copy(w, fresh, w->soa[w->cur], keep); if (rc) return rc; SYNC(w); }
        soa_free(w->soa[w->cur]);
        soa_free(w->soa[w->cur ^ 1]);
        w->soa[w->cur] = fresh;
    }
    // exchange buffer (+ the player-only reaction scratch, one float4 per slot)
    const size_t need_slots = size_t(w->nranks) * (w->nranks == 1 ? w->soa[w->cur].cap : block);
    *need_slots_out = need_slots > w->tiles_cap_slots ? need_slots : 0;
    // exact accumulators of the FAST all-pairs kernel, zero between passes
    if (w->soa[w->cur].cap > w->acc_cap) {
        cudaFree(w->acc); cudaFree(w->touch); w->acc = nullptr; w->touch = nullptr; w->acc_cap = 0;
        const uint32_t cap = w->soa[w->cur].cap;
        int rc = dev_alloc(w, &w->acc, size_t(2 * FX_LIMBS) * cap);
        if (rc) return rc;
        if ((rc = dev_alloc(w, &w->touch, cap))) return rc;
        w->acc_cap = cap;
        w->acc_dirty = true;
    }
    if (w->soa[w->cur].cap > w->inv_cap) {
        if (w->next_order_pending) { CU(w, cudaStreamSynchronize(w->aux)); w->next_order_pending = false; }
        cudaFree(w->inv); cudaFree(w->inv_next); cudaFree(w->slot_of); w->inv = nullptr; w->inv_next = nullptr; w->slot_of = nullptr; w->inv_cap = 0;
        int rc = dev_alloc(w, &w->inv, w->soa[w->cur].cap);
        if (rc) return rc;
        if ((rc = dev_alloc(w, &w->inv_next, w->soa[w->cur].cap))) return rc;
        if ((rc = dev_alloc(w, &w->slot_of, w->soa[w->cur].cap))) return rc;
        w->inv_cap = w->soa[w->cur].cap;
    }
    if (block != w->block || n_present != w->n_total) {   // same table shape: a stale order is still a valid (if less tight) order
        w->order_dirty = true;
        w->kick_valid = false;
    }
    w->block = block;
    w->n_total = n_present;
    const uint64_t lo = uint64_t(w->rank) * block;
    w->n_local = n_present > lo ? uint32_t(std::min<uint64_t>(block, n_present - lo)) : 0u;
    for (int g = 0; g < w->nranks && !w->counts.empty(); ++g) {          // the even split, until sweeps and edits make the blocks ragged
        const uint64_t glo = uint64_t(g) * block;
        w->counts[size_t(g)] = n_present > glo ? uint32_t(std::min<uint64_t>(block, n_present - glo)) : 0u;
    }
    w->counts_stale = false;
    w->tiles_fresh = false;
    return OON_OK;
}

// keep_local: keep every shard's current local bodies (growing a world in place) instead of starting from an empty table
int team_layout(Team& T, uint64_t n_cap, uint64_t n_present, bool keep_local) {
    size_t need = 0;
    for (oon_world* w : T) {
        USE(w);
        size_t nm = 0;
        const uint32_t keep = keep_local ? w->n_local : 0u;
        int rc = layout_one(w, n_cap, n_present, keep, &nm);
        if (rc) return rc;
        if (keep_local) w->n_local = keep;
        need = std::max(need, nm);
    }
    return need ? team_realloc_exchange(T, need) : OON_OK;
}

uint32_t total_tiles(const oon_world* w) { return uint32_t(uint64_t(w->block) * w->nranks / TS); }

// Canonical groups of source tiles: FP32 sums inside a group, exact sums across groups.  Group sizes are a function of
// the world size only (they fix the rounding of the result): 4 tiles (2 below 128 tiles, 1 below 32), and from 128 tiles
// on the last 32-35 tiles of the source range are groups of their own.  How the groups are dealt to CTAs is launch
// geometry and free to choose: the bulk goes to CTAs that walk `gpc` groups (16 tiles), then about tail_waves waves of
// one-group CTAs, then the single tiles — so that the SMs run dry together (a 1/8 shard of the 100k-body world is only
// ~8 waves of uniform CTAs: 8.1 waves cost 9, and the drain after the last wave costs half a CTA's duration).
InteractGeometry grouping(const oon_world* w) {
    InteractGeometry geo{};
    const uint32_t ntiles = total_tiles(w);
    geo.group_tiles = ntiles >= 128 ? 4u : ntiles >= 32 ? 2u : 1u;
    uint32_t n_small = ntiles >= 128 ? 32u : 0u;
    if (w->group_tiles_override) geo.group_tiles = w->group_tiles_override;      // experiments only: changes the rounding
    if (w->small_tiles_override >= 0) n_small = std::min<uint32_t>(uint32_t(w->small_tiles_override), ntiles);   // likewise
    geo.small_begin = n_small ? (ntiles - n_small) / geo.group_tiles * geo.group_tiles : ntiles;
    const uint64_t nbig = (uint64_t(geo.small_begin) + geo.group_tiles - 1) / geo.group_tiles;   // groups before small_begin
    uint64_t target_blocks = (uint64_t(w->block) + K1_TARGETS_PER_CTA - 1) / K1_TARGETS_PER_CTA;
    if (w->dbg_fraction > 1) target_blocks = (target_blocks + w->dbg_fraction - 1) / w->dbg_fraction;
    const uint64_t slots = uint64_t(w->sm_count) * 4;                         // resident CTAs
    // long CTAs: ~24 tiles (r2 sweep with the final kernel, 100 k bodies: 8 tiles 2.791 ms, 12: 2.774, 16: 2.765, 24: 2.756, 32: 2.757 per
    // launch; a 1/8 shard: 0.395 / 0.393 / 0.393 / 0.392 / 0.424), more for very large worlds (keep the grid below ~64 waves)
    uint64_t g = std::max<uint64_t>(1, 24 / geo.group_tiles);
    g = std::max<uint64_t>(g, (target_blocks * nbig + slots * 64 - 1) / (slots * 64));
    g = std::max<uint64_t>(g, (nbig + 50000) / 50001);
    if (w->debug_gpc) g = w->debug_gpc;                                        // tuning experiments only
    g = std::max<uint64_t>(1, std::min<uint64_t>(g, std::max<uint64_t>(1, nbig)));
    // one-group CTAs: about tail_waves_x2 / 2 waves of them
    uint64_t ymid = g > 1 ? (slots * w->tail_waves_x2 / 2 + target_blocks - 1) / target_blocks : 0;
    ymid = std::min<uint64_t>(ymid, std::min<uint64_t>(nbig, 4000));
    const uint64_t yb = (nbig - ymid) / g;
    geo.gpc = uint32_t(g);
    geo.y_big = uint32_t(yb);
    geo.y_mid = uint32_t(nbig - yb * g);
    // single tiles: at least ~2.5 waves of CTAs, up to 4 tiles per CTA when there are many target blocks
    const uint64_t n_single = ntiles - geo.small_begin;
    uint64_t spc = n_single ? target_blocks * n_single * 2 / (slots * 5) : 1;
    geo.spc = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(spc, 4)));
    if (w->debug_spc) geo.spc = w->debug_spc;
    return geo;
}

int ensure_partial(oon_world* w) {
    const size_t need = w->block;
    if (need > w->partial_cap) {
        cudaFree(w->partial); w->partial = nullptr; w->partial_cap = 0;
        int rc = dev_alloc(w, &w->partial, need);
        if (rc) return rc;
        w->partial_cap = need;
    }
    return OON_OK;
}

int ensure_aos(oon_world* w, size_t n) {
    if (n > w->aos_cap) {
        cudaFree(w->aos_dev); w->aos_dev = nullptr; w->aos_cap = 0;
        const size_t cap = n + n / 2 + 64;
        int rc = dev_alloc(w, &w->aos_dev, cap);
        if (rc) return rc;
        w->aos_cap = cap;
    }
    return OON_OK;
}

StepParams step_params(const oon_world* w, float dt) {
    StepParams sp{};
    sp.dt = dt;
    sp.friction = w->props.friction;
    sp.gravity = float(w->props.gravity);
    sp.repulsion = w->props.repulsion_stiffness;
    sp.gravity_mode = w->props.gravity_mode;
    sp.collision_on = w->props.collision_mode != OON_COLLISION_OFF;
    sp.n2n = w->props.interact_n2n != 0;
    sp.full_matrix = w->props.loop_mode == OON_LOOP_FULL_MATRIX;
    sp.exact = w->math_mode == OON_MATH_EXACT;
    return sp;
}

float* slice_of(oon_world* w, float* buffer, int rank) { return buffer + size_t(rank) * (w->block / TS) * TILE_FLOATS; }

ExchangeView xview(const oon_world* w) {
    ExchangeView x{};
    x.nranks = uint32_t(w->nranks); x.tiles_per_rank = w->block / TS; x.rank = uint32_t(w->rank);
    x.epoch = w->p2p ? w->epoch : 0u; x.flags = w->flags; x.timeouts = w->timeouts;
    return x;
}

// Open the next pack: with the peer-to-peer exchange it goes to the other buffer under the next epoch, and the buffer it
// overwrites (epoch - 2) is free on every peer because this stream has waited for every peer's flag of epoch - 1.
int begin_pack(oon_world* w, PackOut& out) {
    out = PackOut{};
    out.rank_stats = w->rank_stats;
    if (w->p2p) {
        if (w->waited_epoch != w->epoch) {
            w->launches += uint64_t(launch_wait_peers(w->stream, xview(w)));
            CU(w, cudaGetLastError());
            w->waited_epoch = w->epoch;
        }
        ++w->epoch;
        w->tiles = w->tiles_buf[w->epoch & 1u];
        out.epoch = w->epoch;
        for (int g = 0; g < w->nranks; ++g) {
            if (g == w->rank) continue;
            out.peer_tiles[out.npeers] = slice_of(w, w->peer_tiles[w->epoch & 1u][g], w->rank);
            out.peer_flags[out.npeers] = w->peer_flags[g] + w->rank;
            ++out.npeers;
        }
    }
    out.tiles_local = slice_of(w, w->tiles, w->rank);
    return OON_OK;
}

// NVTX range per phase of a frame (pack / interact / integrate / sweep / order), visible in nsys timelines and
// usable as an ncu filter (--nvtx --nvtx-include "oon/interact/").
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Timing scaffold for oon_step_profiled: class -> accumulated ms.  Every kernel class is bracketed by its own event
// pair on the launching stream; nothing synchronises with the host until the call has enqueued all its frames (a host
// round trip per kernel would let the ranks of a sharded world drift apart and put launch latency into the numbers).
struct Prof {
    bool on = false;
    float ms[4] = {0, 0, 0, 0};
    uint32_t launches[4] = {0, 0, 0, 0};
    struct Span { cudaEvent_t a, b; int cls; };
    std::vector<Span> spans;
};

template <class F>
int timed(oon_world* w, Prof* prof, int cls, F&& body) {
    Prof::Span sp{nullptr, nullptr, cls};
    if (prof && prof->on) {
        CU(w, cudaEventCreate(&sp.a)); CU(w, cudaEventCreate(&sp.b));
        prof->spans.push_back(sp);
        CU(w, cudaEventRecord(sp.a, w->stream));
    }
    int launched = 0;
    int rc = body(launched);
    if (rc) return rc;
    w->launches += uint64_t(launched);
    if (prof && prof->on) {
        CU(w, cudaEventRecord(sp.b, w->stream));
        prof->launches[cls] += uint32_t(launched);
    }
    return OON_OK;
}

int prof_collect(oon_world* w, Prof* prof) {
    int rc = OON_OK;
    rc = sync_main(w);
    for (Prof::Span& sp : prof->spans) {
        float ms = 0.f;
        if (rc == OON_OK && cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) rc = fail(w, OON_ERR_CUDA, "cudaEventElapsedTime failed");
        prof->ms[sp.cls] += ms;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    prof->spans.clear();
    return rc;
}

int exchange(oon_world* w, Prof* prof) {
    if (w->nranks == 1 || w->p2p) return OON_OK;           // peer-to-peer: the pack kernel has already stored the slice everywhere
    return timed(w, prof, 2, [&](int& launched) -> int {
        const size_t count = size_t(w->block / TS) * TILE_FLOATS;
        NC(w, g_nccl.AllGather(slice_of(w, w->tiles, w->rank), w->tiles, count, ncclFloat, w->comm, w->stream));
        launched = 1;
        return OON_OK;
    });
}

// Space-filling-curve order pays through box culling once the O(N^2) pass dwarfs the O(N) sort (a re-sort costs
// ~0.1 ms; bodies of these worlds cross a tile's width in a few frames, so the order is refreshed every
// `resort_every` frames, default 1).  Small worlds and worlds with mortal bodies keep the identity order.
bool mortal_frame(const oon_world* w, float dt);
// (Round 1 fell back to the identity order for any world with mortal bodies: every tile then takes the checked loop, +30 %.  On
// frames of the mortal schedule the order is now rebuilt from the TRUE positions after every sweep, on the main stream, with the
// multi-kernel radix-sort pipeline (~85 us at 100 k bodies, 3 % of the frame) — the one-CTA-per-block kernel is for hiding
// beside the interact kernel, not for the critical path, and an order computed ahead would index bodies the sweep has moved.)
bool wants_morton(const oon_world* w, const StepParams& sp) {
    return sp.n2n && !sp.exact && w->morton_enabled && w->n_total >= w->morton_min_bodies;
}

// Make inv[] valid for the coming pack: identity for the reference-order kernels, Morton order of the current
// positions for the FAST all-pairs kernel (refreshed every `resort_every` frames: order is a performance matter only).
int ensure_order_buffers(oon_world* w) {
    const uint32_t cap = w->soa[w->cur].cap;
    if (w->order_cap >= cap) return OON_OK;
    if (w->next_order_pending) { CU(w, cudaStreamSynchronize(w->aux)); w->next_order_pending = false; }
    cudaFree(w->order_temp); cudaFree(w->order_keys); cudaFree(w->order_vals); cudaFree(w->order_bounds);
    w->order_temp = nullptr; w->order_keys = nullptr; w->order_vals = nullptr; w->order_bounds = nullptr; w->order_cap = 0;
    w->order_temp_bytes = order_temp_bytes(cap);
    CU(w, cudaMalloc(&w->order_temp, w->order_temp_bytes));
    int rc;
    if ((rc = dev_alloc(w, &w->order_keys, size_t(2) * cap))) return rc;
    if ((rc = dev_alloc(w, &w->order_vals, cap))) return rc;
    if ((rc = dev_alloc(w, &w->order_bounds, 8 * 16))) return rc;     // 8 canonical blocks x ORDER_STAT_WORDS
    w->order_cap = cap;
    return OON_OK;
}

// Make inv[] valid for the coming pack: identity for the reference-order kernels, Hilbert order of the positions for the
// FAST all-pairs kernel of large worlds (order is a performance matter only, never a correctness one).
int ensure_order(oon_world* w, const StepParams& sp, Prof* prof, bool allow_refresh = true) {
    const int kind = wants_morton(w, sp) ? 2 : 1;
    // an order computed ahead on the aux stream from the previous frame's positions: adopt it
    if (allow_refresh && kind == 2 && w->order_kind == 2 && !w->order_dirty && w->next_order_pending) {
        CU(w, cudaStreamWaitEvent(w->stream, w->ev_sorted, 0));
        std::swap(w->inv, w->inv_next);
        w->next_order_pending = false;
        w->order_age = 0;
        w->tiles_fresh = false;
        return OON_OK;
    }
    // allow_refresh == false: between an interact and its integrate the accumulated sums are indexed by the current
    // slots, so only a missing or invalidated order may be (re)built there, never an aged one.
    const bool stale = w->order_dirty || w->order_kind == 0 ||
                       (allow_refresh && (w->order_kind != kind || (kind == 2 && w->order_age >= w->resort_every)));
    if (!stale) return OON_OK;
    if (w->next_order_pending) {                       // the table changed under an order computed ahead: drop it
        CU(w, cudaStreamWaitEvent(w->stream, w->ev_sorted, 0));
        w->next_order_pending = false;
    }
    int rc;
    if (kind == 2 && (rc = ensure_order_buffers(w))) return rc;
    rc = timed(w, prof, 3, [&](int& launched) -> int {
        // (on the critical path — first frame, after edits, after every sweep — the multi-kernel radix-sort pipeline: ~85 us at
        // 100 k bodies against ~300 us for the one-CTA-per-block kernel, which exists to hide beside the interact kernel; same order)
        if (kind == 2)
            launched = launch_order_morton(w->stream, w->soa[w->cur].d, w->d_n, w->block, w->cblock, 0.f, nullptr, w->inv, w->order_temp, w->order_temp_bytes,
                                           w->order_keys, w->order_vals, w->order_bounds);
        else
            launched = launch_order_identity(w->stream, w->inv, w->block);
        CU(w, cudaGetLastError());
        return OON_OK;
    });
    if (rc) return rc;
    w->order_kind = kind; w->order_dirty = false; w->order_age = 0;
    w->tiles_fresh = false;
    return OON_OK;
}

// Start computing the NEXT frame's order on the aux stream from the drift prediction p + v*dt of the state just packed
// (p and v stay untouched until this frame's integrate, which waits for the sort).  Bodies of these worlds move a third
// of a tile per frame, so an order built from this frame's positions would be useless next frame (56 % of the tiles
// checked instead of 8 %); the prediction misses only the coming kick.  Hiding the sort behind the interact kernel
// saves ~85 us per frame.
int start_order_ahead(oon_world* w, const StepParams& sp) {
    NvtxRange range("oon/order_ahead");
    if (!w->async_sort || w->resort_every != 1 || !wants_morton(w, sp) || w->order_kind != 2 || w->next_order_pending) return OON_OK;
    if (mortal_frame(w, sp.dt)) return OON_OK;            // the sweep at the end of this frame shifts indices: the order is rebuilt after it
    int rc = ensure_order_buffers(w);
    if (rc) return rc;
    CU(w, cudaEventRecord(w->ev_main, w->stream));
    CU(w, cudaStreamWaitEvent(w->aux, w->ev_main, 0));
    if (order_fits_one_cta(w->cblock) && w->order_one_cta)
        w->launches += uint64_t(launch_order_block(w->aux, w->soa[w->cur].d, w->d_n, w->block, w->cblock, sp.dt, w->kick_valid ? w->kick : nullptr, w->inv_next));
    else
        w->launches += uint64_t(launch_order_morton(w->aux, w->soa[w->cur].d, w->d_n, w->block, w->cblock, sp.dt, w->kick_valid ? w->kick : nullptr, w->inv_next, w->order_temp,
                                                    w->order_temp_bytes, w->order_keys, w->order_vals, w->order_bounds));
    CU(w, cudaGetLastError());
    CU(w, cudaEventRecord(w->ev_sorted, w->aux));
    w->next_order_pending = true;
    return OON_OK;
}

int do_intrinsic_pack(oon_world* w, const StepParams& sp, bool intrinsic, Prof* prof) {
    NvtxRange range("oon/pack");
    int rc = ensure_order(w, sp, prof);
    if (rc) return rc;
    // The order of the NEXT frame is started here, before this frame's pack and interact kernels are even launched: its
    // CTAs (one per canonical block, a whole SM's registers each) are placed while the GPU is still empty.  Started
    // after the interact kernel they had to wait for an SM to drain (up to one long CTA, ~0.2 ms), and on a 1/8 shard
    // the order then arrived after the interact kernel had finished (measured: +50 us per frame).  The order kernels
    // read p, v, lifetime and the last kick — nothing this frame's pack writes except the player's velocity, which
    // they therefore ignore (order_position).
    if ((rc = start_order_ahead(w, sp))) return rc;
    PackOut out;
    if ((rc = begin_pack(w, out))) return rc;
    rc = timed(w, prof, 1, [&](int& launched) -> int {
        launched = launch_intrinsic_pack(w->stream, w->soa[w->cur].d, w->inv, w->slot_of, w->d_n, w->block, uint32_t(w->rank) * w->block,
                                         out, sp, intrinsic, w->counters);
        CU(w, cudaGetLastError());
        return OON_OK;
    });
    if (rc) return rc;
    rc = exchange(w, prof);
    if (rc) return rc;
    w->tiles_fresh = true;
    return OON_OK;
}

int do_interact(oon_world* w, const StepParams& sp, Prof* prof) {
    NvtxRange range("oon/interact");
    InteractGeometry geo{};
    const bool fast_n2n = sp.n2n && !sp.exact;
    int rc;
    if (fast_n2n) {
        geo = grouping(w);
        if (w->acc_dirty) {                                // only after a (re)allocation or a pass nobody integrated
            CU(w, cudaMemsetAsync(w->acc, 0, sizeof(unsigned long long) * 2 * FX_LIMBS * w->acc_cap, w->stream));
            CU(w, cudaMemsetAsync(w->touch, 0, sizeof(uint32_t) * w->acc_cap, w->stream));
        }
        w->acc_dirty = true;
    } else if ((rc = ensure_partial(w))) return rc;
    const ExchangeView xv = xview(w);
    rc = timed(w, prof, 0, [&](int& launched) -> int {
        if (sp.n2n)
            launched = launch_interact(w->stream, w->tiles, total_tiles(w), geo, uint32_t(w->rank) * w->block, w->block,
                                       w->d_n, w->soa[w->cur].d, uint32_t(w->rank) * w->block, sp, xv, w->acc, w->touch, w->fx_scale, w->partial,
                                       w->counters, w->cap, w->dbg_fraction, w->dbg_part, w->dbg_trace, &w->dbg_trace_ctas,
                                       w->persistent_interact ? w->d_work_ctr : nullptr);
        else
            launched = launch_wait_peers(w->stream, xv) +
                       launch_interact_player_only(w->stream, w->tiles, total_tiles(w), uint32_t(w->rank) * w->block, w->block, w->d_n,
                                                   w->soa[w->cur].d, sp, w->partial, w->scratch4, w->counters, w->cap);
        CU(w, cudaGetLastError());
        return OON_OK;
    });
    if (rc) return rc;
    w->waited_epoch = w->epoch;                            // every kernel above waits for all the slices of this epoch
    return OON_OK;
}

bool wants_kick(const oon_world* w, const StepParams& sp) {   // debug trace, or input of the order prediction
    return w->cap.capacity != 0 || (w->async_sort && w->resort_every == 1 && wants_morton(w, sp));
}

int ensure_kick(oon_world* w, const StepParams& sp) {
    if (!wants_kick(w, sp)) return OON_OK;
    if (w->kick_cap < w->block) {
        cudaFree(w->kick); w->kick = nullptr; w->kick_cap = 0;
        int rc = dev_alloc(w, &w->kick, w->block);
        if (rc) return rc;
        w->kick_cap = w->block;
    }
    return OON_OK;
}

// An order computed ahead is waiting and this integrate may pack the next frame in it (see do_integrate).
bool can_fuse_reorder(const oon_world* w, const StepParams& sp) {
    return w->fuse_reorder && w->async_sort && w->resort_every == 1 && wants_morton(w, sp) && w->order_kind == 2 && !w->order_dirty && w->next_order_pending;
}

int do_integrate(oon_world* w, const StepParams& sp, bool apply_pairwise, bool do_global, bool fuse_next, Prof* prof) {
    NvtxRange range(fuse_next ? "oon/integrate+pack" : "oon/integrate");
    int rc = ensure_kick(w, sp);
    if (rc) return rc;
    if ((rc = ensure_order(w, sp, prof, false))) return rc;
    if (w->next_order_pending && (do_global || apply_pairwise))       // the sort-ahead reads the positions this kernel rewrites
        CU(w, cudaStreamWaitEvent(w->stream, w->ev_sorted, 0));
    // Large FAST worlds re-order their slots every frame (the order of frame k+1 is computed beside frame k's interact
    // kernel).  The fused kernel then packs frame k+1 in the NEW order while it integrates frame k: it runs over the new
    // slots and finds the sums of the body it meets at the slot that body had during the pass (slot_of[], kept by the packs).
    bool reorder = false;
    if (fuse_next && apply_pairwise && can_fuse_reorder(w, sp)) {
        std::swap(w->inv, w->inv_next);
        w->next_order_pending = false;
        w->order_age = 0;
        reorder = true;
    }
    PackOut out{};
    if (fuse_next && (rc = begin_pack(w, out))) return rc;
    rc = timed(w, prof, 1, [&](int& launched) -> int {
        launched = launch_integrate(w->stream, w->soa[w->cur].d, w->inv, w->d_n, w->block, uint32_t(w->rank) * w->block, w->partial,
                                    w->acc, w->touch, w->fx_scale, sp, apply_pairwise, do_global, fuse_next, out,
                                    (apply_pairwise && wants_kick(w, sp)) ? w->kick : nullptr, w->counters, w->slot_of, reorder);
        CU(w, cudaGetLastError());
        return OON_OK;
    });
    if (rc) return rc;
    if (apply_pairwise && wants_kick(w, sp)) w->kick_valid = true;
    if (apply_pairwise && sp.n2n && !sp.exact) w->acc_dirty = false;     // k_integrate has zeroed what it read
    if (do_global || apply_pairwise) w->tiles_fresh = false;
    if (fuse_next) { rc = exchange(w, prof); if (rc) return rc; w->tiles_fresh = true; }
    // the order of the frame after the one just packed: beside the coming interact kernel, from the state this kernel left
    if (reorder && (rc = start_order_ahead(w, sp))) return rc;
    return OON_OK;
}

// ---- in-process collectives of a group ------------------------------------------------------------------------------
// "Every shard has produced its record on its own stream; every shard now needs all of them": each shard records an event
// after producing, waits for the peers' events and pulls their records with peer copies on its own stream.  The events of
// the previous use (ev_grp_done) keep a producer from overwriting a record a peer is still reading.
int group_before_produce(Team& T) {
    for (oon_world* m : T) {
        USE(m);
        for (oon_world* g : T) if (g != m) CU(m, cudaStreamWaitEvent(m->stream, g->ev_grp_done, 0));
    }
    return OON_OK;
}
template <class ArrOf>
int group_allgather(Team& T, ArrOf arr_of, size_t bytes_per_rank) {
    for (oon_world* m : T) { USE(m); CU(m, cudaEventRecord(m->ev_grp, m->stream)); }
    for (oon_world* m : T) {
        USE(m);
        for (oon_world* g : T) {
            if (g == m) continue;
            CU(m, cudaStreamWaitEvent(m->stream, g->ev_grp, 0));
            CU(m, cudaMemcpyPeerAsync(arr_of(m) + size_t(g->rank) * bytes_per_rank, m->device, arr_of(g) + size_t(g->rank) * bytes_per_rank, g->device,
                                      bytes_per_rank, m->stream));
        }
        CU(m, cudaEventRecord(m->ev_grp_done, m->stream));
    }
    return OON_OK;
}
// Stream-ordered all-gather of one small record per rank, whatever kind of team: NCCL between processes, peer copies in a group.
template <class ArrOf>
int team_allgather(Team& T, ArrOf arr_of, size_t bytes_per_rank) {
    if (T.size() > 1) return group_allgather(T, arr_of, bytes_per_rank);
    oon_world* w = T[0];
    if (w->nranks == 1) return OON_OK;
    NC(w, g_nccl.AllGather(arr_of(w) + size_t(w->rank) * bytes_per_rank, arr_of(w), bytes_per_rank, ncclChar, w->comm, w->stream));
    return OON_OK;
}

// The caller's expired-body sweep (OON.cpp:1089-1095) for the whole team.  Sharded: every rank compacts its own index block
// (blocks become ragged: fewer bodies than slots); the one thing that crosses ranks is the parity of a run of expired bodies
// that spans a block boundary: one 16-byte record per rank, all-gathered between the scan and the compaction.  The player
// must live on rank 0.
int team_sweep(Team& T, uint32_t player_ndx, std::vector<Prof>* profs) {
    NvtxRange range("oon/sweep");
    int rc;
    if (T.size() > 1 && (rc = group_before_produce(T))) return rc;
    for (size_t i = 0; i < T.size(); ++i) {
        oon_world* w = T[i];
        USE(w);
        if (w->nranks > 1 && player_ndx >= w->block) return fail(w, OON_ERR_UNSUPPORTED, "sweep on a sharded world: the player must be on rank 0");
        SoAStorage& other = w->soa[w->cur ^ 1];
        if (other.cap < w->soa[w->cur].cap) { if ((rc = soa_alloc(w, other, w->soa[w->cur].cap))) return rc; }
        const size_t tb = sweep_temp_bytes(w->soa[w->cur].cap);
        if (tb > w->sweep_temp_bytes) {
            cudaFree(w->sweep_temp); w->sweep_temp = nullptr;
            CU(w, cudaMalloc(&w->sweep_temp, tb)); w->sweep_temp_bytes = tb;
        }
        if (w->sweep_scratch_cap < w->soa[w->cur].cap) {
            cudaFree(w->sweep_scratch); w->sweep_scratch = nullptr;
            if ((rc = dev_alloc(w, &w->sweep_scratch, size_t(3) * w->soa[w->cur].cap))) return rc;
            w->sweep_scratch_cap = w->soa[w->cur].cap;
        }
        if (w->removed_cap < w->soa[w->cur].cap) {
            cudaFree(w->removed_idx); w->removed_idx = nullptr;
            if ((rc = dev_alloc(w, &w->removed_idx, w->soa[w->cur].cap))) return rc;
            w->removed_cap = w->soa[w->cur].cap;
        }
        const uint32_t first_tested = w->rank == 0 ? player_ndx + 1 : 0u;
        rc = timed(w, profs ? &(*profs)[i] : nullptr, 3, [&](int& launched) -> int {
            CU(w, cudaMemsetAsync(w->d_removed_count, 0, sizeof(uint32_t), w->stream));
            launched = launch_sweep_scan(w->stream, w->soa[w->cur].d, w->d_n, w->block, first_tested, w->sweep_temp, w->sweep_temp_bytes,
                                         w->sweep_scratch, w->nranks > 1 ? w->runinfo + 4 * w->rank : nullptr);
            CU(w, cudaGetLastError());
            return OON_OK;
        });
        if (rc) return rc;
    }
    if ((rc = team_allgather(T, [](oon_world* m) { return reinterpret_cast<char*>(m->runinfo); }, 4 * sizeof(int)))) return rc;
    for (size_t i = 0; i < T.size(); ++i) {
        oon_world* w = T[i];
        USE(w);
        rc = timed(w, profs ? &(*profs)[i] : nullptr, 3, [&](int& launched) -> int {
            CU(w, cudaMemsetAsync(w->d_minfo, 0, 4 * sizeof(uint32_t), w->stream));
            launched = (w->nranks > 1 ? 1 : 0) +
                       launch_sweep_compact(w->stream, w->soa[w->cur].d, w->soa[w->cur ^ 1].d, w->d_n, w->block, w->nranks > 1 ? w->runinfo : nullptr, w->rank,
                                            w->sweep_temp, w->sweep_temp_bytes, w->sweep_scratch, w->removed_idx, w->removed_cap,
                                            w->d_removed_count, w->counters, w->d_minfo);
            CU(w, cudaGetLastError());
            return OON_OK;
        });
        if (rc) return rc;
        w->cur ^= 1;
        w->tiles_fresh = false;
        w->order_dirty = true;          // indices shifted
        w->counts_stale = true;
        w->life_known = false;          // until the next fetch of the counts brings the survivors' summary
    }
    return OON_OK;
}

bool mortal(float lifetime) { return lifetime > 0.f || (lifetime != LIFETIME_UNLIMITED && lifetime <= 0.f); }

// Take the lifetime summaries of all shards (3 words each, see note_lifetime) into the host's schedule state.
void apply_lifetime_info(Team& T, const std::vector<uint32_t>& info /* [nranks][4]: {-, any, pending, ~min bits} */) {
    bool any = false, pending = false;
    double budget = 1e300;
    for (size_t r = 0; r * 4 + 3 < info.size() + 0; ++r) {
        any = any || info[4 * r + 1] != 0;
        pending = pending || info[4 * r + 2] != 0;
        if (info[4 * r + 3]) { const uint32_t bits = ~info[4 * r + 3]; float f; std::memcpy(&f, &bits, sizeof f); budget = std::min(budget, double(f)); }
    }
    for (oon_world* w : T) {
        w->has_mortals = any;
        w->life_known = true;
        w->pending_expired = pending;
        w->life_budget = budget;
    }
}
// A body with this lifetime entered the table through the host (add / set / emit): keep the schedule state conservative.
void note_host_lifetime(Team& T, float life) {
    if (!mortal(life)) return;
    for (oon_world* w : T) {
        w->has_mortals = true;
        if (life > 0.f) w->life_budget = std::min(w->life_budget, double(life)); else w->pending_expired = true;
    }
}
// Does the coming frame need the mortal schedule (sweep, no fusion, identity order)?  Not while the host KNOWS that nothing is
// expired and nothing can expire within this frame and the next (two frames of margin cover the rounding of the device's
// frame-by-frame countdown against the host's running sum).
bool mortal_frame(const oon_world* w, float dt) {
    if (!w->has_mortals) return false;
    if (!w->life_known || w->pending_expired) return true;
    return !(w->life_budget > 3.0 * std::fabs(double(dt)) + 1e-6);
}

// Refresh the host's idea of the body counts after device-side removals (sweeps).
int team_sync_count(Team& T) {
    bool stale = false;
    for (oon_world* w : T) stale = stale || w->counts_stale;
    if (!stale) return OON_OK;
    oon_world* w0 = T[0];
    std::vector<uint32_t> info(size_t(w0->nranks) * 4, 0u);       // per rank {bodies, any mortal, expired pending, ~min positive lifetime}
    if (T.size() > 1 || w0->nranks == 1) {                        // one GPU / group: the host reads every shard's words itself
        for (oon_world* w : T) {
            USE(w);
            CU(w, cudaMemcpyAsync(&info[4 * size_t(w->rank)], w->d_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
            CU(w, cudaMemcpyAsync(&info[4 * size_t(w->rank) + 1], w->d_minfo, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        }
        for (oon_world* w : T) { USE(w); SYNC(w); }
    } else {                                                // one rank of a multi-process world: everybody learns everybody's
        uint32_t* mine = w0->d_counts + 4 * w0->rank;
        CU(w0, cudaMemcpyAsync(mine, w0->d_n, sizeof(uint32_t), cudaMemcpyDeviceToDevice, w0->stream));
        CU(w0, cudaMemcpyAsync(mine + 1, w0->d_minfo, 3 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, w0->stream));
        NC(w0, g_nccl.AllGather(mine, w0->d_counts, 4, ncclUint32, w0->comm, w0->stream));
        CU(w0, cudaMemcpyAsync(info.data(), w0->d_counts, sizeof(uint32_t) * info.size(), cudaMemcpyDeviceToHost, w0->stream));
        SYNC(w0);
    }
    uint64_t total = 0;
    for (int g = 0; g < w0->nranks; ++g) total += info[4 * size_t(g)];
    for (oon_world* w : T) {
        if (w->nranks > 1) for (int g = 0; g < w->nranks; ++g) w->counts[size_t(g)] = info[4 * size_t(g)];
        w->n_local = info[4 * size_t(w->rank)];
        w->n_total = total;
        w->counts_stale = false;
    }
    apply_lifetime_info(T, info);                           // the summaries describe the survivors of the last sweep == the table as it is
    return OON_OK;
}

// Bodies of rank g as the host last learnt them (call team_sync_count first).
uint32_t rank_count(const oon_world* w, int g) { return w->nranks == 1 ? w->n_local : w->counts[size_t(g)]; }
uint64_t rank_first(const oon_world* w, int g) { uint64_t f = 0; for (int h = 0; h < g; ++h) f += rank_count(w, h); return f; }

// nsteps frames for every shard of the team, phase by phase: all packs are enqueued before the first interact kernel that
// waits for them (a group's shards wait for each other's flags ON THE DEVICE; the single host thread must never block on a
// kernel whose peers' producers it has not enqueued yet).
// hook (optional): called with (frame, 0) before the first kernel that belongs to a frame alone and (frame, 1) after its
// last one — the benchmark's per-frame L2 flush and event pairs.  A frame's kernels: [intrinsic+pack unless the previous
// frame's integrate packed it] -> interact -> integrate (+ the next frame's intrinsic+pack when fused) [-> sweep].
typedef std::function<int(uint32_t, int)> FrameHook;
int team_step(Team& T, float dt, uint32_t nsteps, std::vector<Prof>* profs, const FrameHook* hook = nullptr) {
    NvtxRange range("oon/step");
    oon_world* w0 = T[0];
    if (!nsteps || !w0->block) { for (oon_world* w : T) w->steps += nsteps; return OON_OK; }
    const StepParams sp = step_params(w0, dt);
    auto prof_of = [&](size_t i) { return profs ? &(*profs)[i] : nullptr; };
    int rc;
    bool packed = false;                                   // the coming frame's tiles were packed by the previous integrate
    for (uint32_t k = 0; k < nsteps; ++k) {
        // Bodies that can expire bring the sweep, and with it the plain schedule — but only on frames where the host cannot
        // rule an expiry out (mortal_frame); a world whose mortals are all far from their end runs like an immortal one.
        const bool sweep = mortal_frame(w0, dt);
        if (hook && (rc = (*hook)(k, 0))) return rc;
        if (!packed)
            for (size_t i = 0; i < T.size(); ++i) { USE(T[i]); if ((rc = do_intrinsic_pack(T[i], sp, true, prof_of(i)))) return rc; }
        for (size_t i = 0; i < T.size(); ++i) { USE(T[i]); if ((rc = do_interact(T[i], sp, prof_of(i)))) return rc; ++T[i]->order_age; }
        const bool last = k + 1 == nsteps;
        // Without mortal bodies the integrate kernel of frame k also runs frame k+1's intrinsic+pack (one pass over the SoA,
        // one launch) — in the next frame's slot order when that is being refreshed every frame and is ready (do_integrate).
        const bool resort_due = wants_morton(w0, sp) && (w0->order_age >= w0->resort_every || w0->next_order_pending);
        const bool fuse = !sweep && !last && (!resort_due || can_fuse_reorder(w0, sp));      // (mortal_frame's margin covers the next frame's intrinsic too)
        for (size_t i = 0; i < T.size(); ++i) { USE(T[i]); if ((rc = do_integrate(T[i], sp, true, true, fuse, prof_of(i)))) return rc; }
        if (sweep && (rc = team_sweep(T, w0->player_ndx, profs))) return rc;
        for (oon_world* w : T) w->last_frame_swept = sweep;
        for (oon_world* w : T) if (w->life_known) w->life_budget -= double(dt);      // the intrinsic pass counted every lifetime down by dt
        packed = fuse;
        if (hook && (rc = (*hook)(k, 1))) return rc;
    }
    for (oon_world* w : T) w->steps += nsteps;
    return OON_OK;
}

// Host-side edits of the body table must not race with an order being computed ahead on the aux stream.
int quiesce_aux(oon_world* w) {
    if (w->next_order_pending) CU(w, cudaStreamWaitEvent(w->stream, w->ev_sorted, 0));
    return OON_OK;
}
int team_quiesce(Team& T) {
    for (oon_world* w : T) { USE(w); int rc = quiesce_aux(w); if (rc) return rc; }
    return OON_OK;
}
int team_sync(Team& T) {
    for (oon_world* w : T) { USE(w); SYNC(w); }
    return OON_OK;
}

// Which shard of the team (index into T, -1: none of mine) owns global body `ndx`, and where it sits there.
// Needs fresh counts (team_sync_count).
int owner_of(const oon_world* any, uint64_t ndx, uint32_t* local_out) {
    uint64_t first = 0;
    for (int g = 0; g < any->nranks; ++g) {
        const uint32_t c = rank_count(any, g);
        if (ndx < first + c) { *local_out = uint32_t(ndx - first); return g; }
        first += c;
    }
    return -1;
}
oon_world* member_of_rank(Team& T, int rank) {
    for (oon_world* w : T) if (w->rank == rank) return w;
    return nullptr;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// exported entry points
// ------------------------------------------------------------------------------------------------
extern "C" {

int oon_abi_version(void) { return OON_ABI_VERSION; }

int oon_debug_time_interact(oon_world* w, float dt, uint32_t fraction, uint32_t part, float* ms) {
    CHECK_W(w);
    if (!ms) return fail(w, OON_ERR_INVALID, "null output");
    if (!w->kids.empty()) return fail(w, OON_ERR_UNSUPPORTED, "oon_debug_time_interact works on a single-GPU handle");
    if (!w->tiles_fresh || !w->block) return fail(w, OON_ERR_INVALID, "oon_debug_time_interact: call oon_update_intrinsic first");
    const StepParams sp = step_params(w, dt);
    if (!sp.n2n || sp.exact) return fail(w, OON_ERR_UNSUPPORTED, "oon_debug_time_interact times the FAST all-pairs kernel only");
    w->dbg_fraction = fraction; w->dbg_part = part;
    CU(w, cudaEventRecord(w->ev_start, w->stream));
    int rc = do_interact(w, sp, nullptr);
    w->dbg_fraction = 0; w->dbg_part = 0;
    if (rc) return rc;
    CU(w, cudaEventRecord(w->ev_stop, w->stream));
    CU(w, cudaMemsetAsync(w->acc, 0, sizeof(unsigned long long) * 2 * FX_LIMBS * w->acc_cap, w->stream));   // the pass never happened
    CU(w, cudaMemsetAsync(w->touch, 0, sizeof(uint32_t) * w->acc_cap, w->stream));
    w->acc_dirty = false;
    CU(w, cudaEventSynchronize(w->ev_stop));
    CU(w, cudaEventElapsedTime(ms, w->ev_start, w->ev_stop));
    return OON_OK;
}

int oon_debug_trace_interact(oon_world* w, float dt, uint32_t fraction, uint32_t part, uint64_t* trace, uint64_t capacity_ctas, uint64_t* n_ctas_out, float* ms) {
    CHECK_W(w);
    if (!trace || !n_ctas_out || !ms) return fail(w, OON_ERR_INVALID, "null output");
    if (!w->kids.empty()) return fail(w, OON_ERR_UNSUPPORTED, "oon_debug_trace_interact works on a single-GPU handle");
    unsigned long long* dev = nullptr;
    CU(w, cudaMalloc(&dev, sizeof(unsigned long long) * 3 * size_t(capacity_ctas)));
    CU(w, cudaMemsetAsync(dev, 0, sizeof(unsigned long long) * 3 * size_t(capacity_ctas), w->stream));
    w->dbg_trace = dev;
    const int rc = oon_debug_time_interact(w, dt, fraction, part, ms);
    w->dbg_trace = nullptr;
    *n_ctas_out = w->dbg_trace_ctas;
    int rc2 = OON_OK;
    if (rc == OON_OK && w->dbg_trace_ctas > capacity_ctas) rc2 = fail(w, OON_ERR_CAPACITY, "trace needs room for %u CTAs", w->dbg_trace_ctas);
    if (rc == OON_OK && rc2 == OON_OK && cudaMemcpy(trace, dev, sizeof(unsigned long long) * 3 * size_t(w->dbg_trace_ctas), cudaMemcpyDeviceToHost) != cudaSuccess)
        rc2 = fail(w, OON_ERR_CUDA, "trace copy failed");
    cudaFree(dev);
    return rc ? rc : rc2;
}

const char* oon_last_error(const oon_world* w) { return w ? w->error.c_str() : g_create_error.c_str(); }

int oon_props_default(oon_props* out) {
    if (!out) return OON_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->friction = 0.03f;                       // World.cpp:42-50
    out->gravity_mode = OON_GRAVITY_HYPERBOLIC;
    out->gravity = 6.6743e-11;
    out->interact_n2n = 0;
    out->collision_mode = OON_COLLISION_GLIDE_THROUGH;
    out->repulsion_stiffness = 0.0;
    out->loop_mode = OON_LOOP_HALF_MATRIX;
    return OON_OK;
}

int oon_nccl_unique_id(void* out_id) {
    if (!out_id) return OON_ERR_INVALID;
    if (!g_nccl.load()) return fail(nullptr, OON_ERR_NCCL, "%s", g_nccl.error.c_str());
    static_assert(sizeof(ncclUniqueId) == OON_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, OON_ERR_NCCL, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
    std::memcpy(out_id, &id, sizeof id);
    return OON_OK;
}

// parent != nullptr: a shard of a group (no NCCL communicator: the group's host thread does the collectives in-process)
static int create_common(int device, int rank, int nranks, const void* id, oon_world* parent, oon_world** out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return fail(nullptr, OON_ERR_INVALID, "bad arguments to oon_create");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, OON_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    }
    if (device < 0 || device >= count) return fail(nullptr, OON_ERR_INVALID, "device %d out of range (have %d)", device, count);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, OON_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, OON_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major, prop.minor);
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, OON_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    auto* w = new oon_world;
    w->device = device; w->rank = rank; w->nranks = nranks;
    w->team.push_back(w);
    oon_props_default(&w->props);
    w->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("OON_GROUP_TILES")) { long v = atol(e); if (v > 0) w->group_tiles_override = uint32_t(v); }
    if (const char* e = getenv("OON_P2P")) w->p2p_wanted = atoi(e) != 0;
    if (const char* e = getenv("OON_MORTON")) w->morton_enabled = atoi(e) != 0;
    if (const char* e = getenv("OON_DEBUG_GPC")) w->debug_gpc = uint32_t(atoi(e));
    if (const char* e = getenv("OON_DEBUG_TAIL_X2")) w->tail_waves_x2 = uint32_t(atoi(e));
    if (const char* e = getenv("OON_DEBUG_SPC")) w->debug_spc = uint32_t(atoi(e));
    if (const char* e = getenv("OON_SMALL_TILES")) w->small_tiles_override = atoi(e);
    if (const char* e = getenv("OON_ASYNC_SORT")) w->async_sort = atoi(e) != 0;
    if (const char* e = getenv("OON_ORDER_ONE_CTA")) w->order_one_cta = atoi(e) != 0;
    if (const char* e = getenv("OON_MORTON_MIN")) { long v = atol(e); if (v >= 0) w->morton_min_bodies = uint64_t(v); }
    if (const char* e = getenv("OON_RESORT_EVERY")) { long v = atol(e); if (v > 0) w->resort_every = uint32_t(v); }
    if (const char* e = getenv("OON_FUSE_REORDER")) w->fuse_reorder = atoi(e) != 0;
    if (const char* e = getenv("OON_K1_PERSISTENT")) w->persistent_interact = atoi(e) != 0;
    auto bail = [&](int code) { std::string msg = w->error; oon_destroy(w); g_create_error = msg; return code; };
#define CUB_(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { w->error = std::string(#call) + " failed: " + cudaGetErrorString(e_); return bail(OON_ERR_CUDA); } } while (0)
    CUB_(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    CUB_(cudaEventCreate(&w->ev_start)); CUB_(cudaEventCreate(&w->ev_stop));
    CUB_(cudaEventCreate(&w->ev_a)); CUB_(cudaEventCreate(&w->ev_b));
    {
        int lo = 0, hi = 0;
        CUB_(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUB_(cudaStreamCreateWithPriority(&w->aux, cudaStreamNonBlocking, hi));   // small kernels: let them cut in front
    }
    CUB_(cudaEventCreateWithFlags(&w->ev_main, cudaEventDisableTiming)); CUB_(cudaEventCreateWithFlags(&w->ev_sorted, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&w->ev_grp, cudaEventDisableTiming)); CUB_(cudaEventCreateWithFlags(&w->ev_grp_done, cudaEventDisableTiming));
    CUB_(cudaMalloc(&w->d_n, sizeof(uint32_t))); CUB_(cudaMemset(w->d_n, 0, sizeof(uint32_t)));
    CUB_(cudaMalloc(&w->d_removed_count, sizeof(uint32_t))); CUB_(cudaMemset(w->d_removed_count, 0, sizeof(uint32_t)));
    CUB_(cudaMalloc(&w->counters, sizeof(unsigned long long) * C_COUNT));
    CUB_(cudaMemset(w->counters, 0, sizeof(unsigned long long) * C_COUNT));
    CUB_(cudaMalloc(&w->fx_scale, sizeof(int))); CUB_(cudaMemset(w->fx_scale, 0, sizeof(int)));
    CUB_(cudaMalloc(&w->d_minfo, 4 * sizeof(uint32_t))); CUB_(cudaMemset(w->d_minfo, 0, 4 * sizeof(uint32_t)));
    CUB_(cudaMalloc(&w->d_work_ctr, 2 * sizeof(uint32_t))); CUB_(cudaMemset(w->d_work_ctr, 0, 2 * sizeof(uint32_t)));
    CUB_(cudaMalloc(&w->emit_rec, sizeof(EmitBase) * size_t(nranks))); CUB_(cudaMemset(w->emit_rec, 0, sizeof(EmitBase) * size_t(nranks)));
    {
        const uint32_t init[3] = {0u, 0x7f61b1e6u /* BOX_EMPTY */, 0u};
        static_assert(BOX_EMPTY == 3.0e38f, "rank_stats initialiser");
        CUB_(cudaMalloc(&w->rank_stats, sizeof init)); CUB_(cudaMemcpy(w->rank_stats, init, sizeof init, cudaMemcpyHostToDevice));
    }
    if (nranks > 1) {
        CUB_(cudaMalloc(&w->flags, sizeof(uint32_t) * size_t(nranks))); CUB_(cudaMemset(w->flags, 0, sizeof(uint32_t) * size_t(nranks)));
        CUB_(cudaHostAlloc(reinterpret_cast<void**>(&w->timeouts_host), sizeof(uint32_t), cudaHostAllocMapped));
        *w->timeouts_host = 0u;
        CUB_(cudaHostGetDevicePointer(reinterpret_cast<void**>(&w->timeouts), w->timeouts_host, 0));
        CUB_(cudaMalloc(&w->ipc_stage, 256 * size_t(nranks)));
        CUB_(cudaMalloc(&w->runinfo, sizeof(int) * 4 * size_t(nranks))); CUB_(cudaMemset(w->runinfo, 0, sizeof(int) * 4 * size_t(nranks)));
        CUB_(cudaMalloc(&w->d_counts, sizeof(uint32_t) * 4 * size_t(nranks)));     // per rank {bodies, lifetime summary[3]}
        w->counts.assign(size_t(nranks), 0u);
    }
#undef CUB_
    if (nranks > 1 && !parent) {
        if (!id) { w->error = "sharded world needs an NCCL unique id"; return bail(OON_ERR_INVALID); }
        if (!g_nccl.load()) { w->error = g_nccl.error; return bail(OON_ERR_NCCL); }
        ncclUniqueId uid; std::memcpy(&uid, id, sizeof uid);
        ncclResult_t r = g_nccl.CommInitRank(&w->comm, nranks, uid, rank);
        if (r != ncclSuccess) { w->error = std::string("ncclCommInitRank failed: ") + g_nccl.GetErrorString(r); return bail(OON_ERR_NCCL); }
    }
    w->parent = parent;
    *out = w;
    return OON_OK;
}

int oon_create(int device, oon_world** out) { return create_common(device, 0, 1, nullptr, nullptr, out); }
int oon_create_sharded(int device, int rank, int nranks, const void* nccl_unique_id, oon_world** out) {
    return create_common(device, rank, nranks, nccl_unique_id, nullptr, out);
}

// One world over `ndev` GPUs of this process, driven by the calling thread (the reference is one process with a single
// mutator, main.cpp:88-91, World.cpp:196-198): the shards reach each other's exchange buffers through peer access.
int oon_create_multi(const int* devices, int ndev, oon_world** out) {
    if (!out || !devices || ndev < 1 || ndev > 8) return fail(nullptr, OON_ERR_INVALID, "oon_create_multi: 1 to 8 devices");
    *out = nullptr;
    if (ndev == 1) return oon_create(devices[0], out);
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, OON_ERR_INVALID, "oon_create_multi: device %d listed twice", devices[i]);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return fail(nullptr, OON_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path"); }
    for (int i = 0; i < ndev; ++i)
        if (devices[i] < 0 || devices[i] >= count) return fail(nullptr, OON_ERR_INVALID, "device %d out of range (have %d)", devices[i], count);
    for (int i = 0; i < ndev; ++i) {
        if (cudaSetDevice(devices[i]) != cudaSuccess) return fail(nullptr, OON_ERR_CUDA, "cudaSetDevice(%d) failed", devices[i]);
        for (int j = 0; j < ndev; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can)
                return fail(nullptr, OON_ERR_UNSUPPORTED, "oon_create_multi: device %d cannot access device %d (the shards exchange their tiles through peer memory)", devices[i], devices[j]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(nullptr, OON_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", devices[i], devices[j], cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    auto* g = new oon_world;
    g->device = -1; g->rank = 0; g->nranks = ndev;
    oon_props_default(&g->props);
    for (int i = 0; i < ndev; ++i) {
        oon_world* kid = nullptr;
        const int rc = create_common(devices[i], i, ndev, nullptr, g, &kid);
        if (rc) { const std::string msg = g_create_error; oon_destroy(g); g_create_error = msg; return rc; }
        g->kids.push_back(kid);
    }
    g->team = g->kids;
    *out = g;
    return OON_OK;
}

int oon_destroy(oon_world* w) {
    if (!w) return OON_ERR_INVALID;
    if (!w->kids.empty() || w->device < 0) {               // group handle: quiesce every shard, then free them one by one
        for (oon_world* k : w->kids) { cudaSetDevice(k->device); if (k->stream) cudaStreamSynchronize(k->stream); if (k->aux) cudaStreamSynchronize(k->aux); }
        for (oon_world* k : w->kids) { k->p2p = false; oon_destroy(k); }
        delete w;
        return OON_OK;
    }
    cudaSetDevice(w->device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    if (w->aux) { cudaStreamSynchronize(w->aux); cudaStreamDestroy(w->aux); }
    if (w->p2p && w->sticky == OON_OK && !in_group(w)) {
        // Collective: no kernel of any rank still stores into a peer's buffer, and nobody frees a buffer a peer still maps.
        barrier(w);
        p2p_close_tiles(w);
        for (int g = 0; g < 8; ++g) if (w->peer_flags[g]) { cudaIpcCloseMemHandle(w->peer_flags[g]); w->peer_flags[g] = nullptr; }
        barrier(w);
    }
    if (w->ev_main) cudaEventDestroy(w->ev_main);
    if (w->ev_sorted) cudaEventDestroy(w->ev_sorted);
    if (w->ev_grp) cudaEventDestroy(w->ev_grp);
    if (w->ev_grp_done) cudaEventDestroy(w->ev_grp_done);
    cudaFree(w->inv_next); cudaFree(w->slot_of);
    if (w->copy_stream) { cudaStreamSynchronize(w->copy_stream); cudaStreamDestroy(w->copy_stream); cudaEventDestroy(w->ev_render_ready); cudaEventDestroy(w->ev_render_done); }
    if (w->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(w->comm);
    soa_free(w->soa[0]); soa_free(w->soa[1]);
    cudaFree(w->tiles_buf[0]); cudaFree(w->tiles_buf[1]); cudaFree(w->acc); cudaFree(w->touch); cudaFree(w->fx_scale); cudaFree(w->rank_stats);
    cudaFree(w->flags); if (w->timeouts_host) cudaFreeHost(w->timeouts_host); cudaFree(w->ipc_stage); cudaFree(w->runinfo); cudaFree(w->d_counts); cudaFree(w->partial); cudaFree(w->scratch4); cudaFree(w->kick); cudaFree(w->d_n); cudaFree(w->d_removed_count);
    cudaFree(w->counters); cudaFree(w->sweep_temp); cudaFree(w->sweep_scratch); cudaFree(w->removed_idx); cudaFree(w->aos_dev); cudaFree(w->emit_rec); cudaFree(w->d_minfo); cudaFree(w->d_work_ctr);
    cudaFree(w->xyr_dev); cudaFree(w->cap.coll); cudaFree(w->cap.touch);
    cudaFree(w->inv); cudaFree(w->order_temp); cudaFree(w->order_keys); cudaFree(w->order_vals); cudaFree(w->order_bounds);
    if (w->ev_start) cudaEventDestroy(w->ev_start);
    if (w->ev_stop) cudaEventDestroy(w->ev_stop);
    if (w->ev_a) cudaEventDestroy(w->ev_a);
    if (w->ev_b) cudaEventDestroy(w->ev_b);
    if (w->stream) cudaStreamDestroy(w->stream);
    cudaGetLastError();
    delete w;
    return OON_OK;
}

int oon_device_count(const oon_world* w, int* n_out) {
    if (!w || !n_out) return OON_ERR_INVALID;
    *n_out = int(w->team.size());
    return OON_OK;
}

int oon_set_props(oon_world* w, const oon_props* p) {
    CHECK_W(w);
    if (!p) return fail(w, OON_ERR_INVALID, "null props");
    if (p->gravity_mode > OON_GRAVITY_EXPERIMENTAL) return fail(w, OON_ERR_INVALID, "gravity_mode %u out of range", p->gravity_mode);
    if (p->collision_mode > OON_COLLISION_GLIDE_THROUGH) return fail(w, OON_ERR_INVALID, "collision_mode %u out of range", p->collision_mode);
    if (p->loop_mode > OON_LOOP_FULL_MATRIX) return fail(w, OON_ERR_INVALID, "loop_mode %u out of range", p->loop_mode);
    w->props = *p;
    for (oon_world* m : w->team) m->props = *p;
    return OON_OK;
}
int oon_get_props(const oon_world* w, oon_props* out) {
    if (!w || !out) return OON_ERR_INVALID;
    *out = w->props;
    return OON_OK;
}
int oon_set_math_mode(oon_world* w, int mode) {
    CHECK_W(w);
    if (mode != OON_MATH_FAST && mode != OON_MATH_EXACT) return fail(w, OON_ERR_INVALID, "math mode %d out of range", mode);
    w->math_mode = mode;
    for (oon_world* m : w->team) m->math_mode = mode;
    return OON_OK;
}
int oon_get_math_mode(const oon_world* w, int* mode) {
    if (!w || !mode) return OON_ERR_INVALID;
    *mode = w->math_mode;
    return OON_OK;
}
int oon_set_player_index(oon_world* w, uint64_t ndx) {
    CHECK_W(w);
    if (ndx > 0x3fffffffull) return fail(w, OON_ERR_INVALID, "player index %llu out of range", (unsigned long long)ndx);
    w->player_ndx = uint32_t(ndx);
    for (oon_world* m : w->team) m->player_ndx = uint32_t(ndx);
    return OON_OK;
}

// Replace the world by n bodies (with room for n_cap >= n).  `src` = the whole table in host order (is_block == false) or,
// for one rank of a multi-process world, its block of it (is_block == true, `count` bodies starting at global index
// rank * block).  Only a shard's own block crosses the bus.  Whether any body can expire (then the frames include the sweep)
// is found by the unpack kernels, one flag per shard; the ranks of a multi-process world agree on it with a one-word
// all-gather — round 1 walked the whole 60-byte-per-body host table on every rank for this, which cost more than the PCIe
// copy of the block.
static int upload_common(oon_world* h, const oon_body_f32* src, bool is_block, uint64_t count, uint64_t n, uint64_t n_cap) {
    Team& T = h->team;
    int rc = team_quiesce(T);
    if (rc) return rc;
    if (n && !src && !(is_block && !count)) return fail(h, OON_ERR_INVALID, "null body table");
    if (is_block && T.size() > 1) return fail(h, OON_ERR_INVALID, "upload_local on a group handle: the group takes the whole table (oon_upload)");
    if ((rc = team_layout(T, std::max(n, n_cap), n, false))) return rc;
    for (oon_world* w : T) {
        USE(w);
        const uint64_t lo = uint64_t(w->rank) * w->block;
        if (is_block && count != w->n_local)
            return fail(w, OON_ERR_INVALID, "upload_local: rank %d of %d owns %u bodies of a %llu-body world (from index %llu), got %llu",
                        w->rank, w->nranks, w->n_local, (unsigned long long)n, (unsigned long long)lo, (unsigned long long)count);
        const oon_body_f32* mine = is_block ? src : src + lo;
        CU(w, cudaMemsetAsync(w->d_minfo, 0, 4 * sizeof(uint32_t), w->stream));         // lifetime summary of the table, filled by the unpack kernel
        if (w->n_local) {
            if ((rc = ensure_aos(w, w->n_local))) return rc;
            CU(w, cudaMemcpyAsync(w->aos_dev, mine, size_t(w->n_local) * sizeof(oon_body_f32), cudaMemcpyHostToDevice, w->stream));
            w->launches += launch_unpack_f32(w->stream, w->aos_dev, w->soa[w->cur].d, 0, w->n_local, w->d_minfo);
            CU(w, cudaGetLastError());
        }
        CU(w, cudaMemcpyAsync(w->d_n, &w->n_local, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
    }
    std::vector<uint32_t> info(size_t(T[0]->nranks) * 4, 0u);      // per rank {-, any mortal, expired pending, ~min positive lifetime}
    for (oon_world* w : T) {
        USE(w);
        if (T.size() == 1 && w->nranks > 1) {
            uint32_t* mine = w->d_counts + 4 * w->rank;
            CU(w, cudaMemcpyAsync(mine + 1, w->d_minfo, 3 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, w->stream));
            NC(w, g_nccl.AllGather(mine, w->d_counts, 4, ncclUint32, w->comm, w->stream));
            CU(w, cudaMemcpyAsync(info.data(), w->d_counts, sizeof(uint32_t) * info.size(), cudaMemcpyDeviceToHost, w->stream));
        } else {
            CU(w, cudaMemcpyAsync(&info[4 * size_t(w->rank) + 1], w->d_minfo, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
        }
    }
    if ((rc = team_sync(T))) return rc;                    // the caller may reuse the host table immediately (pageable memory)
    apply_lifetime_info(T, info);
    return OON_OK;
}

int oon_upload(oon_world* w, const oon_body_f32* bodies, uint64_t n) {
    CHECK_W(w);
    return upload_common(w, bodies, false, 0, n, n);
}

int oon_upload_local(oon_world* w, const oon_body_f32* block, uint64_t count, uint64_t n_total) {
    CHECK_W(w);
    if (!w->kids.empty()) {                                // a group is one "rank" to its caller: its block is the whole table
        if (count != n_total) return fail(w, OON_ERR_INVALID, "upload_local on a group handle: the group owns the whole table (%llu bodies), got %llu", (unsigned long long)n_total, (unsigned long long)count);
        return upload_common(w, block, false, 0, n_total, n_total);
    }
    return upload_common(w, block, true, count, n_total, n_total);
}

int oon_shard_range(uint64_t n_total, int rank, int nranks, uint64_t* first_out, uint64_t* count_out) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return OON_ERR_INVALID;
    const bool canon = (nranks == 1 || nranks == 2 || nranks == 4 || nranks == 8);      // == layout_one()
    const uint32_t nblocks = canon ? 8u : uint32_t(nranks);
    const uint64_t cblock = std::max<uint64_t>(TS, (((n_total + nblocks - 1) / nblocks) + TS - 1) / TS * TS);
    const uint64_t block = cblock * (nblocks / uint32_t(nranks));
    const uint64_t lo = std::min<uint64_t>(n_total, uint64_t(rank) * block);
    if (first_out) *first_out = lo;
    if (count_out) *count_out = std::min<uint64_t>(block, n_total - lo);
    return OON_OK;
}

int oon_upload_f64(oon_world* w, const oon_body_f64* bodies, uint64_t n) {
    CHECK_W(w);
    if (n && !bodies) return fail(w, OON_ERR_INVALID, "null body table");
    std::vector<oon_body_f32> tmp(n);
    for (uint64_t i = 0; i < n; ++i) {
        const oon_body_f64& s = bodies[i]; oon_body_f32& d = tmp[i];
        std::memset(&d, 0, sizeof d);
        d.gravity_immunity = s.gravity_immunity; d.free_color = s.free_color;
        d.lifetime = float(s.lifetime); d.r = float(s.r); d.density = float(s.density);
        d.px = float(s.px); d.py = float(s.py); d.vx = float(s.vx); d.vy = float(s.vy);
        d.T = s.T; d.color = s.color; d.mass = float(s.mass);
        d.thrust_up = s.thrust_up; d.thrust_down = s.thrust_down; d.thrust_left = s.thrust_left; d.thrust_right = s.thrust_right;
    }
    return oon_upload(w, tmp.data(), n);
}

// The whole world in host order.  One GPU / group: every shard packs its own block and copies it straight to its place in the
// caller's table (a group's shards copy in parallel, one PCIe link each).  One rank of a multi-process world: all-gather of
// the packed blocks, then the whole table (every rank gets everything; use oon_download_local for the rank's own block).
static int download_common(oon_world* h, oon_body_f32* out, uint64_t capacity, uint64_t* n_out) {
    Team& T = h->team;
    int rc = team_sync_count(T);
    if (rc) return rc;
    oon_world* w0 = T[0];
    const uint64_t n = w0->n_total;
    if (n_out) *n_out = n;
    if (n > capacity) return fail(h, OON_ERR_CAPACITY, "download needs room for %llu bodies, got %llu", (unsigned long long)n, (unsigned long long)capacity);
    if (n && !out) return fail(h, OON_ERR_INVALID, "null output");
    if (T.size() == 1 && w0->nranks > 1) {
        oon_world* w = w0;
        const size_t per = size_t(w->block);
        if ((rc = ensure_aos(w, per * w->nranks))) return rc;
        w->launches += launch_pack_f32(w->stream, w->soa[w->cur].d, w->aos_dev + per * w->rank, 0, w->n_local);
        CU(w, cudaGetLastError());
        NC(w, g_nccl.AllGather(w->aos_dev + per * w->rank, w->aos_dev, per * sizeof(oon_body_f32), ncclChar, w->comm, w->stream));
        uint64_t at = 0;
        for (int g = 0; g < w->nranks; ++g) {               // one copy per rank: after a sweep the blocks hold fewer bodies than slots
            const uint32_t c = rank_count(w, g);
            if (c) CU(w, cudaMemcpyAsync(out + at, w->aos_dev + size_t(g) * w->block, size_t(c) * sizeof(oon_body_f32), cudaMemcpyDeviceToHost, w->stream));
            at += c;
        }
        SYNC(w);
        return OON_OK;
    }
    for (oon_world* w : T) {
        USE(w);
        if (!w->n_local) continue;
        if ((rc = ensure_aos(w, w->n_local))) return rc;
        w->launches += launch_pack_f32(w->stream, w->soa[w->cur].d, w->aos_dev, 0, w->n_local);
        CU(w, cudaGetLastError());
        CU(w, cudaMemcpyAsync(out + rank_first(w, w->rank), w->aos_dev, size_t(w->n_local) * sizeof(oon_body_f32), cudaMemcpyDeviceToHost, w->stream));
    }
    return team_sync(T);
}

int oon_download(oon_world* w, oon_body_f32* out, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(w);
    return download_common(w, out, capacity, n_out);
}

// Shard-local read-back: the bodies this rank owns, nothing else — no collective on the data path, one D2H copy of the
// rank's own block (a full oon_download on 8 ranks all-gathers 60 B/body and then copies the whole table on EVERY rank).
int oon_download_local(oon_world* w, oon_body_f32* block_out, uint64_t capacity, uint64_t* first_out, uint64_t* n_out) {
    CHECK_W(w);
    if (!w->kids.empty()) {                                // a group owns the whole table
        if (first_out) *first_out = 0;
        return download_common(w, block_out, capacity, n_out);
    }
    int rc = team_sync_count(w->team);                       // only worlds whose bodies can expire pay for the count exchange
    if (rc) return rc;
    const uint64_t first = rank_first(w, w->rank);
    const uint64_t n = w->n_local;
    if (first_out) *first_out = first;
    if (n_out) *n_out = n;
    if (n > capacity) return fail(w, OON_ERR_CAPACITY, "local download needs room for %llu bodies, got %llu", (unsigned long long)n, (unsigned long long)capacity);
    if (n && !block_out) return fail(w, OON_ERR_INVALID, "null output");
    if (n) {
        if ((rc = ensure_aos(w, n))) return rc;
        w->launches += launch_pack_f32(w->stream, w->soa[w->cur].d, w->aos_dev, 0, uint32_t(n));
        CU(w, cudaGetLastError());
        CU(w, cudaMemcpyAsync(block_out, w->aos_dev, size_t(n) * sizeof(oon_body_f32), cudaMemcpyDeviceToHost, w->stream));
    }
    SYNC(w);
    return OON_OK;
}

int oon_download_f64(oon_world* w, oon_body_f64* out, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(w);
    int rc = team_sync_count(w->team);
    if (rc) return rc;
    std::vector<oon_body_f32> tmp(w->team[0]->n_total);
    uint64_t n = 0;
    rc = oon_download(w, tmp.data(), tmp.size(), &n);
    if (n_out) *n_out = n;
    if (rc) return rc;
    if (n > capacity) return fail(w, OON_ERR_CAPACITY, "download needs room for %llu bodies", (unsigned long long)n);
    for (uint64_t i = 0; i < n; ++i) {
        const oon_body_f32& s = tmp[i]; oon_body_f64& d = out[i];
        std::memset(&d, 0, sizeof d);
        d.gravity_immunity = s.gravity_immunity; d.free_color = s.free_color;
        d.lifetime = s.lifetime; d.r = s.r; d.density = s.density; d.px = s.px; d.py = s.py; d.vx = s.vx; d.vy = s.vy;
        d.T = s.T; d.color = s.color; d.mass = s.mass;
        d.thrust_up = s.thrust_up; d.thrust_down = s.thrust_down; d.thrust_left = s.thrust_left; d.thrust_right = s.thrust_right;
    }
    return OON_OK;
}

// Pack the render read set {x, y, r, colour} of every body on the shards' main streams and copy it to the host on their copy
// streams; the main streams are free for the next frames as soon as the packs are enqueued.  One rank of a multi-process
// world gathers the ranks' blocks first (every rank gets everything).
static int render_enqueue(oon_world* h, float* xyr, uint32_t* colors, uint64_t capacity, uint64_t* n_out) {
    Team& T = h->team;
    int rc = team_sync_count(T);
    if (rc) return rc;
    const uint64_t n = T[0]->n_total;
    if (n_out) *n_out = n;
    if (n > capacity) return fail(h, OON_ERR_CAPACITY, "render download needs room for %llu bodies", (unsigned long long)n);
    if (n && !xyr) return fail(h, OON_ERR_INVALID, "null output");
    for (oon_world* w : T) {
        USE(w);
        if (!w->copy_stream) {
            CU(w, cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
            CU(w, cudaEventCreateWithFlags(&w->ev_render_ready, cudaEventDisableTiming));
            CU(w, cudaEventCreateWithFlags(&w->ev_render_done, cudaEventDisableTiming));
        }
        if (w->render_pending) CU(w, cudaStreamWaitEvent(w->stream, w->ev_render_done, 0));   // the staging buffer is still being read
        const bool gather = T.size() == 1 && w->nranks > 1;
        const size_t per = size_t(w->block) * 4;             // 3 floats + the colour word per slot
        const size_t need = gather ? per * w->nranks : per;
        if (need > w->xyr_cap) {
            if (w->render_pending) CU(w, cudaEventSynchronize(w->ev_render_done));
            cudaFree(w->xyr_dev); w->xyr_dev = nullptr; w->xyr_cap = 0;
            if ((rc = dev_alloc(w, &w->xyr_dev, need + need / 2))) return rc;
            w->xyr_cap = need + need / 2;
        }
        float* mine = w->xyr_dev + (gather ? per * w->rank : 0);
        w->launches += launch_render_pack(w->stream, w->soa[w->cur].d, w->d_n, w->block, mine, reinterpret_cast<uint32_t*>(mine + size_t(w->block) * 3));
        CU(w, cudaGetLastError());
        if (gather) NC(w, g_nccl.AllGather(mine, w->xyr_dev, per, ncclFloat, w->comm, w->stream));
        CU(w, cudaEventRecord(w->ev_render_ready, w->stream));
        CU(w, cudaStreamWaitEvent(w->copy_stream, w->ev_render_ready, 0));
        if (gather) {
            uint64_t at = 0;
            for (int g = 0; g < w->nranks; ++g) {              // one copy per rank (ragged blocks after a sweep)
                const uint32_t c = rank_count(w, g);
                const float* src = w->xyr_dev + per * size_t(g);
                if (c) CU(w, cudaMemcpyAsync(xyr + at * 3, src, size_t(c) * 3 * sizeof(float), cudaMemcpyDeviceToHost, w->copy_stream));
                if (c && colors) CU(w, cudaMemcpyAsync(colors + at, src + size_t(w->block) * 3, size_t(c) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->copy_stream));
                at += c;
            }
        } else if (w->n_local) {
            const uint64_t at = rank_first(w, w->rank);
            CU(w, cudaMemcpyAsync(xyr + at * 3, mine, size_t(w->n_local) * 3 * sizeof(float), cudaMemcpyDeviceToHost, w->copy_stream));
            if (colors) CU(w, cudaMemcpyAsync(colors + at, mine + size_t(w->block) * 3, size_t(w->n_local) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->copy_stream));
        }
        CU(w, cudaEventRecord(w->ev_render_done, w->copy_stream));
        w->render_pending = true;
        w->render_count = n;
    }
    h->render_pending = true;
    h->render_count = n;
    return OON_OK;
}

static int render_wait(oon_world* h, uint64_t* n_out) {
    for (oon_world* w : h->team) {
        USE(w);
        CU(w, cudaEventSynchronize(w->ev_render_done));
        w->render_pending = false;
        int rc = peers_ok(w);
        if (rc) return rc;
    }
    h->render_pending = false;
    if (n_out) *n_out = h->render_count;
    return OON_OK;
}

int oon_download_render(oon_world* w, float* xyr, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(w);
    int rc = render_enqueue(w, xyr, nullptr, capacity, n_out);
    if (rc) return rc;
    return render_wait(w, nullptr);
}

int oon_download_render_rgb(oon_world* w, float* xyr, uint32_t* colors, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(w);
    if (!colors) return fail(w, OON_ERR_INVALID, "null colour output");
    int rc = render_enqueue(w, xyr, colors, capacity, n_out);
    if (rc) return rc;
    return render_wait(w, nullptr);
}

int oon_download_render_begin(oon_world* w, float* xyr_pinned, uint64_t capacity) {
    CHECK_W(w);
    return render_enqueue(w, xyr_pinned, nullptr, capacity, nullptr);
}

int oon_download_render_end(oon_world* w, uint64_t* n_out) {
    CHECK_W(w);
    if (!w->render_pending) return fail(w, OON_ERR_INVALID, "oon_download_render_end without oon_download_render_begin");
    return render_wait(w, n_out);
}

int oon_body_count(oon_world* w, uint64_t* n_out) {
    CHECK_W(w);
    if (!n_out) return fail(w, OON_ERR_INVALID, "null output");
    int rc = team_sync_count(w->team);
    if (rc) return rc;
    *n_out = w->team[0]->n_total;
    return OON_OK;
}

// ---- edits of the body table ------------------------------------------------------------------------------------------
// Sharded worlds (multi-process: every rank makes the same call with the same arguments; group: one call) route an edit to
// the shard that owns the body; the bookkeeping (counts, totals, "bodies can expire") is updated identically everywhere
// without communication.  Appends go to the tail of the table: the last non-empty rank, or the next one when that is full.

// Plan where k more bodies go at the tail of the table: the free slots of the last non-empty rank, then the ranks after it
// (`single`: all on ONE rank — the emitters append from one kernel).  When the tail has no room the whole world is laid out
// afresh for the larger size (through the host: rare).
struct Segment { int rank; uint32_t count; };
static int make_room(oon_world* h, uint64_t k, bool single, std::vector<Segment>* plan) {
    Team& T = h->team;
    plan->clear();
    int rc = team_sync_count(T);
    if (rc) return rc;
    oon_world* w0 = T[0];
    if (k > 0x3fffffffull) return fail(h, OON_ERR_INVALID, "cannot append %llu bodies", (unsigned long long)k);
    if (w0->nranks == 1) {
        if (uint64_t(w0->n_local) + k > w0->block) {
            const uint32_t keep = w0->n_local;
            if ((rc = team_layout(T, uint64_t(keep) + k, keep, true))) return rc;
        }
        plan->push_back({0, uint32_t(k)});
        return OON_OK;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        int last = 0;
        for (int g = 0; g < w0->nranks; ++g) if (rank_count(w0, g)) last = g;
        plan->clear();
        if (single) {
            if (uint64_t(rank_count(w0, last)) + k <= w0->block) { plan->push_back({last, uint32_t(k)}); return OON_OK; }
            if (last + 1 < w0->nranks && k <= w0->block) { plan->push_back({last + 1, uint32_t(k)}); return OON_OK; }
        } else {
            uint64_t remaining = k;
            for (int g = last; g < w0->nranks && remaining; ++g) {
                const uint64_t take = std::min<uint64_t>(w0->block - rank_count(w0, g), remaining);
                if (take) plan->push_back({g, uint32_t(take)});
                remaining -= take;
            }
            if (!remaining) return OON_OK;
        }
        if (attempt) break;
        // no room at the tail: gather the table on the host and lay the world out again with room for k more
        const uint64_t n = w0->n_total;
        std::vector<oon_body_f32> all(n);
        uint64_t got = 0;
        if ((rc = download_common(h, all.data(), n, &got))) return rc;
        const uint64_t n_cap = single ? std::max<uint64_t>(n + k, k * uint64_t(w0->nranks)) : n + k;    // single: a block must hold the burst
        if ((rc = upload_common(h, all.data(), false, 0, n, n_cap))) return rc;     // (the lifetime summary is rebuilt from the table itself)
    }
    return fail(h, OON_ERR_CAPACITY, "cannot append %llu bodies to the sharded world", (unsigned long long)k);
}

static void note_append(Team& T, int rank, uint64_t k) {
    for (oon_world* w : T) {
        if (w->nranks > 1) w->counts[size_t(rank)] += uint32_t(k);
        if (w->rank == rank) w->n_local += uint32_t(k);
        w->n_total += k;
        w->tiles_fresh = false;
        w->order_dirty = true;                             // the slot order does not know the new bodies
        w->kick_valid = false;
    }
}

int oon_add_bodies(oon_world* h, const oon_body_f32* bodies, uint64_t k) {
    CHECK_W(h);
    Team& T = h->team;
    int rc = team_quiesce(T);
    if (rc) return rc;
    if (!k) return OON_OK;
    if (!bodies) return fail(h, OON_ERR_INVALID, "null body table");
    std::vector<Segment> plan;
    if ((rc = make_room(h, k, false, &plan))) return rc;
    for (uint64_t i = 0; i < k; ++i) note_host_lifetime(T, bodies[i].lifetime);
    uint64_t done = 0;
    for (const Segment& seg : plan) {                       // in order: the last non-empty rank is filled to the brim first
        if (oon_world* w = member_of_rank(T, seg.rank)) {
            USE(w);
            if ((rc = ensure_aos(w, seg.count))) return rc;
            const uint32_t at = w->n_local, now = w->n_local + seg.count;
            CU(w, cudaMemcpyAsync(w->aos_dev, bodies + done, size_t(seg.count) * sizeof(oon_body_f32), cudaMemcpyHostToDevice, w->stream));
            w->launches += launch_unpack_f32(w->stream, w->aos_dev, w->soa[w->cur].d, at, seg.count, nullptr);
            CU(w, cudaGetLastError());
            CU(w, cudaMemcpyAsync(w->d_n, &now, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
            SYNC(w);
        }
        note_append(T, seg.rank, seg.count);
        done += seg.count;
    }
    return OON_OK;
}

int oon_emitter_cfg_default(oon_emitter_cfg* out) {
    if (!out) return OON_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    out->v_factor = 0.1f;                            // Emitter.hpp:26-36
    out->offset_velo_factor = 0.2f;
    out->particle_lifetime = LIFETIME_UNLIMITED;
    out->create_mass = 1;
    out->particle_density = 2000.0f * 0.001f;        // Phys::DENSITY_ROCK * 0.001f
    out->position_divergence_x = out->position_divergence_y = 5.f;
    out->velocity_divergence = 1.f;
    out->color = 0x706080u;
    return OON_OK;
}

// The reference's three body-creating loops share one device kernel (k_emit) and this host side.  The emitter, the parent
// and the tail of the table may live on three different shards: the owners publish the fields the loop reads as EmitBase
// records, one small all-gather brings them to the append rank, a second one takes the result (particles appended, the
// emitter's depleted mass and recalculated radius) back.
static int emit_common(oon_world* h, int kind, uint64_t base_ndx, uint64_t parent_ndx, uint32_t n, const oon_emitter_cfg* cfg,
                       const float* nozzles_xy, uint64_t seed, uint32_t* emitted_out, const char* what) {
    Team& T = h->team;
    int rc = team_quiesce(T);
    if (rc) return rc;
    if (emitted_out) *emitted_out = 0;
    if (!cfg) return fail(h, OON_ERR_INVALID, "%s: null configuration", what);
    if ((rc = team_sync_count(T))) return rc;
    oon_world* w0 = T[0];
    if (base_ndx >= w0->n_total || parent_ndx >= w0->n_total)
        return fail(h, OON_ERR_INVALID, "%s: body index %llu out of range (%llu bodies)", what, (unsigned long long)std::max(base_ndx, parent_ndx), (unsigned long long)w0->n_total);
    if (!n) return OON_OK;
    std::vector<Segment> plan;
    if ((rc = make_room(h, n, true, &plan))) return rc;    // room for all n on one rank; the counts are corrected below
    const int arank = plan[0].rank;
    uint32_t elocal = 0, plocal = 0;
    const int erank = owner_of(w0, base_ndx, &elocal), prank = owner_of(w0, parent_ndx, &plocal);
    const bool sharded = w0->nranks > 1;
    auto rec_of = [](oon_world* m) { return reinterpret_cast<char*>(m->emit_rec); };
    if (sharded && T.size() > 1 && (rc = group_before_produce(T))) return rc;
    for (oon_world* w : T) {
        USE(w);
        const int e = w->rank == erank ? int(elocal) : -1, p = w->rank == prank ? int(plocal) : -1;
        if (e >= 0 || p >= 0) { w->launches += uint64_t(launch_emit_fetch(w->stream, w->soa[w->cur].d, e, p, w->emit_rec + w->rank)); CU(w, cudaGetLastError()); }
    }
    if (sharded && (rc = team_allgather(T, rec_of, sizeof(EmitBase)))) return rc;
    if (sharded && T.size() > 1 && (rc = group_before_produce(T))) return rc;
    if (oon_world* w = member_of_rank(T, arank)) {
        USE(w);
        float2* nozzles_dev = nullptr;
        if (nozzles_xy) {
            if ((rc = ensure_aos(w, (size_t(n) * sizeof(float2) + sizeof(oon_body_f32) - 1) / sizeof(oon_body_f32)))) return rc;
            nozzles_dev = reinterpret_cast<float2*>(w->aos_dev);
            CU(w, cudaMemcpyAsync(nozzles_dev, nozzles_xy, size_t(n) * sizeof(float2), cudaMemcpyHostToDevice, w->stream));
        }
        w->launches += uint64_t(launch_emit(w->stream, w->soa[w->cur].d, w->d_n, w->emit_rec + erank, w->emit_rec + prank, w->emit_rec + arank, n, *cfg,
                                            nozzles_dev, seed, kind));
        CU(w, cudaGetLastError());
    }
    if (sharded && (rc = team_allgather(T, rec_of, sizeof(EmitBase)))) return rc;
    if (oon_world* w = member_of_rank(T, erank)) {
        USE(w);
        w->launches += uint64_t(launch_emit_apply(w->stream, w->soa[w->cur].d, elocal, w->emit_rec + arank));
        CU(w, cudaGetLastError());
    }
    EmitBase res{};
    {
        oon_world* w = T[0];                                // every shard holds the append rank's record by now
        USE(w);
        CU(w, cudaMemcpyAsync(&res, w->emit_rec + arank, sizeof res, cudaMemcpyDeviceToHost, w->stream));
    }
    if ((rc = team_sync(T))) return rc;
    note_append(T, arank, res.emitted);
    if (res.emitted && kind != EMIT_SPAWN) note_host_lifetime(T, cfg->particle_lifetime);
    if (emitted_out) *emitted_out = res.emitted;
    return OON_OK;
}

int oon_emit_particles(oon_world* w, uint64_t emitter_ndx, uint32_t n, const oon_emitter_cfg* cfg, const float* nozzles_xy,
                       uint64_t seed, uint32_t* emitted_out) {
    CHECK_W(w);
    return emit_common(w, EMIT_PARTICLES, emitter_ndx, emitter_ndx, n, cfg, nozzles_xy, seed, emitted_out, "emit_particles");
}

int oon_chemtrail_burst(oon_world* w, uint64_t emitter_ndx, uint32_t n, const oon_emitter_cfg* cfg, uint64_t seed, uint32_t* emitted_out) {
    CHECK_W(w);
    return emit_common(w, EMIT_CHEMTRAIL, emitter_ndx, emitter_ndx, n, cfg, nullptr, seed, emitted_out, "chemtrail_burst");
}

int oon_spawn(oon_world* w, uint64_t parent_ndx, uint64_t player_ndx, uint32_t n, uint64_t seed) {
    CHECK_W(w);
    oon_emitter_cfg cfg;
    oon_emitter_cfg_default(&cfg);                         // unused by the spawn loop except as a non-null argument
    return emit_common(w, EMIT_SPAWN, player_ndx, parent_ndx, n, &cfg, nullptr, seed, nullptr, "spawn");
}

int oon_get_body(oon_world* h, uint64_t ndx, oon_body_f32* out) {
    CHECK_W(h);
    if (!out) return fail(h, OON_ERR_INVALID, "null output");
    Team& T = h->team;
    int rc = team_sync_count(T);
    if (rc) return rc;
    oon_world* w0 = T[0];
    if (ndx >= w0->n_total) return fail(h, OON_ERR_INVALID, "body index %llu out of range (%llu bodies)", (unsigned long long)ndx, (unsigned long long)w0->n_total);
    uint32_t local = 0;
    const int orank = owner_of(w0, ndx, &local);
    if (oon_world* w = member_of_rank(T, orank)) {
        USE(w);
        if ((rc = ensure_aos(w, size_t(w->nranks)))) return rc;
        w->launches += launch_pack_f32(w->stream, w->soa[w->cur].d, w->aos_dev + (T.size() == 1 ? w->rank : 0), local, 1);
        CU(w, cudaGetLastError());
        if (T.size() > 1 || w->nranks == 1) {
            CU(w, cudaMemcpyAsync(out, w->aos_dev, sizeof(oon_body_f32), cudaMemcpyDeviceToHost, w->stream));
            SYNC(w);
            return OON_OK;
        }
    }
    // one rank of a multi-process world: the owner's record reaches every rank through a 60-byte-per-rank all-gather
    oon_world* w = w0;
    if ((rc = ensure_aos(w, size_t(w->nranks)))) return rc;
    NC(w, g_nccl.AllGather(w->aos_dev + w->rank, w->aos_dev, sizeof(oon_body_f32), ncclChar, w->comm, w->stream));
    CU(w, cudaMemcpyAsync(out, w->aos_dev + orank, sizeof(oon_body_f32), cudaMemcpyDeviceToHost, w->stream));
    SYNC(w);
    return OON_OK;
}

int oon_set_body(oon_world* h, uint64_t ndx, const oon_body_f32* in) {
    CHECK_W(h);
    Team& T = h->team;
    int rc = team_quiesce(T);
    if (rc) return rc;
    if (!in) return fail(h, OON_ERR_INVALID, "null input");
    if ((rc = team_sync_count(T))) return rc;
    oon_world* w0 = T[0];
    if (ndx >= w0->n_total) return fail(h, OON_ERR_INVALID, "body index %llu out of range (%llu bodies)", (unsigned long long)ndx, (unsigned long long)w0->n_total);
    uint32_t local = 0;
    const int orank = owner_of(w0, ndx, &local);
    if (oon_world* w = member_of_rank(T, orank)) {
        USE(w);
        if ((rc = ensure_aos(w, 1))) return rc;
        CU(w, cudaMemcpyAsync(w->aos_dev, in, sizeof(oon_body_f32), cudaMemcpyHostToDevice, w->stream));
        w->launches += launch_unpack_f32(w->stream, w->aos_dev, w->soa[w->cur].d, local, 1, nullptr);
        CU(w, cudaGetLastError());
        SYNC(w);
    }
    note_host_lifetime(T, in->lifetime);
    for (oon_world* w : T) w->tiles_fresh = false;
    return OON_OK;
}

int oon_remove_body(oon_world* h, uint64_t ndx) {
    CHECK_W(h);
    Team& T = h->team;
    int rc = team_quiesce(T);
    if (rc) return rc;
    if ((rc = team_sync_count(T))) return rc;
    oon_world* w0 = T[0];
    if (ndx >= w0->n_total) return fail(h, OON_ERR_INVALID, "body index %llu out of range (%llu bodies)", (unsigned long long)ndx, (unsigned long long)w0->n_total);
    uint32_t local = 0;
    const int orank = owner_of(w0, ndx, &local);
    if (oon_world* w = member_of_rank(T, orank)) {
        USE(w);
        // erase(begin + ndx): shift the owner's tail down by one, field by field, through the spare SoA buffer
        const uint32_t n = w->n_local, tail = n - local - 1, now = n - 1;
        SoAStorage& cur = w->soa[w->cur];
        SoAStorage& other = w->soa[w->cur ^ 1];
        if (other.cap < cur.cap) { if ((rc = soa_alloc(w, other, cur.cap))) return rc; }
        if (tail) {
#define SHIFT(f) do { CU(w, cudaMemcpyAsync(other.d.f, cur.d.f + local + 1, size_t(tail) * sizeof(*cur.d.f), cudaMemcpyDeviceToDevice, w->stream)); \
                      CU(w, cudaMemcpyAsync(cur.d.f + local, other.d.f, size_t(tail) * sizeof(*cur.d.f), cudaMemcpyDeviceToDevice, w->stream)); } while (0)
            SHIFT(p); SHIFT(v); SHIFT(T); SHIFT(life); SHIFT(mass); SHIFT(r); SHIFT(density); SHIFT(color); SHIFT(flags); SHIFT(thrust);
#undef SHIFT
        }
        CU(w, cudaMemcpyAsync(w->d_n, &now, sizeof(uint32_t), cudaMemcpyHostToDevice, w->stream));
        SYNC(w);
    }
    for (oon_world* w : T) {
        if (w->nranks > 1) w->counts[size_t(orank)] -= 1u;
        if (w->rank == orank) w->n_local -= 1u;
        w->n_total -= 1u;
        w->tiles_fresh = false;
        w->order_dirty = true;
        w->kick_valid = false;
    }
    return OON_OK;
}

int oon_step(oon_world* w, float dt, uint32_t nsteps) {
    CHECK_W(w);
    return team_step(w->team, dt, nsteps, nullptr);
}

// per-class device times of a team: the slowest shard's
static int collect_profs(Team& T, std::vector<Prof>& profs, float ms[4], uint32_t launches[4]) {
    int rc = OON_OK;
    float out[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < T.size(); ++i) {
        if (T[i]->parent) cudaSetDevice(T[i]->device);
        const int crc = prof_collect(T[i], &profs[i]);
        if (rc == OON_OK) rc = crc;
        for (int c = 0; c < 4; ++c) out[c] = std::max(out[c], profs[i].ms[c]);
    }
    if (ms) std::memcpy(ms, out, sizeof out);
    if (launches) std::memcpy(launches, profs[0].launches, sizeof profs[0].launches);
    return rc;
}

int oon_step_profiled(oon_world* w, float dt, uint32_t nsteps, float ms[4], uint32_t launches[4]) {
    CHECK_W(w);
    std::vector<Prof> profs(w->team.size());
    for (Prof& p : profs) p.on = true;
    int rc = team_step(w->team, dt, nsteps, &profs);
    const int crc = collect_profs(w->team, profs, ms, launches);
    return rc ? rc : crc;
}

int oon_update_intrinsic(oon_world* h, float dt) {
    CHECK_W(h);
    Team& T = h->team;
    if (!T[0]->block) return OON_OK;
    for (oon_world* w : T) { USE(w); int rc = do_intrinsic_pack(w, step_params(w, dt), true, nullptr); if (rc) return rc; }
    for (oon_world* w : T) if (w->life_known) w->life_budget -= double(dt);          // every lifetime was counted down by dt
    return OON_OK;
}

int oon_update_pairwise(oon_world* h, float dt) {
    CHECK_W(h);
    Team& T = h->team;
    if (!T[0]->block) return OON_OK;
    const StepParams sp = step_params(T[0], dt);
    int rc;
    for (oon_world* w : T) { USE(w); if (!w->tiles_fresh && (rc = do_intrinsic_pack(w, sp, false, nullptr))) return rc; }
    for (oon_world* w : T) { USE(w); if ((rc = do_interact(w, sp, nullptr))) return rc; ++w->order_age; }
    for (oon_world* w : T) { USE(w); if ((rc = do_integrate(w, sp, true, false, false, nullptr))) return rc; }
    return OON_OK;
}

int oon_update_global(oon_world* h, float dt) {
    CHECK_W(h);
    Team& T = h->team;
    if (!T[0]->block) return OON_OK;
    for (oon_world* w : T) { USE(w); int rc = do_integrate(w, step_params(w, dt), false, true, false, nullptr); if (rc) return rc; }
    return OON_OK;
}

int oon_sweep_expired(oon_world* h, uint64_t player_ndx) {
    CHECK_W(h);
    Team& T = h->team;
    for (oon_world* w : T) w->last_frame_swept = false;
    if (!T[0]->block || !T[0]->has_mortals) return OON_OK;
    for (oon_world* w : T) w->last_frame_swept = true;
    return team_sweep(T, uint32_t(player_ndx), nullptr);
}

int oon_synchronize(oon_world* w) {
    CHECK_W(w);
    return team_sync(w->team);
}

int oon_exchange_kind(const oon_world* w, int* kind) {
    if (!w || !kind) return OON_ERR_INVALID;
    const oon_world* m = w->team[0];
    *kind = m->nranks == 1 ? OON_EXCHANGE_NONE : (m->p2p ? OON_EXCHANGE_PEER_STORES : OON_EXCHANGE_NCCL_ALLGATHER);
    return OON_OK;
}

int oon_drain_events(oon_world* h, oon_events* out) {
    CHECK_W(h);
    if (!out) return fail(h, OON_ERR_INVALID, "null output");
    Team& T = h->team;
    std::vector<unsigned long long> c(size_t(C_COUNT) * T.size(), 0ull);
    for (size_t i = 0; i < T.size(); ++i) {
        oon_world* w = T[i];
        USE(w);
        CU(w, cudaMemcpyAsync(&c[size_t(C_COUNT) * i], w->counters, sizeof(unsigned long long) * C_COUNT, cudaMemcpyDeviceToHost, w->stream));
        CU(w, cudaMemsetAsync(w->counters, 0, sizeof(unsigned long long) * C_CAPTURE_COLL, w->stream));   // capture cursors live on
    }
    int rc = team_sync(T);
    if (rc) return rc;
    std::memset(out, 0, sizeof *out);
    out->steps = T[0]->steps;
    for (size_t i = 0; i < T.size(); ++i) {                 // each visit is counted on exactly one shard (the target's owner)
        const unsigned long long* k = &c[size_t(C_COUNT) * i];
        T[i]->steps = 0;
        out->collisions += k[C_COLLISIONS]; out->touches += k[C_TOUCHES]; out->player_touches += k[C_PLAYER_TOUCHES];
        out->terminated += k[C_TERMINATED]; out->removed += k[C_REMOVED];
        out->warp_tiles += k[C_WARP_TILES]; out->warp_tiles_checked += k[C_WARP_TILES_CHECKED];
    }
    return OON_OK;
}

int oon_timer_start(oon_world* h) {
    CHECK_W(h);
    for (oon_world* w : h->team) { USE(w); CU(w, cudaEventRecord(w->ev_start, w->stream)); }
    return OON_OK;
}
int oon_timer_stop(oon_world* h, float* elapsed_ms) {
    CHECK_W(h);
    if (!elapsed_ms) return fail(h, OON_ERR_INVALID, "null output");
    *elapsed_ms = 0.f;
    for (oon_world* w : h->team) { USE(w); CU(w, cudaEventRecord(w->ev_stop, w->stream)); }
    for (oon_world* w : h->team) {
        USE(w);
        float ms = 0.f;
        CU(w, cudaEventSynchronize(w->ev_stop));
        CU(w, cudaEventElapsedTime(&ms, w->ev_start, w->ev_stop));
        *elapsed_ms = std::max(*elapsed_ms, ms);           // a group's frame ends when its slowest shard does
        int rc = peers_ok(w);
        if (rc) return rc;
    }
    return OON_OK;
}

static int bench_frames_impl(oon_world* h, float dt, uint32_t nsteps, uint64_t flush_bytes, double* ms_sum, std::vector<Prof>* profs) {
    CHECK_W(h);
    if (!ms_sum) return fail(h, OON_ERR_INVALID, "null output");
    *ms_sum = 0.0;
    if (!nsteps) return OON_OK;
    Team& T = h->team;
    const size_t M = T.size();
    std::vector<void*> flush(M, nullptr);
    std::vector<cudaEvent_t> ev(size_t(2) * nsteps * M, nullptr);
    int rc = OON_OK;
    auto body = [&]() -> int {
        for (size_t i = 0; i < M; ++i) {
            USE(T[i]);
            if (flush_bytes) CU(T[i], cudaMalloc(&flush[i], flush_bytes));
            for (uint32_t k = 0; k < 2 * nsteps; ++k) CU(T[i], cudaEventCreate(&ev[(size_t(k) * M) + i]));
        }
        // ONE multi-frame step (the schedule oon_step(w, dt, nsteps) runs, fused launches included) with the flush and the
        // event pair of every frame hooked in at the frame's boundaries.
        const FrameHook hook = [&](uint32_t k, int phase) -> int {
            if (phase == 1) {
                for (size_t i = 0; i < M; ++i) { USE(T[i]); CU(T[i], cudaEventRecord(ev[(size_t(2 * k + 1) * M) + i], T[i]->stream)); }
                return OON_OK;
            }
            for (size_t i = 0; i < M; ++i) {
                oon_world* w = T[i];
                USE(w);
                if (flush_bytes) CU(w, cudaMemsetAsync(flush[i], int(k & 0xff), flush_bytes, w->stream));
                // The shards of a world wait for each other inside the frame: line them up again after the untimed flush (a
                // one-word all-gather / a round of event waits), or the slowest shard's flush would be charged to everybody's frame.
                if (flush_bytes && M == 1 && w->nranks > 1)
                    NC(w, g_nccl.AllGather(static_cast<char*>(w->ipc_stage) + size_t(w->rank) * 4, w->ipc_stage, 4, ncclChar, w->comm, w->stream));
                if (M > 1) CU(w, cudaEventRecord(w->ev_grp, w->stream));
            }
            for (size_t i = 0; i < M; ++i) {
                oon_world* w = T[i];
                USE(w);
                if (M > 1) for (oon_world* g : T) if (g != w) CU(w, cudaStreamWaitEvent(w->stream, g->ev_grp, 0));
                CU(w, cudaEventRecord(ev[(size_t(2 * k) * M) + i], w->stream));
            }
            return OON_OK;
        };
        int src = team_step(T, dt, nsteps, profs, &hook);
        if (src) return src;
        src = team_sync(T);
        if (src) return src;
        for (uint32_t k = 0; k < nsteps; ++k) {
            float worst = 0.f;
            for (size_t i = 0; i < M; ++i) {
                USE(T[i]);
                float ms = 0.f;
                CU(T[i], cudaEventElapsedTime(&ms, ev[(size_t(2 * k) * M) + i], ev[(size_t(2 * k + 1) * M) + i]));
                worst = std::max(worst, ms);
            }
            *ms_sum += double(worst);
        }
        return OON_OK;
    };
    rc = body();
    for (size_t i = 0; i < M; ++i) {
        if (T[i]->parent) cudaSetDevice(T[i]->device);
        if (rc != OON_OK) cudaStreamSynchronize(T[i]->stream);
        for (uint32_t k = 0; k < 2 * nsteps; ++k) if (ev[(size_t(k) * M) + i]) cudaEventDestroy(ev[(size_t(k) * M) + i]);
        if (flush[i]) cudaFree(flush[i]);
    }
    return rc;
}

int oon_bench_frames(oon_world* w, float dt, uint32_t nsteps, uint64_t flush_bytes, double* ms_sum) {
    return bench_frames_impl(w, dt, nsteps, flush_bytes, ms_sum, nullptr);
}

int oon_bench_frames_profiled(oon_world* w, float dt, uint32_t nsteps, uint64_t flush_bytes, double* ms_sum, float ms[4], uint32_t launches[4]) {
    if (!w) return OON_ERR_INVALID;
    std::vector<Prof> profs(w->team.size());
    for (Prof& p : profs) p.on = true;
    const int rc = bench_frames_impl(w, dt, nsteps, flush_bytes, ms_sum, &profs);
    const int crc = collect_profs(w->team, profs, ms, launches);
    return rc ? rc : crc;
}

int oon_launch_count(const oon_world* w, uint64_t* n_out) {
    if (!w || !n_out) return OON_ERR_INVALID;
    *n_out = 0;
    for (const oon_world* m : w->team) *n_out += m->launches;
    return OON_OK;
}

int oon_measure_fp32_peak(int device, double* tflops_out) {
    if (!tflops_out) return OON_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, OON_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", device); }
    const float t = measure_fp32_peak_tflops(nullptr);
    if (t <= 0) return fail(nullptr, OON_ERR_CUDA, "FFMA probe failed");
    *tflops_out = t;
    return OON_OK;
}
int oon_measure_copy_bandwidth(int device, double* gbs_out) {
    if (!gbs_out) return OON_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, OON_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", device); }
    const float g = measure_copy_gbs(nullptr);
    if (g <= 0) return fail(nullptr, OON_ERR_CUDA, "copy probe failed");
    *gbs_out = g;
    return OON_OK;
}

int oon_debug_ieee_selftest(int device, uint64_t n, uint64_t seed, uint64_t* mismatches_out) {
    if (!mismatches_out) return OON_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, OON_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", device); }
    unsigned long long* d = nullptr;
    unsigned long long h = 0;
    if (cudaMalloc(&d, sizeof h) != cudaSuccess || cudaMemset(d, 0, sizeof h) != cudaSuccess) return fail(nullptr, OON_ERR_CUDA, "selftest allocation failed");
    run_ieee_selftest(nullptr, n, seed, d);
    const cudaError_t e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(nullptr, OON_ERR_CUDA, "selftest failed: %s", cudaGetErrorString(e));
    *mismatches_out = h;
    return OON_OK;
}

int oon_debug_capture_pairs(oon_world* h, uint64_t capacity) {
    CHECK_W(h);
    if (capacity > 0x7fffffffull) return fail(h, OON_ERR_INVALID, "capture capacity too large");
    for (oon_world* w : h->team) {
        USE(w);
        SYNC(w);
        cudaFree(w->cap.coll); cudaFree(w->cap.touch);
        w->cap = CaptureBuf{};
        if (capacity) {
            int rc;
            if ((rc = dev_alloc(w, &w->cap.coll, capacity))) return rc;
            if ((rc = dev_alloc(w, &w->cap.touch, capacity))) return rc;
            w->cap.capacity = uint32_t(capacity);
        }
        CU(w, cudaMemset(w->counters + C_CAPTURE_COLL, 0, 2 * sizeof(unsigned long long)));
        w->kick_valid = false;
    }
    return OON_OK;
}

int oon_debug_fetch_pairs(oon_world* h, int which, uint32_t* pairs, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(h);
    if (which != 0 && which != 1) return fail(h, OON_ERR_INVALID, "which must be 0 or 1");
    const int ci = which == 0 ? C_CAPTURE_COLL : C_CAPTURE_TOUCH;
    uint64_t total = 0;
    for (oon_world* w : h->team) {                          // a group's lists are concatenated (each visit is recorded by the target's owner)
        USE(w);
        if (!w->cap.capacity) return fail(h, OON_ERR_INVALID, "pair capture is off");
        unsigned long long cnt = 0;
        CU(w, cudaMemcpyAsync(&cnt, w->counters + ci, sizeof cnt, cudaMemcpyDeviceToHost, w->stream));
        SYNC(w);
        if (cnt > w->cap.capacity) { if (n_out) *n_out = total + cnt; return fail(h, OON_ERR_CAPACITY, "pair capture overflowed: %llu pairs, capacity %u", cnt, w->cap.capacity); }
        if (total + cnt > capacity) { if (n_out) *n_out = total + cnt; return fail(h, OON_ERR_CAPACITY, "output too small for %llu pairs", (unsigned long long)(total + cnt)); }
        if (cnt) {
            if (!pairs) return fail(h, OON_ERR_INVALID, "null output");
            CU(w, cudaMemcpyAsync(pairs + 2 * total, which == 0 ? w->cap.coll : w->cap.touch, size_t(cnt) * sizeof(uint2), cudaMemcpyDeviceToHost, w->stream));
        }
        CU(w, cudaMemsetAsync(w->counters + ci, 0, sizeof(unsigned long long), w->stream));
        SYNC(w);
        total += cnt;
    }
    if (n_out) *n_out = total;
    return OON_OK;
}

int oon_debug_fetch_kick(oon_world* h, float* dv, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(h);
    Team& T = h->team;
    if (T.size() == 1 && T[0]->nranks > 1) return fail(h, OON_ERR_UNSUPPORTED, "kick trace on one rank of a multi-process world");
    int rc = team_sync_count(T);
    if (rc) return rc;
    const uint64_t n = T[0]->n_total;
    if (n_out) *n_out = n;
    if (n > capacity) return fail(h, OON_ERR_CAPACITY, "output too small for %llu bodies", (unsigned long long)n);
    for (oon_world* w : T) {
        USE(w);
        if (!w->kick_valid) return fail(h, OON_ERR_INVALID, "no kick recorded: enable oon_debug_capture_pairs and run a pairwise pass first");
        if (w->n_local) CU(w, cudaMemcpyAsync(dv + 2 * rank_first(w, w->rank), w->kick, size_t(w->n_local) * sizeof(float2), cudaMemcpyDeviceToHost, w->stream));
    }
    return team_sync(T);
}

// Indices erased by the last sweep, at the time of removal like the oracle's trace (OON.cpp:1089-1095 erases in ascending
// order, so when a body goes every rank before it has finished: its index then = survivors of the earlier ranks + its local
// index at removal).
int oon_debug_fetch_removed(oon_world* h, uint32_t* idx, uint64_t capacity, uint64_t* n_out) {
    CHECK_W(h);
    Team& T = h->team;
    int rc = team_sync_count(T);
    if (rc) return rc;
    uint64_t total = 0;
    for (oon_world* w : T) {
        USE(w);
        if (!w->last_frame_swept) continue;                  // the last frame ran without a sweep (no expiry was possible): nothing was removed
        uint32_t cnt = 0;
        CU(w, cudaMemcpyAsync(&cnt, w->d_removed_count, sizeof cnt, cudaMemcpyDeviceToHost, w->stream));
        SYNC(w);
        if (total + cnt > capacity) { if (n_out) *n_out = total + cnt; return fail(h, OON_ERR_CAPACITY, "output too small for %llu indices", (unsigned long long)(total + cnt)); }
        if (cnt) {
            if (!idx) return fail(h, OON_ERR_INVALID, "null output");
            CU(w, cudaMemcpyAsync(idx + total, w->removed_idx, size_t(cnt) * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream));
            SYNC(w);
            const uint32_t off = uint32_t(rank_first(w, w->rank));
            for (uint32_t k = 0; k < cnt; ++k) idx[total + k] += off;
        }
        total += cnt;
    }
    if (n_out) *n_out = total;
    return OON_OK;
}

}  // extern "C"
