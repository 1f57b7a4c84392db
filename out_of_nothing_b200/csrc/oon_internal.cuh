// oon_internal.cuh — shared declarations between the kernels (oon_kernels.cu) and the C ABI (oon_capi.cu).
// sm_100a only.  Names follow the reference's domain (bodies, sources, targets, tiles, shards).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/oon_gpu.h"

namespace oon {

// ---- geometry of the source exchange buffer ---------------------------------------------------
// Sources live in "tiles" of TS bodies laid out as five contiguous arrays
//     x[TS], y[TS], m[TS], r[TS], id[TS]   (id = global body index, uint32 bits)
// followed by an 8-float header {min x, min y, max x, max y of the live bodies, max live radius, -, -, -}
// (AoSoA).  One tile is one 5152-byte TMA bulk copy into shared memory and lands exactly in the layout the
// packed-FP32 inner loop reads (two sources per 64-bit register pair).  Slot s of the buffer holds body
// inv[s]: in FAST all-pairs mode the slots follow a Hilbert curve through the positions, so a tile is a compact
// patch of space and its header box lets a warp skip the collision tests of far-away tiles wholesale.
constexpr int TS = 256;                    // bodies per tile
constexpr int TILE_HDR = 5 * TS;           // float offset of the header
constexpr int TILE_FLOATS = 5 * TS + 8;    // floats per tile
constexpr int TILE_BYTES = TILE_FLOATS * 4;
constexpr int CTA_THREADS = 256;
#ifndef OON_K1_THREADS
#define OON_K1_THREADS 256
#endif
#ifndef OON_K1_T
#define OON_K1_T 1
#endif
constexpr int K1_TARGETS_PER_CTA = OON_K1_THREADS * OON_K1_T;   // target slots per CTA of the FAST interact kernel
constexpr float BOX_EMPTY = 3.0e38f;       // |coordinate| of an empty box (finite on purpose)
constexpr uint32_t ID_NONE = 0xffffffffu;
constexpr uint32_t ID_PLAYER_BIT = 0x80000000u;   // id word of a tile slot: global body index (< 2^30) | this bit for a body with thrusters

// A dead slot (padding, or a body with lifetime == 0) is parked far away with no mass and a
// negative radius: it attracts nothing, never collides and never comes near anything.
constexpr float DEAD_COORD = 1.0e18f;
constexpr float DEAD_RADIUS = -1.0e30f;

constexpr float EPS_COLLISION = 50000000.0f / 10;   // World.hpp:91, World.cpp:315 (a float constant)
constexpr float THRUST_UNSET = 2e31f;               // Math::unset<float>() as stored in the fixtures
constexpr float LIFETIME_UNLIMITED = -1.0f;         // Entity::Unlimited, Entity.hpp:34

// per-body flag bits (device SoA)
enum : uint8_t { F_IMMUNE = 1, F_FREE_COLOR = 2, F_PLAYER = 4 };

// counters kept on the device (uint64 each)
enum { C_COLLISIONS = 0, C_TOUCHES, C_PLAYER_TOUCHES, C_TERMINATED, C_REMOVED, C_WARP_TILES, C_WARP_TILES_CHECKED, C_CAPTURE_COLL, C_CAPTURE_TOUCH, C_COUNT };

// One step's constant parameters, passed by value to the kernels.
struct StepParams {
    float dt;
    float friction;
    float gravity;              // G as float (NumType(props.gravity))
    double repulsion;           // repulsion_stiffness (double in the reference)
    uint32_t gravity_mode;      // OON_GRAVITY_*
    uint32_t collision_on;      // collision_mode != Off
    uint32_t n2n;               // interact_n2n
    uint32_t full_matrix;       // loop_mode == Full_Matrix
    uint32_t exact;             // OON_MATH_EXACT
};

// Device view of one shard's private state (structure of arrays, local index order == host order).
struct BodySoA {
    float2* p;          // position
    float2* v;          // velocity
    float* T;           // temperature
    float* life;        // lifetime
    float* mass;
    float* r;           // radius
    float* density;
    uint32_t* color;
    uint8_t* flags;     // F_*
    float4* thrust;     // up, down, left, right
};

struct CaptureBuf {
    uint2* coll;        // (target, source)
    uint2* touch;
    uint32_t capacity;  // 0 = capture off
};

// Where a pack kernel writes this rank's slice of the exchange tiles.
constexpr int MAX_PEERS = 7;
struct PackOut {
    float* tiles_local;              // this rank's slice inside its own exchange buffer
    uint32_t* rank_stats;            // {max |coordinate| bits, min non-zero |mass| bits, block ticket}
    uint32_t npeers;                 // peer-to-peer exchange: remote ranks the slice is also stored to (0: NCCL all-gather / one GPU)
    uint32_t epoch;                  // value this rank's flag is raised to once the slice has landed everywhere
    float* peer_tiles[MAX_PEERS];    // this rank's slice inside each peer's exchange buffer (NVLink peer memory)
    uint32_t* peer_flags[MAX_PEERS]; // this rank's flag at each peer
};

// What the FAST all-pairs kernel needs to know about the exchange.
struct ExchangeView {
    uint32_t nranks;                 // slices in the exchange buffer
    uint32_t tiles_per_rank;
    uint32_t rank;                   // own slice
    uint32_t epoch;                  // peer-to-peer exchange: slice g may be read once flags[g] >= epoch; 0: no flags to wait for
    const uint32_t* flags;           // [nranks], written by the peers
    uint32_t* timeouts;              // counts waits that gave up (a peer that never arrived): sticky error on the host
};

// ---- exact accumulation of the per-group sums ---------------------------------------------------
// The FAST all-pairs kernel sums in FP32 inside a canonical GROUP of source tiles (a function of the world size
// only) and adds the group sums of a target in fixed point: value = (w0 + w1*2^40 + w2*2^80) * 2^B with three
// signed 64-bit limbs per component (carry-save: limbs are summed independently with integer atomics and
// normalised once in k_integrate).  Integer addition is associative, so the result does not depend on which
// CTA, in which order, on which GPU handled a group: any launch geometry and any sharding give the same bits.
// B comes from the world statistics in the tile headers; the 120-bit window leaves 2^88 for the product of the
// mass ratio and the distance ratio of the world before the smallest addend loses FP32 precision.
constexpr int FX_LIMB_BITS = 40;
constexpr int FX_LIMBS = 3;
constexpr int FX_MAX_SHIFT = 96;     // 24-bit mantissa << 96 = 120 bits

// ---- launchers (oon_kernels.cu) ------------------------------------------------------------------
// All take the stream to launch on and return the number of kernels launched.
// `inv` maps a local slot of the exchange buffer to the local body stored there (identity unless sorted).

// AoS (60 B records) <-> SoA
// minfo (optional, 4 words, zeroed by the caller): what the unpacked bodies can still do — [0] any finite lifetime, [1] any
// expired body (lifetime <= 0, not Unlimited), [2] ~bits of the smallest positive lifetime (0: none)
int launch_unpack_f32(cudaStream_t st, const oon_body_f32* aos, BodySoA soa, uint32_t first, uint32_t count, uint32_t* minfo);
int launch_pack_f32(cudaStream_t st, BodySoA soa, oon_body_f32* aos, uint32_t first, uint32_t count);

// The reference's three body-creating loops on the device: up to n new bodies appended at index *d_n (which is
// advanced).  The body they are created around (spawn: the player) and the body a spawned one inherits T and v from are
// passed as EmitBase records, so that they may live on another rank than the tail of the table: launch_emit_fetch (owner
// ranks) -> records travel -> launch_emit (append rank; fills the result part of *result) -> launch_emit_apply (emitter's owner).
struct EmitBase {
    float2 ep, ev;                   // emitter: position, velocity
    float er, em, edens;             // emitter: radius, mass, density
    float pT;                        // parent: temperature
    float2 pv;                       // parent: velocity
    float new_mass, new_r;           // result: the emitter after the burst ...
    uint32_t write_mass, write_r;    // ... and which of the two fields changed
    uint32_t emitted;                // result: particles appended
    uint32_t _pad;
};
static_assert(sizeof(EmitBase) == 64, "EmitBase is one 64-byte record of the all-gather staging");
enum { EMIT_PARTICLES = 0, EMIT_CHEMTRAIL = 1, EMIT_SPAWN = 2 };
int launch_emit_fetch(cudaStream_t st, BodySoA soa, int emitter_local /* -1: not mine */, int parent_local /* -1: not mine */, EmitBase* out);
int launch_emit(cudaStream_t st, BodySoA soa, uint32_t* d_n, const EmitBase* emitter, const EmitBase* parent, EmitBase* result, uint32_t n,
                const oon_emitter_cfg& cfg, const float2* nozzles, unsigned long long seed, int kind);
int launch_emit_apply(cudaStream_t st, BodySoA soa, uint32_t emitter_local, const EmitBase* result);

// slot order: identity, or Hilbert-curve order of the positions (FAST all-pairs mode)
size_t order_temp_bytes(uint32_t slots);
int launch_order_identity(cudaStream_t st, uint32_t* inv, uint32_t slots);
int launch_order_morton(cudaStream_t st, BodySoA soa, const uint32_t* d_n, uint32_t slots, uint32_t cblock, float predict_dt,
                        const float2* last_kick, uint32_t* inv, void* temp, size_t temp_bytes, unsigned long long* keys2 /* 2*slots */, uint32_t* vals_in /* slots */,
                        uint32_t* stats /* 8 x 16 words */);

// The same order in one launch, one CTA per canonical block, when a block fits one CTA (order_fits_one_cta).
bool order_fits_one_cta(uint32_t cblock);
int launch_order_block(cudaStream_t st, BodySoA soa, const uint32_t* d_n, uint32_t slots, uint32_t cblock, float predict_dt,
                       const float2* last_kick, uint32_t* inv);

// update_intrinsic + pack the shard's slice of the exchange buffer (one thread per slot, one block per tile).
// body_base = global index of local body 0.  Slots >= n are written dead.
// slot_of[body] = slot it was packed into (read by a later launch_integrate that packs in a NEW order).
int launch_intrinsic_pack(cudaStream_t st, BodySoA soa, const uint32_t* inv, uint32_t* slot_of, const uint32_t* d_n, uint32_t slots,
                          uint32_t body_base, const PackOut& out, StepParams sp, bool do_intrinsic, unsigned long long* counters);

// How the source tiles of the FAST all-pairs pass are grouped (canonical: a function of the world size only, it fixes
// the rounding) and how the groups are dealt to CTAs (launch geometry: free, the sums across groups are exact).
struct InteractGeometry {
    uint32_t group_tiles;    // canonical: tiles [0, small_begin) are summed in FP32 in groups of this many tiles ...
    uint32_t small_begin;    // canonical: ... tiles [small_begin, ntiles) one by one (a multiple of group_tiles; ntiles if none)
    uint32_t gpc;            // geometry: grid rows [0, y_big) walk gpc groups,
    uint32_t y_big;
    uint32_t y_mid;          // geometry: the next y_mid rows one group (y_big * gpc + y_mid == groups before small_begin),
    uint32_t spc;            // geometry: the rows after them spc single tiles
};

// The all-pairs pass.  tiles: whole gathered buffer (ntiles tiles).  Targets = local slots [0, slots) living at
// global slot target_base.  FAST: sources are cut in canonical groups and dealt to CTAs as `geo` says; group sums go to
// acc [2 components][FX_LIMBS][slots] (64-bit integer atomics) and touch visits to touch[slots].
// EXACT: partial[slots] = {new v.x, new v.y, touch visits, -}.
int launch_interact(cudaStream_t st, const float* tiles, uint32_t ntiles, const InteractGeometry& geo,
                    uint32_t target_base, uint32_t slots, const uint32_t* d_n_local, BodySoA soa, uint32_t body_base,
                    StepParams sp, const ExchangeView& xv, unsigned long long* acc, uint32_t* touch, int* fx_scale, float4* partial,
                    unsigned long long* counters, CaptureBuf cap, uint32_t dbg_fraction = 0, uint32_t dbg_part = 0,
                    unsigned long long* trace = nullptr /* tuning aid: 3 words per CTA */, uint32_t* trace_ctas = nullptr,
                    uint32_t* work_ctr = nullptr /* FAST: {next item, CTAs finished}, both 0 between launches: persistent CTAs draw their work from it */);
// Peer-to-peer exchange: hold the stream until every peer's slice of epoch xv.epoch has landed (for the kernels that
// read the exchange buffer without waiting themselves).
int launch_wait_peers(cudaStream_t st, const ExchangeView& xv);

// interact_n2n == false: only body 0 is a source (and, in half-matrix mode, the target of everyone).  Identity order.
int launch_interact_player_only(cudaStream_t st, const float* tiles, uint32_t ntiles, uint32_t target_base,
                                uint32_t slots, const uint32_t* d_n_local, BodySoA soa,
                                StepParams sp, float4* partial, float4* scratch, unsigned long long* counters, CaptureBuf cap);

// update_global (+ touch heat + kick application), optionally fused with the next step's
// update_intrinsic + pack.  kick_out (optional, body order) receives the velocity change of the pairwise pass.
// FAST all-pairs results come from acc/touch (normalised, converted and zeroed here; *fx_scale = the exponent B the
// interact kernel used), EXACT / player-only results from partial[slots].
int launch_integrate(cudaStream_t st, BodySoA soa, const uint32_t* inv, const uint32_t* d_n, uint32_t slots, uint32_t body_base,
                     const float4* partial, unsigned long long* acc, uint32_t* touch, const int* fx_scale,
                     StepParams sp, bool apply_pairwise, bool do_global,
                     bool fuse_next_intrinsic, const PackOut& out, float2* kick_out, unsigned long long* counters,
                     uint32_t* slot_of, bool reorder /* inv changed since the pass that filled acc: a body's sums sit at slot_of[body] */);

// The caller's expired-body sweep (OON.cpp:1089-1095) as a stable device compaction from `src` into `dst`.
// d_n is updated in place.  removed_idx (optional, capacity removed_cap) gets the indices at removal time.
size_t sweep_temp_bytes(uint32_t slots);
// Two phases so that a sharded world can all-gather the ranks' run bookkeeping in between (a run of expired bodies may
// cross the boundary between two ranks' index blocks): scan = flags + "last non-expired index" (+ this rank's triple
// {bodies, expired run the block ends with, whole block is one run} into my_runinfo); compact = keep flags from the run
// parity (runinfo = all ranks' triples, or nullptr), new indices, stable scatter, new count.  first_tested = first local
// index the sweep looks at (player + 1 on the rank that owns the player, 0 on the ranks after it).
int launch_sweep_scan(cudaStream_t st, BodySoA src, const uint32_t* d_n, uint32_t slots, uint32_t first_tested,
                      void* temp, size_t temp_bytes, int* scratch_i32 /* 3*slots ints */, int* my_runinfo);
int launch_sweep_compact(cudaStream_t st, BodySoA src, BodySoA dst, uint32_t* d_n, uint32_t slots, const int* runinfo, int rank,
                         void* temp, size_t temp_bytes, int* scratch_i32, uint32_t* removed_idx, uint32_t removed_cap,
                         uint32_t* d_removed_count, unsigned long long* counters, uint32_t* minfo /* like launch_unpack_f32, for the survivors */);

// render read set: {x, y, r} per body + the colour word
int launch_render_pack(cudaStream_t st, BodySoA soa, const uint32_t* d_n, uint32_t slots, float* xyr, uint32_t* colors);

// self-test of the EXACT kernel's written-out IEEE sqrt / division fast paths against the intrinsics (n random operand sets)
int run_ieee_selftest(cudaStream_t st, unsigned long long n, unsigned long long seed, unsigned long long* d_mismatches);

// microbenchmarks
float measure_fp32_peak_tflops(cudaStream_t st);
float measure_copy_gbs(cudaStream_t st);

}  // namespace oon
