// oon_oracle_capi.cpp — C entry points over the CPU oracle, for ctypes (tests, bench cpu_baseline).
// TEST INFRASTRUCTURE ONLY — see the header of oon_oracle.hpp.  Nothing under
// out_of_nothing_b200/ links or loads this library.
#include "oon_oracle.hpp"

#include <chrono>
#include <cstring>
#include <type_traits>
#include <variant>

using namespace oon_oracle;

namespace {

// Same 60-byte record as the reference's float-build Entity and as
// oon_body_f32 in include/oon_gpu.h.  Re-declared here on purpose: the oracle must not
// depend on product headers.
#pragma pack(push, 1)
struct BodyF32 {
    uint8_t gravity_immunity, free_color, pad[2];
    float lifetime, r, density, px, py, vx, vy, T;
    uint32_t color;
    float mass, thrust_up, thrust_down, thrust_left, thrust_right;
};
#pragma pack(pop)
static_assert(sizeof(BodyF32) == 60, "record layout");

struct Handle {
    bool is_double;
    World<float> wf;
    World<double> wd;
    Trace trace;
    bool tracing = false;
};

template <class Num>
void set_props(World<Num>& w, float friction, unsigned gm, double gravity, int n2n, double rep, unsigned cm, unsigned lm) {
    w.props.friction = friction;
    w.props.gravity_mode = GravityMode(gm);
    w.props.gravity = Num(gravity);
    w.props.interact_n2n = n2n != 0;
    w.props.repulsion_stiffness = rep;
    w.props.collision_mode = CollisionMode(cm);
    w.set_loop_mode(LoopMode(lm));
}

template <class Num>
void set_bodies(World<Num>& w, const BodyF32* aos, uint64_t n, int recalc) {
    w.bodies.clear();
    w.bodies.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
        const BodyF32& b = aos[i];
        Entity<Num> e;
        e.superpower.gravity_immunity = b.gravity_immunity != 0;
        e.superpower.free_color = b.free_color != 0;
        e.lifetime = b.lifetime; e.r = b.r; e.density = b.density;
        e.p = {Num(b.px), Num(b.py)}; e.v = {Num(b.vx), Num(b.vy)};
        e.T = b.T; e.color = b.color; e.mass = b.mass;
        e.thrust_up = b.thrust_up; e.thrust_down = b.thrust_down;
        e.thrust_left = b.thrust_left; e.thrust_right = b.thrust_right;
        w.add_body(e, recalc != 0);
    }
}

template <class Num>
void get_bodies(const World<Num>& w, BodyF32* out) {
    for (size_t i = 0; i < w.bodies.size(); ++i) {
        const Entity<Num>& e = *w.bodies[i];
        BodyF32& b = out[i];
        std::memset(&b, 0, sizeof b);
        b.gravity_immunity = e.superpower.gravity_immunity; b.free_color = e.superpower.free_color;
        b.lifetime = float(e.lifetime); b.r = float(e.r); b.density = float(e.density);
        b.px = float(e.p.x); b.py = float(e.p.y); b.vx = float(e.v.x); b.vy = float(e.v.y);
        b.T = e.T; b.color = e.color; b.mass = float(e.mass);
        b.thrust_up = e.thrust_up; b.thrust_down = e.thrust_down;
        b.thrust_left = e.thrust_left; b.thrust_right = e.thrust_right;
    }
}

template <class Num>
void get_pv64(const World<Num>& w, double* out) {  // [n][4] = px,py,vx,vy at full oracle precision
    for (size_t i = 0; i < w.bodies.size(); ++i) {
        const Entity<Num>& e = *w.bodies[i];
        out[4 * i + 0] = double(e.p.x); out[4 * i + 1] = double(e.p.y);
        out[4 * i + 2] = double(e.v.x); out[4 * i + 3] = double(e.v.y);
    }
}

template <class Num>
void phase(World<Num>& w, int which, float dt) {
    switch (which) {
    case 0: w.update_intrinsic(dt); break;
    case 1: w.update_pairwise(dt); break;
    case 2: w.update_global(dt); break;
    case 3: w.sweep_expired(0); break;
    }
}

}  // namespace

#define DISPATCH(h, expr_f, expr_d) do { if ((h)->is_double) { auto& w = (h)->wd; expr_d; } else { auto& w = (h)->wf; expr_f; } } while (0)

extern "C" {

void* oracle_create(int use_double) {
    auto* h = new Handle;
    h->is_double = use_double != 0;
    return h;
}
void oracle_destroy(void* p) { delete static_cast<Handle*>(p); }

void oracle_set_props(void* p, float friction, unsigned gravity_mode, double gravity, int interact_n2n,
                      double repulsion_stiffness, unsigned collision_mode, unsigned loop_mode) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, set_props(w, friction, gravity_mode, gravity, interact_n2n, repulsion_stiffness, collision_mode, loop_mode),
                set_props(w, friction, gravity_mode, gravity, interact_n2n, repulsion_stiffness, collision_mode, loop_mode));
}

void oracle_set_bodies_f32(void* p, const void* aos, uint64_t n, int recalc) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, set_bodies(w, static_cast<const BodyF32*>(aos), n, recalc),
                set_bodies(w, static_cast<const BodyF32*>(aos), n, recalc));
}

uint64_t oracle_body_count(void* p) {
    auto* h = static_cast<Handle*>(p);
    return h->is_double ? h->wd.bodies.size() : h->wf.bodies.size();
}

void oracle_get_bodies_f32(void* p, void* out) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, get_bodies(w, static_cast<BodyF32*>(out)), get_bodies(w, static_cast<BodyF32*>(out)));
}

void oracle_get_pv_f64(void* p, double* out) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, get_pv64(w, out), get_pv64(w, out));
}

void oracle_trace(void* p, int on) {
    auto* h = static_cast<Handle*>(p);
    h->tracing = on != 0;
    h->trace.clear();
    h->wf.trace = h->tracing ? &h->trace : nullptr;
    h->wd.trace = h->tracing ? &h->trace : nullptr;
}
void oracle_trace_clear(void* p) { static_cast<Handle*>(p)->trace.clear(); }

// counts[0..4] = collisions, touches, expired, terminated, pairs_visited
void oracle_trace_counts(void* p, uint64_t* counts) {
    auto& t = static_cast<Handle*>(p)->trace;
    counts[0] = t.collisions.size(); counts[1] = t.touches.size(); counts[2] = t.expired.size();
    counts[3] = t.terminated.size(); counts[4] = t.pairs_visited;
}
// which: 0 collisions (pairs: target,source), 1 touches (pairs), 2 expired (singles), 3 terminated (singles)
void oracle_trace_fetch(void* p, int which, uint32_t* out) {
    auto& t = static_cast<Handle*>(p)->trace;
    switch (which) {
    case 0: std::memcpy(out, t.collisions.data(), t.collisions.size() * sizeof(PairEvent)); break;
    case 1: std::memcpy(out, t.touches.data(), t.touches.size() * sizeof(PairEvent)); break;
    case 2: std::memcpy(out, t.expired.data(), t.expired.size() * sizeof(uint32_t)); break;
    case 3: std::memcpy(out, t.terminated.data(), t.terminated.size() * sizeof(uint32_t)); break;
    }
}

// which: 0 update_intrinsic, 1 update_pairwise, 2 update_global, 3 caller sweep
void oracle_phase(void* p, int which, float dt) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, phase(w, which, dt), phase(w, which, dt));
}

// the caller's sweep with the app's player index (OON.cpp:1089: `for (i = player_entity_ndx() + 1; ...`)
void oracle_sweep(void* p, uint64_t player_ndx) {
    auto* h = static_cast<Handle*>(p);
    DISPATCH(h, w.sweep_expired(size_t(player_ndx)), w.sweep_expired(size_t(player_ndx)));
}

void oracle_step(void* p, float dt, uint32_t nsteps, int sweep) {
    auto* h = static_cast<Handle*>(p);
    for (uint32_t k = 0; k < nsteps; ++k) DISPATCH(h, w.step(dt, sweep != 0), w.step(dt, sweep != 0));
}

// Wall-clock seconds for nsteps frames (steady_clock around the loop, no I/O) — the CPU baseline.
double oracle_time_steps(void* p, float dt, uint32_t nsteps, int sweep) {
    auto t0 = std::chrono::steady_clock::now();
    oracle_step(p, dt, nsteps, sweep);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Emitter::emit_particles (Emitter.cpp:22-92).  cfg = 15 floats in the order of oon_emitter_cfg (include/oon_gpu.h):
// eject_velocity.xy, eject_offset.xy, v_factor, offset_velo_factor, particle_lifetime, create_mass, particle_density,
// position_divergence.xy, velocity_divergence, particle_mass_min, particle_mass_max, color (as uint32 bits).
// kind: 0 Emitter::emit_particles, 1 OONApp::chemtrail_burst (same 15-float configuration)
uint32_t oracle_emit_kind(void* p, int kind, uint64_t emitter_id, uint32_t n, const float* c, const float* nozzles, uint64_t seed);
uint32_t oracle_emit_particles(void* p, uint64_t emitter_id, uint32_t n, const float* c, const float* nozzles, uint64_t seed) {
    return oracle_emit_kind(p, 0, emitter_id, n, c, nozzles, seed);
}
uint32_t oracle_chemtrail_burst(void* p, uint64_t emitter_id, uint32_t n, const float* c, uint64_t seed) {
    return oracle_emit_kind(p, 1, emitter_id, n, c, nullptr, seed);
}
void oracle_spawn(void* p, uint64_t parent_id, uint64_t player_id, uint32_t n, uint64_t seed) {
    auto* h = static_cast<Handle*>(p);
    if (h->is_double) { if (parent_id < h->wd.bodies.size() && player_id < h->wd.bodies.size()) h->wd.spawn(size_t(parent_id), n, seed, size_t(player_id)); }
    else { if (parent_id < h->wf.bodies.size() && player_id < h->wf.bodies.size()) h->wf.spawn(size_t(parent_id), n, seed, size_t(player_id)); }
}
uint32_t oracle_emit_kind(void* p, int kind, uint64_t emitter_id, uint32_t n, const float* c, const float* nozzles, uint64_t seed) {
    auto* h = static_cast<Handle*>(p);
    auto run = [&](auto& w) -> uint32_t {
        using Num = std::decay_t<decltype(w.bodies[0]->r)>;
        if (emitter_id >= w.bodies.size()) return 0;
        EmitterConfig<Num> cfg;
        cfg.eject_velocity = {Num(c[0]), Num(c[1])}; cfg.eject_offset = {Num(c[2]), Num(c[3])};
        cfg.v_factor = Num(c[4]); cfg.offset_velo_factor = Num(c[5]); cfg.particle_lifetime = c[6]; cfg.create_mass = c[7] != 0.f;
        cfg.particle_density = Num(c[8]); cfg.position_divergence = {Num(c[9]), Num(c[10])}; cfg.velocity_divergence = Num(c[11]);
        cfg.particle_mass_min = Num(c[12]); cfg.particle_mass_max = Num(c[13]);
        uint32_t color; std::memcpy(&color, &c[14], 4); cfg.color = color;
        std::vector<V2<Num>> nz;
        if (nozzles) for (uint32_t i = 0; i < n; ++i) nz.push_back({Num(nozzles[2 * i]), Num(nozzles[2 * i + 1])});
        if (kind == 1) return w.chemtrail_burst(size_t(emitter_id), n, cfg, seed);
        return w.emit_particles(size_t(emitter_id), n, cfg, nozzles ? nz.data() : nullptr, seed);
    };
    return h->is_double ? run(h->wd) : run(h->wf);
}

int oracle_rand(uint64_t seed, uint64_t k) { return oon_rand(seed, k); }

float oracle_radius_from_mass_and_density_f32(float mass, float density) {
    return Phys<float>::radius_from_mass_and_density(mass, density);
}

}  // extern "C"
