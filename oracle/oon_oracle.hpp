// oon_oracle.hpp — CPU ORACLE for the Out_of_Nothing per-frame world update.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link, import or execute anything under oracle/.  The product path
// (out_of_nothing_b200/csrc + host) never calls it and has no CPU fallback.
//
// What this is: a single-threaded restatement of the reference's model update,
// templated on the number type (float: the type every shipped fixture was written
// with; double: the reference's current Cfg.hpp:6 setting), keeping the reference's
// data layout (std::vector<std::shared_ptr<Entity>>, half-matrix pair loop) so that
// it is also an honest CPU baseline.  It follows, function by function:
//   /root/reference/src/app/model/World.cpp:42-50    Properties::defaults
//   /root/reference/src/app/model/World.cpp:55-78    add_body / remove_body
//   /root/reference/src/app/model/World.cpp:103-188  update_intrinsic
//   /root/reference/src/app/model/World.cpp:192-438  update_pairwise
//   /root/reference/src/app/model/World.cpp:442-463  update_global
//   /root/reference/src/app/model/World.cpp:475-482  is_colliding
//   /root/reference/src/app/model/World.cpp:532-548  touch_hook
//   /root/reference/src/app/model/Entity.hpp:25-98, Entity.cpp:14-26
//   /root/reference/src/app/OON.cpp:1083-1095        the caller's expired-body sweep
//   /root/reference/src/app/model/Emitter.hpp:24-37, Emitter.cpp:22-92   Emitter::Config, emit_particles
//   /root/reference/src/app/OON.cpp:778-803, 842-861, 984-1029           add_random_body_near, spawn, chemtrail_burst
//   (the reference draws from the C library's unseeded rand(); here every draw is oon_rand(seed, k), a
//   counter-based stand-in with the MSVC RAND_MAX the reference ships with — see EmitterConfig below)
//
// Why a restatement: the reference cannot be compiled here.  Every translation unit
// includes headers of the author's separate engine (O2N 0.98: Szim/..., sz/..., Tracy),
// reached through a dangling Windows symlink (extern/DEPS.cfg:6-7).  The arithmetic
// that lives there (Math::Vector2 operator-, length(), normalized(), operator*;
// Phys::radius_from_mass_and_density; Phys::T_to_RGB_and_BV; DeltaT) is restated
// from its call sites; the choices made are listed below and were each checked
// against the reference's golden pair KAT-1 (tests/golden/kat1_{start,end}.state,
// i.e. test/regression/_baseline-2024-09-18/1000_bodies-START.state ->
// test/regression/tc-smoke/REFERENCE-END.state, 1010 bodies, 20 steps).
//
// PARITY PIN (what KAT-1 proves for Oracle<float>, see tests/test_oracle_golden.py):
//   * T bit-exact on 1010/1010 bodies  => every collision and touch decision of all
//     20 steps (3590 colliding pair-steps, 322 touches) and the cooling law match;
//   * r, mass, density, lifetime, flags unchanged, body count unchanged;
//   * positions: ~64 % bit-exact, max rel. error 1.6e-4; velocities: median 1 ulp.
//   * radius formula: r = powf((m/rho)/float(4/3*pi), float(1/3)) reproduces the stored
//     radius of all 1010 bodies bit-exactly.
//   The last-ulp rounding sequence of the 2024-09 binary's Vector2 is NOT recoverable
//   (64 variants tried: reciprocal vs division normalise, hypot, double internals,
//   double dt, a*dt first, (G*m)/d — none reaches 100 % on p/v), so p/v parity is
//   pinned to ~1 ulp per operation, not to the bit.  `color` (Phys::T_to_RGB_and_BV is
//   absent) and the every-100th-cooling recalc() are excluded: "parity unpinned" for
//   those two fields only.  The repulsion (`repulsion_stiffness != 0`) path has no
//   fixture: its mixed float/double evaluation order is restated literally from
//   World.cpp:396-416 but is unpinned.
//
// Choices for the absent engine pieces (each is the most literal reading):
//   length()      = sqrt(x*x + y*y)           (no hypot, no FMA: build with -ffp-contract=off)
//   normalized()  = {x / length(), y / length()}
//   V2 * scalar   = {x*s, y*s} in Num;  `n * a * dt` associates left to right
//   DeltaT        = float (hook signatures use `float dt`, World.hpp:139-140)
//   Phys::G       = 6.6743e-11 (header of kat3_end.state, written from the runtime default)
//   DENSITY_ROCK  = 2000 (Entity default density DENSITY_ROCK/2 = 1000 in every fixture)
//   unset<float>  = 2e31 (thruster fields of every non-player body in the fixtures)
//   collide_hook  = no-op returning false (World.cpp:327-349 has only comments in its body)
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace oon_oracle {

using DeltaT = float;

// --- minimal stand-in for Szim::Math::Vector2<T> (only what World.cpp uses) ---------------
template <class T>
struct V2 {
    T x{0}, y{0};
    V2 operator-(const V2& o) const { return {x - o.x, y - o.y}; }
    V2 operator*(T s) const { return {x * s, y * s}; }
    V2 operator/(T s) const { return {x / s, y / s}; }
    V2& operator+=(const V2& o) { x += o.x; y += o.y; return *this; }
    V2& operator-=(const V2& o) { x -= o.x; y -= o.y; return *this; }
    T length() const { return std::sqrt(x * x + y * y); }
    V2 normalized() const { T l = length(); return {x / l, y / l}; }
};

// --- stand-in for Szim::Model::Physics<T> ---------------------------------------------------
template <class T>
struct Phys {
    static constexpr T G = T(6.6743e-11);
    static constexpr T DENSITY_ROCK = T(2000);
    // r = cbrt(V / (4/3 pi)); evaluated the way that reproduces every stored radius of KAT-1:
    // float: powf(V / float(4/3 pi), float(1/3)), correctly rounded.
    static T radius_from_mass_and_density(T mass, T density) {
        const T four_thirds_pi = T(4.0 / 3.0 * 3.14159265358979323846);
        const T third = T(1) / T(3);
        return T(std::pow(double(mass / density / four_thirds_pi), double(third)));
    }
};

inline constexpr float THRUST_UNSET = 2e31f;  // Math::unset<float>()
inline constexpr float CFG_GLOBE_RADIUS = 50000000.0f;  // World.hpp:91
inline constexpr float EPS_COLLISION = CFG_GLOBE_RADIUS / 10;  // World.cpp:315

enum class GravityMode : unsigned { Off = 0, Hyperbolic, Realistic, Experimental, Default = Hyperbolic };
enum class CollisionMode : unsigned { Off = 0, Glide_Through = 1 };
enum class LoopMode : unsigned { Half_Matrix = 0, Full_Matrix = 1, Default = Half_Matrix };

// --- Entity (Entity.hpp:25-98) -----------------------------------------------------------------
template <class Num>
struct Entity {
    static constexpr float Unlimited = -1.f;
    struct { bool gravity_immunity = false; bool free_color = false; } superpower;
    Num lifetime = Unlimited;
    Num r = 0;
    Num density = Phys<Num>::DENSITY_ROCK / 2;
    V2<Num> p{}, v{};
    float T = 0;
    uint32_t color = 0;
    Num mass = 0;
    float thrust_up = THRUST_UNSET, thrust_down = THRUST_UNSET, thrust_left = THRUST_UNSET, thrust_right = THRUST_UNSET;

    // Entity.cpp:14-20.  Colour mapping is absent from the reference tree -> not restated.
    void recalc() { r = Phys<Num>::radius_from_mass_and_density(mass, density); }
    bool can_expire() const noexcept { return lifetime > 0; }
    void terminate() noexcept { lifetime = 0; }
    bool terminated() const noexcept { return lifetime == 0; }
    void on_event_terminated() { r *= 0.1f; }  // Entity.cpp:23-26
    bool has_thruster() const noexcept { return thrust_up != THRUST_UNSET; }
    bool is_player() const noexcept { return has_thruster(); }
};

template <class Num>
struct Properties {  // World.hpp:59-69, defaults World.cpp:42-50
    float friction = 0.03f;
    GravityMode gravity_mode = GravityMode::Default;
    Num gravity = Phys<Num>::G;
    bool interact_n2n = false;
    double repulsion_stiffness = 0;
    CollisionMode collision_mode = CollisionMode::Glide_Through;
};

// --- Emitter (Emitter.hpp:24-37) ----------------------------------------------------------------
// rand() replacement ("a seeded RNG replacing rand()"): draw k of a burst is
// oon_rand(seed, k) in [0, RAND_MAX], RAND_MAX = 32767 as in the MSVC runtime the reference ships with.
// Particle i of a burst uses draws 5i+0 (mass), 5i+1, 5i+2 (p.x, p.y), 5i+3, 5i+4 (v.x, v.y) — the order in
// which Emitter.cpp:49-71 evaluates its rand() calls — whether or not the particle is emitted.
inline constexpr int OON_RAND_MAX = 32767;
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline int oon_rand(uint64_t seed, uint64_t k) { return int(splitmix64(seed ^ (k * 0xD1342543DE82EF95ull)) >> 49); }

template <class Num>
struct EmitterConfig {
    V2<Num> eject_velocity{};
    V2<Num> eject_offset{};
    Num v_factor = Num(0.1f);
    Num offset_velo_factor = Num(0.2f);
    float particle_lifetime = -1;
    bool create_mass = true;
    Num particle_density = Phys<Num>::DENSITY_ROCK * Num(0.001f);
    V2<Num> position_divergence{Num(5), Num(5)};
    Num velocity_divergence = Num(1);
    Num particle_mass_min{};
    Num particle_mass_max{};
    uint32_t color = 0x706080;
};

struct PairEvent { uint32_t target, source; };

// Optional per-step trace, filled when World::trace != nullptr.
struct Trace {
    std::vector<PairEvent> collisions;  // every visit where is_colliding() was true
    std::vector<PairEvent> touches;     // the subset that reached touch_hook
    std::vector<uint32_t> expired;      // indices (at the time of removal) erased by the caller sweep
    std::vector<uint32_t> terminated;   // indices terminated by update_intrinsic
    uint64_t pairs_visited = 0;
    void clear() { collisions.clear(); touches.clear(); expired.clear(); terminated.clear(); pairs_visited = 0; }
};

// --- World ----------------------------------------------------------------------------------------
template <class Num>
class World {
public:
    using EntityT = Entity<Num>;
    Properties<Num> props;
    std::vector<std::shared_ptr<EntityT>> bodies;
    Trace* trace = nullptr;
    // recalc() on touch recomputes a radius that is a pure function of (mass, density), both
    // constant on this path; the golden END confirms r is unchanged on all touched bodies.
    // Kept as a switch only so tests can show it is idempotent.
    bool recalc_on_touch = false;

    LoopMode loop_mode() const { return loop_mode_; }
    void set_loop_mode(LoopMode m) { loop_mode_ = m; }

    size_t add_body(const EntityT& obj, bool recalc = true) {  // World.cpp:55-62
        bodies.push_back(std::make_shared<EntityT>(obj));
        if (recalc) bodies.back()->recalc();
        return bodies.size() - 1;
    }
    void remove_body(size_t ndx) { bodies.erase(bodies.begin() + ptrdiff_t(ndx)); }  // World.cpp:72-78

    // World.cpp:103-188
    void update_intrinsic(DeltaT dt) {
        if (dt == 0.f) return;
        for (size_t i = 0; i < bodies.size(); ++i) {
            auto& body = bodies[i];
            if (body->can_expire()) {
                body->lifetime -= static_cast<Num>(dt);
                if (body->lifetime <= 0) {
                    body->terminate();
                    body->on_event_terminated();
                    if (trace) trace->terminated.push_back(uint32_t(i));
                    continue;
                }
            }
            if (body->T > 0) body->T *= 0.996f;
            if (body->has_thruster()) {
                V2<Num> F_thr{Num((-body->thrust_left + body->thrust_right) * dt),
                              Num((body->thrust_up - body->thrust_down) * dt)};
                body->v += F_thr / body->mass;
            }
        }
    }

    // World.cpp:475-482
    bool is_colliding(const EntityT* a, const EntityT* b, Num distance) const {
        return props.collision_mode == CollisionMode::Off ? false : distance <= a->r + b->r;
    }

    // World.cpp:532-548 (audio is a front-end side effect; not part of the model state)
    void touch_hook(EntityT* a, EntityT* b) {
        a->T += 100;
        b->T += 100;
        if (recalc_on_touch) { a->recalc(); b->recalc(); }
    }

    // World.cpp:192-438
    void update_pairwise(DeltaT dt) {
        const size_t obj_cnt = bodies.size();
        const size_t n_src = props.interact_n2n ? obj_cnt : (obj_cnt ? 1 : 0);
        for (size_t s = 0; s < n_src; ++s) {
            if (bodies[s]->terminated()) continue;
            for (size_t t = (loop_mode_ == LoopMode::Full_Matrix ? 0 : s + 1); t < obj_cnt; ++t) {
                if (s == t) continue;
                EntityT* target = bodies[t].get();
                if (target->terminated()) continue;
                EntityT* source = bodies[s].get();
                if (trace) ++trace->pairs_visited;

                auto dist_vect = source->p - target->p;
                auto dist_normal = dist_vect.normalized();
                auto distance = dist_vect.length();

                Num g_factor;
                switch (props.gravity_mode) {
                case GravityMode::Hyperbolic: g_factor = props.gravity / distance; break;
                case GravityMode::Realistic:
                case GravityMode::Experimental: g_factor = props.gravity / (distance * distance); break;
                default: g_factor = 0; break;
                }

                if (is_colliding(target, source, distance)) {
                    if (trace) trace->collisions.push_back({uint32_t(t), uint32_t(s)});
                    if (std::abs(distance - (target->r + source->r)) < EPS_COLLISION) {
                        if (trace) trace->touches.push_back({uint32_t(t), uint32_t(s)});
                        touch_hook(target, source);
                    }
                    // collide_hook: no-op, glide through (no velocity change while overlapping)
                } else if (loop_mode_ == LoopMode::Full_Matrix) {
                    if (!target->superpower.gravity_immunity) {
                        Num a = g_factor * source->mass;
                        target->v += dist_normal * a * Num(dt);
                    }
                } else {
                    double sculptor_field_correction;
                    if (props.repulsion_stiffness == 0) {
                        sculptor_field_correction = 0;
                    } else {
                        sculptor_field_correction = g_factor * props.repulsion_stiffness
                                                    * (source->mass / distance) * (target->mass / distance);
                    }
                    if (!target->superpower.gravity_immunity) {
                        auto a = g_factor * source->mass - sculptor_field_correction;
                        target->v += dist_normal * Num(a) * Num(dt);
                    }
                    if (!source->superpower.gravity_immunity) {
                        auto a = -(g_factor * target->mass - sculptor_field_correction);
                        source->v += dist_normal * Num(a) * Num(dt);
                    }
                }
            }
        }
    }

    // World.cpp:442-463
    void update_global(DeltaT dt) {
        for (size_t n = bodies.size(), i = 0; i < n; ++i) {
            auto& body = bodies[i];
            auto friction_decel = body->v * Num(props.friction);
            body->v -= friction_decel * Num(dt);
            body->p += body->v * Num(dt);
        }
    }

    // The caller's clean-up after update_world(), OON.cpp:1089-1095 — including its quirk:
    // no index fix-up after erase, so the body sliding into slot i is not tested this frame.
    void sweep_expired(size_t player_ndx = 0) {
        for (size_t i = player_ndx + 1; i < bodies.size(); ++i) {
            auto& e = *bodies[i];
            if (e.lifetime != EntityT::Unlimited && e.lifetime <= 0) {
                if (trace) trace->expired.push_back(uint32_t(i));
                remove_body(i);
            }
        }
    }

    // Emitter::emit_particles, Emitter.cpp:22-92.  nozzles: n positions relative to the emitter, or nullptr.
    // Returns the number of particles appended.
    unsigned emit_particles(size_t emitter_id, unsigned n, const EmitterConfig<Num>& cfg, const V2<Num>* nozzles, uint64_t seed) {
        const Num RM = Num(OON_RAND_MAX);
        unsigned emitted = 0;
        // (the reference holds a reference to the emitter; bodies are shared_ptrs, so push_back does not move it)
        EntityT& emitter = *bodies[emitter_id];
        V2<Num> p_range{cfg.position_divergence.x * emitter.r, cfg.position_divergence.y * emitter.r};
        V2<Num> p_offset = cfg.eject_offset;
        {
            // emitter.v.normalized() * cfg.offset_velo_factor * emitter.r; the null vector normalises to itself
            // ("always checked implicitly", Emitter.cpp:41-43)
            const Num l = emitter.v.length();
            V2<Num> nv = l == 0 ? V2<Num>{} : V2<Num>{emitter.v.x / l, emitter.v.y / l};
            V2<Num> add = nv * cfg.offset_velo_factor * emitter.r;
            p_offset += add;
        }
        const Num v_range = emitter.r * cfg.velocity_divergence;
        for (unsigned i = 0; i < n; ++i) {
            const uint64_t k = uint64_t(i) * 5;
            const Num particle_mass = cfg.particle_mass_min + (cfg.particle_mass_max - cfg.particle_mass_min) * float(oon_rand(seed, k)) / RM;
            if (!cfg.create_mass && emitter.mass < particle_mass) continue;
            V2<Num> p{(Num(oon_rand(seed, k + 1)) * p_range.x) / RM - p_range.x / 2 + emitter.p.x + p_offset.x,
                      (Num(oon_rand(seed, k + 2)) * p_range.y) / RM - p_range.y / 2 + emitter.p.y + p_offset.y};
            if (nozzles) p += nozzles[i] * emitter.r;
            EntityT e;
            e.lifetime = cfg.particle_lifetime;
            e.density = cfg.particle_density;
            e.p = p;
            e.v = {(Num(oon_rand(seed, k + 3)) * v_range) / RM - v_range / 2 + emitter.v.x * cfg.v_factor + cfg.eject_velocity.x,
                   (Num(oon_rand(seed, k + 4)) * v_range) / RM - v_range / 2 + emitter.v.y * cfg.v_factor + cfg.eject_velocity.y};
            e.color = cfg.color;
            e.mass = particle_mass;
            add_body(e);
            ++emitted;
            if (!cfg.create_mass) emitter.mass -= particle_mass;
        }
        if (!cfg.create_mass) emitter.recalc();
        return emitted;
    }

    // OONApp::chemtrail_burst, OON.cpp:984-1029 (the appcfg values arrive through cfg: v_factor, offset_velo_factor =
    // chemtrail_offset_factor, particle_lifetime, create_mass, particle_density, velocity_divergence = chemtrail_divergence,
    // particle_mass_min/max = M_min/M_max).  6 draws per particle: mass, p.x, p.y, v.x, v.y, colour.
    unsigned chemtrail_burst(size_t emitter_id, unsigned n, const EmitterConfig<Num>& cfg, uint64_t seed) {
        const Num RM = Num(OON_RAND_MAX);
        unsigned emitted = 0;
        EntityT& emitter = *bodies[emitter_id];
        const Num p_range = emitter.r * 5;
        const Num v_range = Num(CFG_GLOBE_RADIUS) * cfg.velocity_divergence;
        for (unsigned i = 0; i++ < n;) {
            const uint64_t k = uint64_t(i - 1) * 6;
            const Num particle_mass = cfg.particle_mass_min + (cfg.particle_mass_max - cfg.particle_mass_min) * float(oon_rand(seed, k)) / RM;
            if (!cfg.create_mass && emitter.mass < particle_mass) continue;
            EntityT e;
            e.lifetime = cfg.particle_lifetime;
            e.density = cfg.particle_density;
            e.p = {(Num(oon_rand(seed, k + 1)) * p_range) / RM - p_range / 2 + emitter.p.x + emitter.v.x * cfg.offset_velo_factor,
                   (Num(oon_rand(seed, k + 2)) * p_range) / RM - p_range / 2 + emitter.p.y + emitter.v.y * cfg.offset_velo_factor};
            e.v = {(Num(oon_rand(seed, k + 3)) * v_range) / RM - v_range / 2 + emitter.v.x * cfg.v_factor,
                   (Num(oon_rand(seed, k + 4)) * v_range) / RM - v_range / 2 + emitter.v.y * cfg.v_factor};
            e.color = (uint32_t)(float)0xffffff * uint32_t(oon_rand(seed, k + 5));
            e.mass = particle_mass;
            add_body(e);
            ++emitted;
            if (!cfg.create_mass) emitter.mass -= particle_mass;
        }
        emitter.recalc();
        return emitted;
    }

    // OONApp::add_random_body_near, OON.cpp:778-803.  7 draws from k0: p.x, p.y, v.x, v.y, colour, colour, mass.
    size_t add_random_body_near(size_t base_id, uint64_t seed, uint64_t k0) {
        const Num RM = Num(OON_RAND_MAX);
        const Num p_range = Num(CFG_GLOBE_RADIUS * 30), v_range = Num(CFG_GLOBE_RADIUS * 10);
        const EntityT base = *bodies[base_id];
        const Num M_min = base.mass / 9, M_max = base.mass * 3;
        EntityT e;
        e.p = {(Num(oon_rand(seed, k0)) * p_range) / RM - p_range / 2 + base.p.x,
               (Num(oon_rand(seed, k0 + 1)) * p_range) / RM - p_range / 2 + base.p.y};
        e.v = {(Num(oon_rand(seed, k0 + 2)) * v_range) / RM - v_range / 2 + base.v.x * 0.05f,
               (Num(oon_rand(seed, k0 + 3)) * v_range) / RM - v_range / 2 + base.v.y * 0.05f};
        e.color = 0xffffff & (uint32_t(oon_rand(seed, k0 + 4)) * uint32_t(oon_rand(seed, k0 + 5)));
        e.mass = M_min + (M_max - M_min) * float(oon_rand(seed, k0 + 6)) / RM;
        return add_body(e);
    }

    // OONApp::spawn, OON.cpp:842-861: bodies near the PLAYER, temperature and velocity inherited from the parent.
    void spawn(size_t parent_id, unsigned n, uint64_t seed, size_t player_id = 0) {
        const EntityT parent = *bodies[parent_id];
        for (unsigned i = 0; i < n; ++i) {
            const size_t ndx = add_random_body_near(player_id, seed, uint64_t(i) * 7);
            EntityT& newborn = *bodies[ndx];
            newborn.lifetime = EntityT::Unlimited;
            newborn.T = parent.T;
            newborn.v = parent.v;
        }
    }

    // One frame of OONApp::updates_for_next_frame's model part (OON.cpp:1085-1095).
    void step(DeltaT dt, bool sweep = true) {
        update_intrinsic(dt);
        update_pairwise(dt);
        update_global(dt);
        if (sweep) sweep_expired(0);
    }

private:
    LoopMode loop_mode_ = LoopMode::Default;
};

}  // namespace oon_oracle
