// oon_world.hpp — C++ host mirror of the reference's model interface over the C ABI (include/oon_gpu.h).
//
// Same names, argument meaning and error behaviour as
//   OON::Model::World       /root/reference/src/app/model/World.hpp:82-158   (props, add_body, remove_body,
//                                                                             update_intrinsic/_pairwise/_global,
//                                                                             loop_mode/set_loop_mode, save/load)
//   OON::Model::Entity      /root/reference/src/app/model/Entity.hpp:25-98   (float-build layout, recalc, terminated...)
//   Properties, GravityMode, CollisionMode   World.hpp:42-69
// plus the caller-side frame of OONApp::updates_for_next_frame (OON.cpp:1085-1095) as World::update(dt).
//
// All physics runs on the device through liboon_gpu.so; this header only keeps the host AoS and the device
// table coherent: the host copy is authoritative after any host-side mutation (entity(i), add_body, remove_body,
// load), the device copy after any update_* call.  There is no CPU implementation of the update here.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iosfwd>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/oon_gpu.h"

namespace OON::Model {

static constexpr char const* VERSION = "0.1.5";   // World.hpp:36

enum class GravityMode : unsigned { Off = 0, Hyperbolic, Realistic, Experimental, Default = Hyperbolic };
enum class CollisionMode : unsigned { Off = 0, Glide_Through = 1 };
enum class LoopMode : unsigned { Half_Matrix = 0, Full_Matrix = 1, Default = Half_Matrix };

namespace Phys {
inline constexpr float G = 6.6743e-11f;
inline constexpr float DENSITY_ROCK = 2000.f;
// Entity::recalc()'s radius (Entity.cpp:17); the evaluation order reproduces the stored radius of every body of the
// reference's KAT-1 and KAT-3 fixtures bit-exactly (tests/test_oracle_golden.py).
inline float radius_from_mass_and_density(float mass, float density) {
    const float four_thirds_pi = float(4.0 / 3.0 * 3.14159265358979323846);
    return float(std::pow(double(mass / density / four_thirds_pi), double(1.0f / 3.0f)));
}
}  // namespace Phys

inline constexpr float THRUST_UNSET = 2e31f;   // Math::unset<float>()

// Byte-identical to the reference's float-build Entity and to oon_body_f32.
struct Entity {
    static constexpr float Unlimited = -1.f;
    struct { bool gravity_immunity = false; bool free_color = false; } superpower;
    uint8_t _pad[2] = {0, 0};
    float lifetime = Unlimited;
    float r = 0;
    float density = Phys::DENSITY_ROCK / 2;
    struct { float x = 0, y = 0; } p, v;
    float T = 0;
    uint32_t color = 0;
    float mass = 0;
    float thrust_up = THRUST_UNSET, thrust_down = THRUST_UNSET, thrust_left = THRUST_UNSET, thrust_right = THRUST_UNSET;

    void recalc() { r = Phys::radius_from_mass_and_density(mass, density); }   // colour mapping absent from the reference tree
    bool can_expire() const noexcept { return lifetime > 0; }
    void terminate() noexcept { lifetime = 0; }
    bool terminated() const noexcept { return lifetime == 0; }
    bool has_thruster() const noexcept { return thrust_up != THRUST_UNSET; }
    void add_thrusters() noexcept { thrust_up = thrust_down = thrust_left = thrust_right = 0; }
    bool is_player() const noexcept { return has_thruster(); }
};
static_assert(sizeof(Entity) == sizeof(oon_body_f32) && sizeof(Entity) == 60, "Entity must keep the reference's 60-byte layout");

struct Properties {   // World.hpp:59-69, defaults World.cpp:42-50
    float friction = 0.03f;
    GravityMode gravity_mode = GravityMode::Default;
    float gravity = Phys::G;
    bool interact_n2n = false;
    double repulsion_stiffness = 0;
    CollisionMode collision_mode = CollisionMode::Glide_Through;
};

using EntityID = size_t;
using DeltaT = float;

class World {
public:
    static constexpr float CFG_GLOBE_RADIUS = 50000000.0f;   // World.hpp:91

    Properties props;

    explicit World(int device = 0) {
        if (int rc = oon_create(device, &h_)) throw std::runtime_error(std::string("oon_create: ") + oon_last_error(nullptr) + " (" + std::to_string(rc) + ")");
    }
    // One world over several GPUs of this process (the reference is one process with a single mutator, main.cpp:88-91):
    // same interface, the shards are the library's business.
    explicit World(std::vector<int> const& devices) {
        if (int rc = oon_create_multi(devices.data(), int(devices.size()), &h_))
            throw std::runtime_error(std::string("oon_create_multi: ") + oon_last_error(nullptr) + " (" + std::to_string(rc) + ")");
    }
    ~World() { if (h_) oon_destroy(h_); }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    // ---- setup (World.hpp:115-121, World.cpp:55-90) --------------------------------------------
    // init(): the reference's World::init() only calls back into the app (`app.init_world_hook()`, World.cpp:82-90); the
    // mirror has no app object, so the host installs the hook it wants run (oon_headless: OONApp::init_world_hook's body).
    std::function<void(World&)> init_world_hook;
    void init() { if (init_world_hook) init_world_hook(*this); }

    EntityID add_body(Entity const& obj) {
        pull();
        bodies_.push_back(obj);
        bodies_.back().recalc();
        host_dirty_ = true;
        return bodies_.size() - 1;
    }
    EntityID add_body(Entity&& obj) {                      // World.cpp:64-70: recalc the throw-away original, then append
        obj.recalc();
        pull();
        bodies_.push_back(obj);
        host_dirty_ = true;
        return bodies_.size() - 1;
    }
    // World.cpp:468-482.  The two-argument overload is the reference's stub (always false); the decision the update
    // itself takes is the three-argument one — here for host code that wants to ask the same question (GUI picking, tests);
    // the device evaluates it per visited pair with the identical IEEE sequence (exact_decide, csrc/oon_kernels.cu).
    bool is_colliding(Entity const* /*obj1*/, Entity const* /*obj2*/) const { return false; }
    bool is_colliding(Entity const* obj1, Entity const* obj2, float distance) const {
        return props.collision_mode == CollisionMode::Off ? false : distance <= obj1->r + obj2->r;
    }
    void remove_body(EntityID ndx) {
        pull();
        if (ndx >= bodies_.size()) throw std::out_of_range("remove_body");
        bodies_.erase(bodies_.begin() + std::ptrdiff_t(ndx));
        host_dirty_ = true;
    }
    size_t entity_count() { pull(); return bodies_.size(); }
    Entity const& const_entity(EntityID i) { pull(); return bodies_.at(i); }
    Entity& entity(EntityID i) { pull(); host_dirty_ = true; return bodies_.at(i); }   // mutable access: host becomes authoritative
    std::vector<Entity> const& bodies() { pull(); return bodies_; }

    LoopMode loop_mode() const { return loop_mode_; }
    void set_loop_mode(LoopMode m) { loop_mode_ = m; }
    void set_math_mode(int oon_math_mode) { check(oon_set_math_mode(h_, oon_math_mode)); }

    // ---- modelling (World.hpp:126-128) ---------------------------------------------------------
    void update_intrinsic(DeltaT dt) { push(); check(oon_update_intrinsic(h_, dt)); device_ahead_ = true; }
    void update_pairwise(DeltaT dt) { push(); check(oon_update_pairwise(h_, dt)); device_ahead_ = true; }
    void update_global(DeltaT dt) { push(); check(oon_update_global(h_, dt)); device_ahead_ = true; }

    // One model frame of OONApp::updates_for_next_frame (OON.cpp:1085-1095): the three updates + the expired-body sweep.
    void update(DeltaT dt, unsigned frames = 1) { push(); check(oon_step(h_, dt, frames)); device_ahead_ = true; }

    // Emitter::emit_particles (Emitter.cpp:22-92) on the device: no upload, the host copy is refreshed lazily.
    // `seed` replaces the reference's unseeded rand() stream (include/oon_gpu.h, oon_emit_particles).
    unsigned emit_particles(EntityID emitter, unsigned n, oon_emitter_cfg const& cfg, uint64_t seed, float const* nozzles_xy = nullptr) {
        push();
        uint32_t emitted = 0;
        check(oon_emit_particles(h_, emitter, n, &cfg, nozzles_xy, seed, &emitted));
        if (emitted || !cfg.create_mass) device_ahead_ = true;
        return emitted;
    }
    static oon_emitter_cfg emitter_defaults() { oon_emitter_cfg c{}; oon_emitter_cfg_default(&c); return c; }   // Emitter::Config{}

    // What the renderer reads every frame: x, y, r per body (OONMainDisplay_sfml.cpp:181-211), without a full download.
    std::vector<float> const& render_xyr() {
        push();
        uint64_t n = 0;
        check(oon_body_count(h_, &n));
        xyr_.resize(size_t(n) * 3);
        check(oon_download_render(h_, xyr_.data(), n, &n));
        xyr_.resize(size_t(n) * 3);
        return xyr_;
    }
    // The same read-back beside the simulation: begin() after update(), draw the previous frame, end().
    // `pinned` = cudaHostAlloc'ed / cudaHostRegister'ed memory for 3 floats per body, valid until render_end().
    void render_begin(float* pinned, size_t capacity_bodies) { push(); check(oon_download_render_begin(h_, pinned, capacity_bodies)); }
    size_t render_end() { uint64_t n = 0; check(oon_download_render_end(h_, &n)); return size_t(n); }

    oon_events drain_events() { oon_events ev{}; check(oon_drain_events(h_, &ev)); return ev; }

    // ---- persistence (World_SaveLoad.cpp) --------------------------------------------------------
    bool save(std::ostream& out, const char* version = nullptr);
    static bool load(std::istream& in, World* result);

    int device_count() const { int n = 1; oon_device_count(h_, &n); return n; }
    oon_world* handle() { return h_; }

private:
    void check(int rc) {
        if (rc) throw std::runtime_error(std::string("oon: ") + oon_last_error(h_) + " (" + std::to_string(rc) + ")");
    }
    void push_props() {
        oon_props p{};
        p.friction = props.friction; p.gravity_mode = unsigned(props.gravity_mode); p.gravity = props.gravity;
        p.interact_n2n = props.interact_n2n; p.collision_mode = unsigned(props.collision_mode);
        p.repulsion_stiffness = props.repulsion_stiffness; p.loop_mode = unsigned(loop_mode_);
        check(oon_set_props(h_, &p));
    }
    void push() {   // host -> device when the host copy was touched; knobs always (they are tiny)
        push_props();
        if (host_dirty_) {
            check(oon_upload(h_, reinterpret_cast<const oon_body_f32*>(bodies_.data()), bodies_.size()));
            host_dirty_ = false;
            device_ahead_ = false;
        }
    }
    void pull() {   // device -> host when frames ran since the last download
        if (!device_ahead_) return;
        uint64_t n = 0;
        check(oon_body_count(h_, &n));
        bodies_.resize(size_t(n));
        check(oon_download(h_, reinterpret_cast<oon_body_f32*>(bodies_.data()), n, &n));
        bodies_.resize(size_t(n));
        device_ahead_ = false;
    }

    oon_world* h_ = nullptr;
    std::vector<Entity> bodies_;
    std::vector<float> xyr_;
    LoopMode loop_mode_ = LoopMode::Default;
    bool host_dirty_ = false, device_ahead_ = false;
};

}  // namespace OON::Model
