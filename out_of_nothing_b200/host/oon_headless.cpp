// oon_headless.cpp — headless driver over the C++ host mirror (the new "headless benchmark driver" of the north star).
//
// Speaks the command line of the reference's regression scripts
//   test/regression/tc-smoke.cmd:30-43, test-headless.cmd:22-24, test-realg.cmd:1-4, fixrun.cmd:9
// i.e. --headless --cfg= --interact --friction= --fixed-dt= --fps-limit= --loop-cap= --exit-on-finish --session=
//      --session-save-as= --no-save-compressed --bodies= --g-mode=R|H|0 --log-level=   (unknown engine flags are ignored)
// and follows OONApp::init_world_hook (src/app/OON.cpp:241-292) for the initial world.
// Extras: --math=fast|exact, --device=N, --gpus=N (one world over N GPUs of this process: oon_create_multi), --loop-mode=half|full,
//         --convert-only (load + save, no device), --report.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "oon_snapshot.hpp"

using namespace OON::Model;

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string get(const std::string& k, const std::string& d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
};

Args parse(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        if (s.rfind("--", 0) != 0) continue;
        s = s.substr(2);
        const size_t eq = s.find('=');
        if (eq == std::string::npos) a.kv[s] = "";
        else a.kv[s.substr(0, eq)] = s.substr(eq + 1);
    }
    return a;
}

// counter-based generator (the reference uses an unseeded rand(), OON.cpp:796-801; same distributions)
uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
double uniform(uint64_t seed, uint64_t index, uint64_t field) {
    const uint64_t bits = splitmix(splitmix(seed * 0x100000001B3ull + index * 64 + field));
    return double(bits >> 11) * (1.0 / 9007199254740992.0);
}

void add_random_body_near(World& w, EntityID base_id, uint64_t index) {   // OON.cpp:778-803
    const float p_range = World::CFG_GLOBE_RADIUS * 30, v_range = World::CFG_GLOBE_RADIUS * 10;
    const Entity base = w.const_entity(base_id);
    const float m_min = base.mass / 9, m_max = base.mass * 3;
    Entity e;
    e.p.x = float(uniform(20240918, index, 0) * p_range - p_range / 2 + base.p.x);
    e.p.y = float(uniform(20240918, index, 1) * p_range - p_range / 2 + base.p.y);
    e.v.x = float(uniform(20240918, index, 2) * v_range - v_range / 2 + base.v.x * 0.05f);
    e.v.y = float(uniform(20240918, index, 3) * v_range - v_range / 2 + base.v.y * 0.05f);
    e.color = uint32_t(uniform(20240918, index, 5) * 0xffffff);
    e.mass = float(m_min + (m_max - m_min) * uniform(20240918, index, 4));
    w.add_body(e);
}

int usage() {
    std::puts("oon_headless --session=START.state --interact --friction=0.01 --fixed-dt=0.033 --loop-cap=20 --exit-on-finish\n"
              "             --session-save-as=END.state --no-save-compressed [--math=exact|fast] [--g-mode=R|H|0] [--bodies=N]\n"
              "             [--loop-mode=half|full] [--device=0] [--gpus=1] [--report] [--convert-only]");
    return 2;
}

}  // namespace

int main(int argc, char** argv) {
    const Args args = parse(argc, argv);
    if (args.has("help") || argc == 1) return usage();
    try {
        if (args.has("convert-only")) {            // snapshot codec only: no device is touched
            std::ifstream in(args.get("session"), std::ios::binary);
            SnapshotData s; std::string err;
            if (!in || !read_snapshot(in, s, err)) { std::fprintf(stderr, "load failed: %s\n", err.c_str()); return 1; }
            std::ofstream out(args.get("session-save-as"), std::ios::binary);
            const std::string ver = args.get("save-version", s.version);
            if (!write_snapshot(out, s, ver.c_str(), !args.has("no-save-compressed"))) { std::fprintf(stderr, "save failed\n"); return 1; }
            std::printf("converted %zu bodies (records of %zu bytes, version %s%s)\n", s.bodies.size(), s.record_bytes, s.version.c_str(),
                        s.was_compressed ? ", zstd" : "");
            return 0;
        }

        const int first_device = std::atoi(args.get("device", "0").c_str());
        const int ngpus = std::max(1, std::atoi(args.get("gpus", "1").c_str()));
        std::vector<int> devices;
        for (int g = 0; g < ngpus; ++g) devices.push_back(first_device + g);
        World world(devices);                      // one GPU: oon_create; several: oon_create_multi (single process, single mutator)
        bool load_failed = false;
        // ---- init_world_hook (OON.cpp:241-292), run through World::init() like the engine does (World.cpp:82-90)
        world.init_world_hook = [&](World& world) {
        if (args.has("session") && !args.get("session").empty()) {
            std::ifstream in(args.get("session"), std::ios::binary);
            if (!in || !World::load(in, &world)) { std::fprintf(stderr, "Failed to load session '%s'\n", args.get("session").c_str()); load_failed = true; }
        } else {
            Entity player;                         // "Player Superglobe" (test/OON.cfg:16-18)
            player.density = 550.f; player.mass = 1e27f; player.color = 0xffff20;
            player.add_thrusters(); player.superpower.free_color = true;
            player.superpower.gravity_immunity = !args.has("no-player-antigravity");   // sim/player_antigravity defaults to true
            world.add_body(player);
            if (args.has("bodies")) {
                const long n = std::max(0l, std::atol(args.get("bodies").c_str()));
                for (long i = 0; i < n; ++i) add_random_body_near(world, 0, uint64_t(i + 1));
            } else {                               // the two default moons (OON.cpp:273-281)
                const float R = World::CFG_GLOBE_RADIUS;
                Entity a; a.p = {R * 2, 0}; a.v = {0, -R * 2}; a.color = 0xff2020; a.mass = 3e24f; world.add_body(a);
                Entity b; b.p = {-R * 1.6f, R * 1.2f}; b.v = {-R * 1.8f, -R * 1.5f}; b.color = 0x3060ff; b.mass = 3e24f; world.add_body(b);
            }
        }
        };
        world.init();
        if (load_failed) return 1;
        if (args.has("interact")) world.props.interact_n2n = true;
        if (args.has("friction")) world.props.friction = std::stof(args.get("friction"));
        if (args.has("g-mode")) {                  // OONConfig.cpp:82-91
            const std::string g = args.get("g-mode");
            world.props.gravity_mode = (g == "R" || g == "r") ? GravityMode::Realistic : (g == "0" ? GravityMode::Off : GravityMode::Hyperbolic);
        }
        if (args.get("loop-mode") == "full") world.set_loop_mode(LoopMode::Full_Matrix);
        world.set_math_mode(args.get("math", "fast") == "exact" ? OON_MATH_EXACT : OON_MATH_FAST);

        const float dt = args.has("fixed-dt") ? std::stof(args.get("fixed-dt")) : 1.0f / 60;
        const long cap = std::atol(args.get("loop-cap", "0").c_str());
        if (cap <= 0 && !args.has("session-save-as")) { std::fprintf(stderr, "headless run needs --loop-cap=N > 0\n"); return 2; }

        const size_t n0 = world.entity_count();
        const auto t0 = std::chrono::steady_clock::now();
        if (cap > 0) world.update(dt, unsigned(cap));
        const size_t n1 = world.entity_count();        // synchronises
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        if (args.has("session-save-as")) {
            std::ofstream out(args.get("session-save-as"), std::ios::binary);
            SnapshotData s; s.props = world.props; s.bodies = world.bodies();
            if (!write_snapshot(out, s, nullptr, !args.has("no-save-compressed"))) { std::fprintf(stderr, "save failed\n"); return 1; }
        }
        if (args.has("report")) {
            const oon_events ev = world.drain_events();
            const double inter = world.props.interact_n2n ? double(cap) * double(n0) * double(n0 - 1) : double(cap) * double(n0 - 1);
            std::printf("{\"gpus\": %d, \"frames\": %ld, \"bodies_start\": %zu, \"bodies_end\": %zu, \"seconds\": %.6f, \"steps_per_sec\": %.3f, "
                        "\"interactions_per_sec\": %.6e, \"collisions\": %llu, \"touches\": %llu, \"removed\": %llu}\n",
                        world.device_count(), cap, n0, n1, secs, cap / secs, inter / secs, (unsigned long long)ev.collisions, (unsigned long long)ev.touches,
                        (unsigned long long)ev.removed);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oon_headless: %s\n", e.what());
        return 1;
    }
}
