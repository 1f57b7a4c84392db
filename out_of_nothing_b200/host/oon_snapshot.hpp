// oon_snapshot.hpp — reader / writer of the reference's world snapshot format (C++ host side, no device needed).
//
// Follows World::save / World::load (/root/reference/src/app/model/World_SaveLoad.cpp:36-72, :80-165) and
// Entity::save / Entity::load (/root/reference/src/app/model/Entity_SaveLoad.cpp:23-49, :53-77):
// text header `name = value`, the mandatory `- - -` separator, one `<ndx> : "<raw Entity bytes, " and \ escaped>"`
// line per body; optionally the whole file inside one zstd frame (engine option save_compressed).
// Records are the raw struct bytes: 60 B (float build, every shipped fixture but one) or 96 B (double build).
#pragma once

#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <map>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "oon_world.hpp"

namespace OON::Model {

struct SnapshotData {
    std::string version = VERSION;
    Properties props;
    std::vector<Entity> bodies;
    bool was_compressed = false;
    size_t record_bytes = sizeof(Entity);
};

namespace detail {

inline bool version_parse(const std::string& s, int out[3]) {
    out[0] = out[1] = out[2] = 0;
    return std::sscanf(s.c_str(), "%d.%d.%d", &out[0], &out[1], &out[2]) >= 2;
}
inline int version_cmp(const std::string& a, const std::string& b) {
    int x[3], y[3];
    version_parse(a, x); version_parse(b, y);
    for (int i = 0; i < 3; ++i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}

// zstd through dlopen: the image ships libzstd.so.1 without headers
struct Zstd {
    void* lib = nullptr;
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*bound)(size_t) = nullptr;
    unsigned long long (*content_size)(const void*, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libzstd.so.1", RTLD_NOW);
        if (!lib) return false;
        decompress = reinterpret_cast<decltype(decompress)>(dlsym(lib, "ZSTD_decompress"));
        compress = reinterpret_cast<decltype(compress)>(dlsym(lib, "ZSTD_compress"));
        bound = reinterpret_cast<decltype(bound)>(dlsym(lib, "ZSTD_compressBound"));
        content_size = reinterpret_cast<decltype(content_size)>(dlsym(lib, "ZSTD_getFrameContentSize"));
        is_error = reinterpret_cast<decltype(is_error)>(dlsym(lib, "ZSTD_isError"));
        return decompress && compress && bound && content_size && is_error;
    }
};
inline Zstd& zstd() { static Zstd z; return z; }

struct BodyF64 {   // double-build Entity, 96 bytes (test/user/session/"1body-realg [double].save")
    uint8_t gravity_immunity, free_color, pad[6];
    double lifetime, r, density, px, py, vx, vy;
    float T;
    uint32_t color;
    double mass;
    float thrust[4];
};
static_assert(sizeof(BodyF64) == 96, "double-build record");

}  // namespace detail

inline bool read_snapshot(std::istream& in, SnapshotData& out, std::string& err) {
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    out = SnapshotData{};
    if (data.size() >= 4 && std::memcmp(data.data(), "\x28\xb5\x2f\xfd", 4) == 0) {
        auto& z = detail::zstd();
        if (!z.load()) { err = "snapshot is zstd-compressed but libzstd.so.1 cannot be loaded"; return false; }
        unsigned long long size = z.content_size(data.data(), data.size());
        if (size == ~0ull || size == ~0ull - 1) size = 64ull * data.size() + (1 << 20);
        std::string raw(size_t(size), '\0');
        const size_t got = z.decompress(raw.data(), raw.size(), data.data(), data.size());
        if (z.is_error(got)) { err = "zstd frame could not be inflated"; return false; }
        raw.resize(got);
        data.swap(raw);
        out.was_compressed = true;
    }
    const size_t sep = data.find("- - -");
    if (sep == std::string::npos) { err = "Failed to read world data! (no '- - -' separator)"; return false; }
    std::map<std::string, std::string> kv;
    {
        std::istringstream hs(data.substr(0, sep));
        for (std::string name, eq, val; hs >> name >> eq >> val && eq == "=";) kv[name] = val;
    }
    out.version = kv.count("MODEL_VERSION") ? kv["MODEL_VERSION"] : "0.0.1";
    if (detail::version_cmp(out.version, VERSION) > 0) { err = "Unsupported snapshot version:" + out.version; return false; }
    try {
        out.props.friction = std::stof(kv.at("drag"));
        out.props.interact_n2n = std::stof(kv.at("interactions")) != 0;
        if (detail::version_cmp(out.version, "0.1.0") >= 0) out.props.gravity_mode = GravityMode(std::stoul(kv.at("gravity_mode")));
        if (detail::version_cmp(out.version, "0.1.2") >= 0) out.props.gravity = std::stof(kv.at("gravity_strength"));
    } catch (...) { err = "Invalid (type of) property in the loaded snapshot."; return false; }
    size_t count = 0;
    try { count = size_t(std::stoul(kv.at("objects"))); } catch (...) { err = "snapshot has no object count"; return false; }

    size_t pos = data.find('\n', sep);
    pos = pos == std::string::npos ? data.size() : pos + 1;
    out.bodies.reserve(count);
    std::string rec;
    for (size_t n = 0; n < count; ++n) {
        const size_t colon = data.find(':', pos);
        const size_t q = colon == std::string::npos ? colon : data.find('"', colon);
        if (q == std::string::npos) { err = "Loading entity #" + std::to_string(n) + " failed!"; return false; }
        if (size_t(std::strtoull(data.c_str() + pos, nullptr, 10)) != n) { err = "body index out of sequence at #" + std::to_string(n); return false; }
        rec.clear();
        size_t i = q + 1;
        bool closed = false;
        while (i < data.size()) {                      // std::quoted(s, '"', '\\') extraction
            const char c = data[i];
            if (c == '\\' && i + 1 < data.size()) { rec.push_back(data[i + 1]); i += 2; }
            else if (c == '"') { closed = true; ++i; break; }
            else { rec.push_back(c); ++i; }
        }
        if (!closed) { err = "unterminated body record #" + std::to_string(n); return false; }
        if (n == 0) out.record_bytes = rec.size();
        if (rec.size() != out.record_bytes || (rec.size() != sizeof(Entity) && rec.size() != sizeof(detail::BodyF64))) {
            err = "Failed to load object! Bytes expected: " + std::to_string(sizeof(Entity)) + ", found: " + std::to_string(rec.size());
            return false;
        }
        Entity e;
        if (rec.size() == sizeof(Entity)) std::memcpy(static_cast<void*>(&e), rec.data(), sizeof e);
        else {
            detail::BodyF64 d;
            std::memcpy(&d, rec.data(), sizeof d);
            e.superpower.gravity_immunity = d.gravity_immunity != 0; e.superpower.free_color = d.free_color != 0;
            e.lifetime = float(d.lifetime); e.r = float(d.r); e.density = float(d.density);
            e.p.x = float(d.px); e.p.y = float(d.py); e.v.x = float(d.vx); e.v.y = float(d.vy);
            e.T = d.T; e.color = d.color; e.mass = float(d.mass);
            e.thrust_up = d.thrust[0]; e.thrust_down = d.thrust[1]; e.thrust_left = d.thrust[2]; e.thrust_right = d.thrust[3];
        }
        out.bodies.push_back(e);
        pos = i;
        while (pos < data.size() && (data[pos] == ' ' || data[pos] == '\r' || data[pos] == '\n' || data[pos] == '\t')) ++pos;
    }
    return true;
}

inline bool write_snapshot(std::ostream& os, const SnapshotData& s, const char* version = nullptr, bool compress = false) {
    const std::string ver = version ? version : VERSION;
    std::ostringstream out;
    out << "MODEL_VERSION = " << ver << '\n';
    out << "drag = " << s.props.friction << '\n';
    out << "interactions = " << s.props.interact_n2n << '\n';
    if (detail::version_cmp(ver, "0.1.0") >= 0) out << "gravity_mode = " << unsigned(s.props.gravity_mode) << '\n';
    if (detail::version_cmp(ver, "0.1.2") >= 0) out << "gravity_strength = " << s.props.gravity << '\n';
    out << "objects = " << s.bodies.size() << '\n';
    out << "- - -" << '\n';
    for (size_t n = 0; n < s.bodies.size(); ++n) {
        out << n << " : \"";
        const char* raw = reinterpret_cast<const char*>(&s.bodies[n]);
        for (size_t i = 0; i < sizeof(Entity); ++i) {
            if (raw[i] == '"' || raw[i] == '\\') out.put('\\');
            out.put(raw[i]);
        }
        out << "\"\n";
    }
    std::string text = out.str();
    if (compress) {
        auto& z = detail::zstd();
        if (!z.load()) return false;
        std::string packed(z.bound(text.size()), '\0');
        const size_t got = z.compress(packed.data(), packed.size(), text.data(), text.size(), 3);
        if (z.is_error(got)) return false;
        packed.resize(got);
        text.swap(packed);
    }
    os.write(text.data(), std::streamsize(text.size()));
    return os && !os.bad();
}

// ---- World::save / World::load in terms of the codec --------------------------------------------------------
inline bool World::save(std::ostream& out, const char* version) {
    SnapshotData s;
    s.props = props;
    s.bodies = bodies();
    return write_snapshot(out, s, version, false);
}

inline bool World::load(std::istream& in, World* result) {
    if (!result) return false;                 // "verify only" is not implemented in the reference either (World_SaveLoad.cpp:121-124)
    SnapshotData s;
    std::string err;
    if (!read_snapshot(in, s, err)) return false;
    const LoopMode lm = result->loop_mode();
    result->props.friction = s.props.friction;
    result->props.interact_n2n = s.props.interact_n2n;
    if (detail::version_cmp(s.version, "0.1.0") >= 0) result->props.gravity_mode = s.props.gravity_mode;
    if (detail::version_cmp(s.version, "0.1.2") >= 0) result->props.gravity = s.props.gravity;
    result->set_loop_mode(lm);
    while (result->entity_count()) result->remove_body(result->entity_count() - 1);
    for (auto const& e : s.bodies) result->add_body(e);    // add_body() recalculates the radius (World.cpp:64-70)
    return true;
}

}  // namespace OON::Model
