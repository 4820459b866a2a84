// On-disk cache of imported meshes and built hierarchies.  The reference re-runs the assimp import and
// the FCL BVH build in every lambda it starts (include/mpl/demo/load_mesh.hpp:232-293, se3_rigid_body_scenario.hpp:58-59);
// here a scene's construction looks the result up by content:
//   mesh file   key = hash(file bytes, shiftToCenter, identityRootTransform)  -> triangle soup + centre
//   hierarchy   key = hash(triangle bytes, build parameters)                  -> nodes + FP32 triangles + depth/radius
// Entries are written to a temporary name and renamed into place, carry a checksum of their payload and the format
// version, and are ignored (rebuilt and rewritten) whenever anything about them does not match.  Off unless a directory
// is given (mplb_set_cache_dir or MPLB_CACHE_DIR).
#include "mplb_host.hpp"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

namespace mplb {

namespace {

constexpr uint32_t kFormat = 3;   // bump when BvhNode / TriF32 / the builder's output changes
std::mutex gMu;
std::string gDir;
bool gDirRead = false;
std::atomic<uint64_t> gHits{0}, gMisses{0};

std::string cacheDir() {
    std::lock_guard<std::mutex> g(gMu);
    if (!gDirRead) {
        gDirRead = true;
        if (const char *e = std::getenv("MPLB_CACHE_DIR")) gDir = e;
    }
    return gDir;
}

// 128-bit content hash: two independent 64-bit multiply-xorshift lanes over 8-byte words (not cryptographic: the
// checksum and size checks catch a damaged entry, a collision needs two different meshes agreeing in 128 bits)
struct Hash128 {
    uint64_t a = 0x9e3779b97f4a7c15ull, b = 0xc2b2ae3d27d4eb4full;
    void word(uint64_t w) {
        a = (a ^ w) * 0xff51afd7ed558ccdull;
        a ^= a >> 32;
        b = (b + w) * 0xc4ceb9fe1a85ec53ull;
        b ^= b >> 29;
    }
    void bytes(const void *p, size_t n) {
        const unsigned char *c = (const unsigned char *)p;
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            std::memcpy(&w, c + i, 8);
            word(w);
        }
        uint64_t w = 0;
        if (i < n) std::memcpy(&w, c + i, n - i);
        word(w ^ ((uint64_t)n << 56));
    }
    std::string hex() const {
        char buf[40];
        std::snprintf(buf, sizeof(buf), "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
        return buf;
    }
};

struct Header {
    char magic[8];
    uint32_t format, kind;
    uint64_t count0, count1;
    int32_t depth, pad;
    double absMax, radius, center[3];
    uint64_t checkA, checkB;
};

bool readAll(const std::string &path, std::vector<unsigned char> &out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamoff n = f.tellg();
    if (n < 0) return false;
    out.resize((size_t)n);
    f.seekg(0);
    f.read((char *)out.data(), n);
    return (bool)f;
}

void writeAtomically(const std::string &path, const Header &h, const void *p0, size_t n0, const void *p1, size_t n1) {
    char tmp[64];
    std::snprintf(tmp, sizeof(tmp), ".tmp.%ld.%p", (long)getpid(), (const void *)&h);
    const std::string t = path + tmp;
    {
        std::ofstream f(t, std::ios::binary | std::ios::trunc);
        if (!f) return;
        f.write((const char *)&h, sizeof(h));
        if (n0) f.write((const char *)p0, (std::streamsize)n0);
        if (n1) f.write((const char *)p1, (std::streamsize)n1);
        if (!f) {
            std::remove(t.c_str());
            return;
        }
    }
    if (std::rename(t.c_str(), path.c_str()) != 0) std::remove(t.c_str());   // best effort: a cache may fail to write
}

Header makeHeader(uint32_t kind, uint64_t c0, uint64_t c1, const void *p0, size_t n0, const void *p1, size_t n1) {
    Header h{};
    std::memcpy(h.magic, "MPLBVH\0", 8);
    h.format = kFormat;
    h.kind = kind;
    h.count0 = c0;
    h.count1 = c1;
    Hash128 k;
    k.bytes(p0, n0);
    k.bytes(p1, n1);
    h.checkA = k.a;
    h.checkB = k.b;
    return h;
}

bool headerOk(const std::vector<unsigned char> &file, uint32_t kind, size_t elem0, size_t elem1, Header &h) {
    if (file.size() < sizeof(Header)) return false;
    std::memcpy(&h, file.data(), sizeof(Header));
    if (std::memcmp(h.magic, "MPLBVH\0", 8) != 0 || h.format != kFormat || h.kind != kind) return false;
    if (h.count0 > (1ull << 40) || h.count1 > (1ull << 40)) return false;
    const size_t n0 = (size_t)h.count0 * elem0, n1 = (size_t)h.count1 * elem1;
    if (file.size() != sizeof(Header) + n0 + n1) return false;
    Hash128 k;
    k.bytes(file.data() + sizeof(Header), n0);
    k.bytes(file.data() + sizeof(Header) + n0, n1);
    return k.a == h.checkA && k.b == h.checkB;
}

}  // namespace

int setCacheDir(const char *dir) {
    std::lock_guard<std::mutex> g(gMu);
    gDirRead = true;
    gDir.clear();
    if (!dir || !*dir) return 0;
    struct stat st;
    if (::stat(dir, &st) != 0) {
        if (::mkdir(dir, 0777) != 0) return -1;
    } else if (!S_ISDIR(st.st_mode)) {
        return -1;
    }
    gDir = dir;
    return 0;
}

void cacheStats(uint64_t *hits, uint64_t *misses) {
    if (hits) *hits = gHits.load();
    if (misses) *misses = gMisses.load();
}

FlatBvh buildBvhCached(const double *tris, size_t n) {
    const std::string dir = cacheDir();
    if (dir.empty() || n == 0) return buildBvh(tris, n);
    Hash128 key;
    key.word(kFormat);
    key.word(n);
    if (const char *e = std::getenv("MPLB_BVH_SPLIT")) key.bytes(e, std::strlen(e));   // (diagnostic builder switch)
    key.bytes(tris, n * 9 * sizeof(double));
    const std::string path = dir + "/bvh-" + key.hex() + ".mplb";
    std::vector<unsigned char> file;
    Header h;
    if (readAll(path, file) && headerOk(file, 1, sizeof(BvhNode), sizeof(TriF32), h) && h.count1 == n) {
        FlatBvh b;
        b.nodes.resize((size_t)h.count0);
        b.tris32.resize((size_t)h.count1);
        std::memcpy(b.nodes.data(), file.data() + sizeof(Header), b.nodes.size() * sizeof(BvhNode));
        std::memcpy(b.tris32.data(), file.data() + sizeof(Header) + b.nodes.size() * sizeof(BvhNode), b.tris32.size() * sizeof(TriF32));
        b.tris64.assign(tris, tris + 9 * n);
        b.depth = h.depth;
        b.absMax = h.absMax;
        b.radius = h.radius;
        ++gHits;
        return b;
    }
    ++gMisses;
    FlatBvh b = buildBvh(tris, n);
    Header w = makeHeader(1, b.nodes.size(), b.tris32.size(), b.nodes.data(), b.nodes.size() * sizeof(BvhNode), b.tris32.data(),
                          b.tris32.size() * sizeof(TriF32));
    w.depth = b.depth;
    w.absMax = b.absMax;
    w.radius = b.radius;
    writeAtomically(path, w, b.nodes.data(), b.nodes.size() * sizeof(BvhNode), b.tris32.data(), b.tris32.size() * sizeof(TriF32));
    return b;
}

TriangleSoup loadColladaCached(const std::string &meshPath, bool shiftToCenter, bool identityRootTransform) {
    const std::string dir = cacheDir();
    std::vector<unsigned char> src;
    if (dir.empty() || !readAll(meshPath, src)) return loadCollada(meshPath, shiftToCenter, identityRootTransform);   // (also raises the reference's error for a missing file)
    Hash128 key;
    key.word(kFormat);
    key.word((shiftToCenter ? 1u : 0u) | (identityRootTransform ? 2u : 0u));
    key.bytes(src.data(), src.size());
    const std::string path = dir + "/mesh-" + key.hex() + ".mplb";
    std::vector<unsigned char> file;
    Header h;
    if (readAll(path, file) && headerOk(file, 2, 9 * sizeof(double), 1, h) && h.count1 == 0) {
        TriangleSoup s;
        s.tris.resize((size_t)h.count0 * 9);
        std::memcpy(s.tris.data(), file.data() + sizeof(Header), s.tris.size() * sizeof(double));
        std::memcpy(s.center, h.center, sizeof(s.center));
        ++gHits;
        return s;
    }
    ++gMisses;
    TriangleSoup s = loadCollada(meshPath, shiftToCenter, identityRootTransform);
    Header w = makeHeader(2, s.size(), 0, s.tris.data(), s.tris.size() * sizeof(double), nullptr, 0);
    std::memcpy(w.center, s.center, sizeof(s.center));
    writeAtomically(path, w, s.tris.data(), s.tris.size() * sizeof(double), nullptr, 0);
    return s;
}

}  // namespace mplb
