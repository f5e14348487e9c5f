// obj_reader.cpp — host mirror of obj_reader.zig:21-198, built for meshes of config-4 size.
//
// Same grammar and the same error behaviour as the reference's line loop, but the file is read once into memory and
// parsed by line-aligned chunks on separate host threads, without a heap allocation per token:
//   pass 1 (parallel)  every chunk parses its `v` lines into a local vertex list and its `f` lines into index tuples,
//                      remembering how many of its own vertices preceded each face;
//   pass 2 (parallel)  after a prefix sum over the chunks' vertex and triangle counts, vertices are copied to their
//                      global positions, then every face checks its indices against the number of vertices that
//                      had been read when the reference reached that line (obj_reader.zig:54-62 indexes
//                      `vertexes.items`, a bounds-checked slice) and emits its fan of triangles in file order.
// Numbers go through std::from_chars (correctly rounded, like std.fmt.parseFloat); anything it does not take in one
// piece (a leading '+', hex floats, out-of-range values) falls back to strtof, the previous implementation.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/zrt_host.h"
#include "../zrt_internal.h"

namespace {

constexpr size_t kMaxLine = 20000; // readUntilDelimiterAlloc(.., '\n', 20000) obj_reader.zig:138: longer lines are an error

struct Face {
    uint64_t idx[6];
    uint32_t n;           // 3..6 vertices
    uint32_t local_verts; // vertices of this chunk read before the face line
};

struct Chunk {
    const char *begin = nullptr, *end = nullptr;
    std::vector<zrt_vec3> verts;
    std::vector<Face> faces;
    size_t n_tris = 0;
    bool error = false;
};

// std.mem.tokenize(u8, s, delim): next token of [p, e) split on runs of the delimiter; false when none is left
inline bool nextToken(const char *&p, const char *e, char delim, const char *&tb, const char *&te) {
    while (p < e && *p == delim) p++;
    if (p >= e) return false;
    tb = p;
    while (p < e && *p != delim) p++;
    te = p;
    return true;
}

bool parseFloat(const char *b, const char *e, float *out) { // std.fmt.parseFloat(f32, ..)
    if (b >= e) return false;
    const auto r = std::from_chars(b, e, *out);
    if (r.ec == std::errc() && r.ptr == e) return true;
    const std::string s(b, e);
    char *end = nullptr;
    *out = std::strtof(s.c_str(), &end);
    return end && *end == '\0' && end != s.c_str();
}

bool parseThreeFloats(const char *p, const char *e, zrt_vec3 *v) { // after the keyword: x y z, more tokens are ignored
    const char *tb, *te;
    if (!nextToken(p, e, ' ', tb, te)) return false; // eats "v" / "vn"
    float f[3];
    for (int k = 0; k < 3; k++)
        if (!nextToken(p, e, ' ', tb, te) || !parseFloat(tb, te, &f[k])) return false;
    *v = zrt_vec3{f[0], f[1], f[2]};
    return true;
}

bool parseUnsigned(const char *b, const char *e, bool allow_sign, uint64_t *out) {
    if (allow_sign && b < e && (*b == '+' || *b == '-')) b++;
    if (b >= e) return false;
    uint64_t v = 0;
    bool overflow = false;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return false;
        const uint64_t d = (uint64_t)(*b - '0');
        if (v > (UINT64_MAX - d) / 10) overflow = true;
        v = v * 10 + d;
    }
    *out = overflow ? UINT64_MAX : v; // never a valid vertex index
    return true;
}

// parseFaceVertex obj_reader.zig:21-43: "v", "v/vt", "v/vt/vn", "v//vn"; only v is used (Q10), the other two must parse
bool parseFaceVertex(const char *b, const char *e, uint64_t *v) {
    const char *p = b, *tb, *te;
    if (!nextToken(p, e, '/', tb, te) || !parseUnsigned(tb, te, false, v)) return false;
    uint64_t ignored;
    for (int k = 0; k < 2; k++) {
        if (!nextToken(p, e, '/', tb, te)) break;
        if (!parseUnsigned(tb, te, true, &ignored)) return false;
    }
    return true;
}

void parseChunk(Chunk &c) {
    { // size the two lists once: a counting scan over the line starts is far cheaper than growing them
        size_t nv = 0, nf = 0;
        for (const char *q = c.begin; q < c.end;) {
            if (q + 1 < c.end && q[1] == ' ') { nv += q[0] == 'v'; nf += q[0] == 'f'; }
            const char *nl = (const char *)std::memchr(q, '\n', (size_t)(c.end - q));
            if (!nl) break;
            q = nl + 1;
        }
        c.verts.reserve(nv);
        c.faces.reserve(nf);
    }
    const char *p = c.begin;
    while (p < c.end) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(c.end - p));
        const char *ls = p, *le = nl ? nl : c.end; // chunks end on a newline, so nl is never null in practice
        p = le + 1;
        if ((size_t)(le - ls) > kMaxLine) { c.error = true; return; } // error.StreamTooLong
        if (le - ls < 1) continue;
        if (le[-1] == '\r') le--; // obj_reader.zig:144-147
        if (le - ls < 2) continue;
        if (ls[0] == 'v' && ls[1] == ' ') { // obj_reader.zig:148-156
            zrt_vec3 v;
            if (!parseThreeFloats(ls, le, &v)) { c.error = true; return; }
            c.verts.push_back(v);
        } else if (ls[0] == 'f' && ls[1] == ' ') { // obj_reader.zig:157-172
            Face f;
            f.n = 0;
            f.local_verts = (uint32_t)c.verts.size();
            const char *q = ls, *tb, *te;
            nextToken(q, le, ' ', tb, te); // eats "f"
            bool ok = true;
            while (ok && nextToken(q, le, ' ', tb, te)) {
                uint64_t v;
                ok = parseFaceVertex(tb, te, &v) && f.n < 6; // a 7th vertex: ParseError.WrongNumberOfFaceVertexes
                if (ok) f.idx[f.n++] = v;
            }
            if (!ok || f.n < 3) { c.error = true; return; }
            c.n_tris += f.n - 2; // parseTriangles obj_reader.zig:64-111: fan (0,1,2),(2,3,0),(3,4,0),(4,5,0)
            c.faces.push_back(f);
        } else if (le - ls >= 3 && ls[0] == 'v' && ls[1] == 'n' && ls[2] == ' ') { // obj_reader.zig:173-181
            zrt_vec3 v; // parsed and dropped by the reference, but a malformed normal is still an error
            if (!parseThreeFloats(ls, le, &v)) { c.error = true; return; }
        }
    }
}

template <class F>
void forEachChunk(std::vector<Chunk> &chunks, F fn) {
    std::vector<zrt::Worker> th(chunks.size() - 1);
    for (size_t i = 1; i < chunks.size(); i++) th[i - 1] = zrt::Worker([&, i] { fn(i); });
    fn(0);
    for (auto &t : th) t.join();
}

} // namespace

extern "C" int zrt_host_read_obj(const char *path, uint32_t material, zrt_triangle **triangles, uint32_t *n_triangles) {
    if (!path || !triangles || !n_triangles) return ZRT_ERR_INVALID;
    zrt::BuildLap lap; // ZRT_TIMING=1
    // plain files are mapped, gzip files (this repository's assets) inflated into one buffer
    struct Text {
        const char *p = nullptr;
        size_t n = 0;
        void *mapped = nullptr;
        char *owned = nullptr;
        ~Text() {
            if (mapped) munmap(mapped, n);
            std::free(owned);
        }
    } text;
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) return ZRT_ERR_IO;
        unsigned char magic[2] = {0, 0};
        struct stat st;
        const bool have = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
        const ssize_t got = pread(fd, magic, 2, 0);
        if (have && !(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
            if (st.st_size > 0) {
                void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
                if (m == MAP_FAILED) { close(fd); return ZRT_ERR_IO; }
                text.mapped = m;
                text.p = (const char *)m;
                text.n = (size_t)st.st_size;
            }
            close(fd);
        } else {
            close(fd);
            gzFile f = gzopen(path, "rb");
            if (!f) return ZRT_ERR_IO;
            gzbuffer(f, 1u << 20);
            size_t cap = 1u << 22, used = 0;
            text.owned = (char *)std::malloc(cap);
            for (; text.owned;) {
                if (used == cap) {
                    char *grown = (char *)std::realloc(text.owned, cap *= 2);
                    if (!grown) { std::free(text.owned); text.owned = nullptr; break; }
                    text.owned = grown;
                }
                const int r = gzread(f, text.owned + used, (unsigned)std::min<size_t>(cap - used, 1u << 30));
                if (r < 0) { gzclose(f); return ZRT_ERR_IO; }
                if (r == 0) break;
                used += (size_t)r;
            }
            gzclose(f);
            if (!text.owned) return ZRT_ERR_OOM;
            text.p = text.owned;
            text.n = used;
        }
    }
    lap("obj: read file");
    // readUntilDelimiterAlloc returns EndOfStream for bytes after the last '\n': an unterminated last line is not read
    size_t len = text.n;
    while (len > 0 && text.p[len - 1] != '\n') len--;
    if (text.n - len > kMaxLine) return ZRT_ERR_INVALID; // the reference fails with StreamTooLong before it sees the end

    const size_t hw = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    const size_t n_chunks = std::max<size_t>(1, std::min(hw, len / (256u << 10)));
    std::vector<Chunk> chunks(n_chunks);
    try {
        const char *base = text.p;
        size_t at = 0;
        for (size_t i = 0; i < n_chunks; i++) {
            size_t stop = (i + 1 == n_chunks) ? len : std::max(at, len * (i + 1) / n_chunks);
            while (stop < len && stop > 0 && base[stop - 1] != '\n') stop++; // chunks end just after a newline
            chunks[i].begin = base + at;
            chunks[i].end = base + stop;
            at = stop;
        }
        forEachChunk(chunks, [&](size_t i) { parseChunk(chunks[i]); });
        for (const Chunk &c : chunks)
            if (c.error) return ZRT_ERR_INVALID;
        lap("obj: parse lines");

        std::vector<size_t> vert_base(n_chunks + 1, 0), tri_base(n_chunks + 1, 0);
        for (size_t i = 0; i < n_chunks; i++) {
            vert_base[i + 1] = vert_base[i] + chunks[i].verts.size();
            tri_base[i + 1] = tri_base[i] + chunks[i].n_tris;
        }
        if (tri_base[n_chunks] > 0xFFFFFFFFull) return ZRT_ERR_INVALID;
        std::vector<zrt_vec3> vertexes(vert_base[n_chunks]);
        forEachChunk(chunks, [&](size_t i) {
            if (!chunks[i].verts.empty())
                std::memcpy(vertexes.data() + vert_base[i], chunks[i].verts.data(), chunks[i].verts.size() * sizeof(zrt_vec3));
        });
        const size_t n_tris = tri_base[n_chunks];
        zrt_triangle *out = (zrt_triangle *)std::malloc(sizeof(zrt_triangle) * (n_tris ? n_tris : 1));
        if (!out) return ZRT_ERR_OOM;
        forEachChunk(chunks, [&](size_t i) {
            Chunk &c = chunks[i];
            zrt_triangle *t = out + tri_base[i];
            for (const Face &fc : c.faces) {
                const uint64_t seen = vert_base[i] + fc.local_verts; // vertexes.items.len when the reference got here
                for (uint32_t k = 0; k < fc.n; k++)
                    if (fc.idx[k] < 1 || fc.idx[k] > seen) { c.error = true; return; }
                auto tri = [&](uint32_t a, uint32_t b, uint32_t cc) {
                    *t++ = zrt_triangle{vertexes[fc.idx[a] - 1], vertexes[fc.idx[b] - 1], vertexes[fc.idx[cc] - 1], material};
                };
                tri(0, 1, 2);
                for (uint32_t k = 3; k < fc.n; k++) tri(k - 1, k, 0);
            }
        });
        for (const Chunk &c : chunks)
            if (c.error) { std::free(out); return ZRT_ERR_INVALID; }
        *triangles = out;
        *n_triangles = (uint32_t)n_tris;
        lap("obj: emit triangles");
    } catch (const std::bad_alloc &) {
        return ZRT_ERR_OOM;
    }
    return ZRT_OK;
}
