// obj_reader.cpp — host mirror of obj_reader.zig:21-198.
#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/zrt_host.h"

namespace {

// std.mem.tokenize(u8, line, " "): split on runs of the delimiter
std::vector<std::string> tokenize(const std::string &s, char delim) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && s[i] == delim) i++;
        size_t j = i;
        while (j < s.size() && s[j] != delim) j++;
        if (j > i) out.emplace_back(s, i, j - i);
        i = j;
    }
    return out;
}

bool parseFloat(const std::string &s, float *out) { // std.fmt.parseFloat(f32, ..): correctly rounded here
    if (s.empty()) return false;
    char *end = nullptr;
    *out = std::strtof(s.c_str(), &end);
    return end && *end == '\0';
}

// parseFaceVertex obj_reader.zig:21-43: "v", "v/vt", "v/vt/vn", "v//vn"; only v is used (Q10)
bool parseFaceVertex(const std::string &s, uint64_t *v) {
    const std::vector<std::string> parts = tokenize(s, '/');
    if (parts.empty()) return false;
    char *end = nullptr;
    *v = std::strtoull(parts[0].c_str(), &end, 10);
    if (!end || *end != '\0' || parts[0].empty() || parts[0][0] == '-' || parts[0][0] == '+') return false;
    for (size_t k = 1; k < parts.size() && k < 3; k++) { // texture / normal indices must parse, then are ignored
        std::strtoull(parts[k].c_str(), &end, 10);
        if (!end || *end != '\0') return false;
    }
    return true;
}

} // namespace

extern "C" int zrt_host_read_obj(const char *path, uint32_t material, zrt_triangle **triangles, uint32_t *n_triangles) {
    if (!path || !triangles || !n_triangles) return ZRT_ERR_INVALID;
    gzFile f = gzopen(path, "rb"); // transparently reads plain files too
    if (!f) return ZRT_ERR_IO;
    std::vector<zrt_vec3> vertexes;
    std::vector<zrt_triangle> tris;
    std::string line;
    std::vector<char> buf(20001); // readUntilDelimiterAlloc(.., '\n', 20000) obj_reader.zig:135
    int rc = ZRT_OK;
    while (gzgets(f, buf.data(), (int)buf.size())) {
        line.assign(buf.data());
        if (!line.empty() && line.back() == '\n') line.pop_back();
        if (line.size() < 1) continue;
        if (line.back() == '\r') line.pop_back(); // obj_reader.zig:144-146
        if (line.size() < 2) continue;
        if (line[0] == 'v' && line[1] == ' ') { // obj_reader.zig:147-154
            const auto tok = tokenize(line, ' ');
            zrt_vec3 v;
            if (tok.size() < 4 || !parseFloat(tok[1], &v.x) || !parseFloat(tok[2], &v.y) || !parseFloat(tok[3], &v.z)) {
                rc = ZRT_ERR_INVALID;
                break;
            }
            vertexes.push_back(v);
        } else if (line[0] == 'f' && line[1] == ' ') { // obj_reader.zig:155-169
            const auto tok = tokenize(line, ' ');
            std::vector<uint64_t> idx;
            bool ok = true;
            for (size_t k = 1; k < tok.size() && ok; k++) {
                uint64_t v;
                ok = parseFaceVertex(tok[k], &v) && v >= 1 && v <= vertexes.size();
                idx.push_back(v);
            }
            // parseTriangles obj_reader.zig:64-111: 3..6 vertices, fan (0,1,2),(2,3,0),(3,4,0),(4,5,0)
            if (!ok || idx.size() < 3 || idx.size() > 6) {
                rc = ZRT_ERR_INVALID; // ParseError.WrongNumberOfFaceVertexes
                break;
            }
            auto tri = [&](size_t a, size_t b, size_t c) {
                tris.push_back(zrt_triangle{vertexes[idx[a] - 1], vertexes[idx[b] - 1], vertexes[idx[c] - 1], material});
            };
            tri(0, 1, 2);
            for (size_t k = 3; k < idx.size(); k++) tri(k - 1, k, 0);
        }
        // `vn` lines are parsed and dropped by the reference (obj_reader.zig:170-177); other lines ignored
    }
    gzclose(f);
    if (rc != ZRT_OK) return rc;
    zrt_triangle *out = (zrt_triangle *)std::malloc(sizeof(zrt_triangle) * (tris.empty() ? 1 : tris.size()));
    if (!out) return ZRT_ERR_OOM;
    if (!tris.empty()) std::memcpy(out, tris.data(), sizeof(zrt_triangle) * tris.size());
    *triangles = out;
    *n_triangles = (uint32_t)tris.size();
    return ZRT_OK;
}
