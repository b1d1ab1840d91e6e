// offline <in.glb> <skybox.exr> <out.exr> [spp=16] [--extent WxH] [--device N] [--dump]
//
// The reference's headless entry point (offline/main.zig:27-50, 80-203) as a native executable over the C ABI of libmoonshine.so -
// what `offline/main.zig` becomes once `Scene.fromGlbExr` feeds HdMoonshine* calls instead of the Vulkan managers.
// It implements the reference's import rules itself (no Python, no third-party loader):
//   * GLB: engine/hrtsystem/World.zig:44-228 (material mapping), :233-363 (one instance per node with a mesh, one mesh per
//     primitive, `sampled` <=> material name starts with "Emitter", rows (x, z, y) of the global transform = Y-up -> Z-up),
//     engine/hrtsystem/Camera.zig:26-51 (first camera node), deps/zgltf/src/main.zig:236-276 + helpers.zig:17-101 (TRS recompose,
//     parent chain, binary32 matrix product in zgltf's order);
//   * PNG: RGB8, non-interlaced (World.zig:50-59: img.pixels.rgb24), inflate by zlib;
//   * EXR: engine/fileformats/exr.zig:137-231 - skybox read as RGBA32F (scanline, NONE / ZIPS / ZIP, HALF / FLOAT), result written as
//     three FLOAT channels B, G, R, uncompressed.
// `--dump` prints a digest of the imported scene instead of rendering (tests/test_offline_native.py compares it with the Python
// front end's import of the same file, field by field, without a GPU); `--exr-copy in.exr out.exr` exercises the EXR reader and
// writer. Multi-GPU rendering stays with `python -m moonshine_b200.offline --gpus N`.
#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/moonshine_ext.h"

namespace {

[[noreturn]] void fail(const std::string& what) { throw std::runtime_error(what); }

std::vector<uint8_t> readFile(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail("cannot open " + path);
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d((size_t)n);
    if (n && fread(d.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); fail("short read on " + path); }
    fclose(f);
    return d;
}
template <class T>
T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof v); return v; }

std::vector<uint8_t> inflateAll(const uint8_t* src, size_t n, size_t expected) {
    std::vector<uint8_t> out(expected);
    uLongf len = (uLongf)expected;
    if (uncompress(out.data(), &len, src, (uLong)n) != Z_OK || len != expected) fail("zlib: corrupt or unexpectedly sized stream");
    return out;
}

// ---- JSON (the subset glTF uses) ----------------------------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order kept (glTF attribute maps are iterated in file order)
    const Json* find(const char* key) const {
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return find(key) != nullptr; }
    const Json& at(const char* key) const { const Json* j = find(key); if (!j) fail(std::string("glTF: missing key ") + key); return *j; }
    const Json& at(size_t i) const { if (i >= arr.size()) fail("glTF: index out of range"); return arr[i]; }
    double number(const char* key, double dflt) const { const Json* j = find(key); return j && j->kind == Num ? j->num : dflt; }
    size_t index(const char* key) const { return (size_t)at(key).num; }
    static const Json& empty() { static const Json e; return e; }
    const Json& get(const char* key) const { const Json* j = find(key); return j ? *j : empty(); }
};
struct JsonParser {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    Json parse() {
        ws();
        if (p >= end) fail("json: unexpected end");
        Json j;
        if (*p == '{') {
            j.kind = Json::Obj; p++; ws();
            if (*p == '}') { p++; return j; }
            for (;;) {
                ws(); Json k = parse(); if (k.kind != Json::Str) fail("json: key is not a string");
                ws(); if (*p++ != ':') fail("json: ':' expected");
                j.obj.emplace_back(k.str, parse());
                ws(); if (*p == ',') { p++; continue; }
                if (*p == '}') { p++; return j; }
                fail("json: ',' or '}' expected");
            }
        }
        if (*p == '[') {
            j.kind = Json::Arr; p++; ws();
            if (*p == ']') { p++; return j; }
            for (;;) {
                j.arr.push_back(parse());
                ws(); if (*p == ',') { p++; continue; }
                if (*p == ']') { p++; return j; }
                fail("json: ',' or ']' expected");
            }
        }
        if (*p == '"') {
            j.kind = Json::Str; p++;
            while (p < end && *p != '"') {
                if (*p == '\\' && p + 1 < end) {
                    p++;
                    switch (*p) {
                        case 'n': j.str += '\n'; break; case 't': j.str += '\t'; break; case 'r': j.str += '\r'; break;
                        case 'b': j.str += '\b'; break; case 'f': j.str += '\f'; break;
                        case 'u': { const unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); j.str += cp < 128 ? (char)cp : '?'; p += 4; break; }
                        default: j.str += *p;
                    }
                    p++;
                } else j.str += *p++;
            }
            p++;
            return j;
        }
        if (!strncmp(p, "true", 4)) { j.kind = Json::Bool; j.b = true; p += 4; return j; }
        if (!strncmp(p, "false", 5)) { j.kind = Json::Bool; p += 5; return j; }
        if (!strncmp(p, "null", 4)) { p += 4; return j; }
        char* e = nullptr;
        j.kind = Json::Num; j.num = strtod(p, &e);
        if (e == p) fail("json: unexpected character");
        p = e;
        return j;
    }
};

// ---- scene as imported (mirrors moonshine_b200/scenes.py Scene) -------------------------------------------------------------------
struct Tex { std::vector<uint8_t> data; uint32_t w = 1, h = 1, fmt = 0; };
struct Mat { uint32_t type = 0, normal = 0, emissive = 0, color = 0, metalness = 0, roughness = 0; float ior = 1.5f; };
struct Mesh { std::vector<float> pos, nrm, uv; std::vector<uint32_t> idx; };
struct Geo { uint32_t mesh, material; bool sampled; };
struct Inst { float xf[12]; std::vector<Geo> geoms; };
struct Scene {
    std::vector<Tex> tex; std::vector<Mat> mat; std::vector<Mesh> mesh; std::vector<Inst> inst;
    Lens lens{};
    template <class T> uint32_t solid(std::initializer_list<T> v, uint32_t fmt) {
        Tex t; t.fmt = fmt; t.data.resize(v.size() * sizeof(T)); memcpy(t.data.data(), v.begin(), t.data.size());
        tex.push_back(std::move(t));
        return (uint32_t)tex.size() - 1;
    }
};

// ---- zgltf transforms: matrices are [column][row], binary32, accumulated in zgltf's order (helpers.zig:17-101) ---------------------
struct M4 { float m[4][4]; };
M4 identity() { M4 r{}; for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f; return r; }
M4 mul(const M4& left, const M4& right) {
    M4 res{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) { const float prod = left.m[k][r] * right.m[c][k]; s = s + prod; }
            res.m[c][r] = s;
        }
    return res;
}
M4 recompose(const float t[3], const float q[4], const float s[3]) {
    M4 T = identity(); T.m[3][0] = t[0]; T.m[3][1] = t[1]; T.m[3][2] = t[2];
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    M4 R = identity();
    R.m[0][0] = 1.0f - 2.0f * (yy + zz); R.m[0][1] = 2.0f * (xy + wz); R.m[0][2] = 2.0f * (xz - wy);
    R.m[1][0] = 2.0f * (xy - wz); R.m[1][1] = 1.0f - 2.0f * (xx + zz); R.m[1][2] = 2.0f * (yz + wx);
    R.m[2][0] = 2.0f * (xz + wy); R.m[2][1] = 2.0f * (yz - wx); R.m[2][2] = 1.0f - 2.0f * (xx + yy);
    M4 S = identity(); S.m[0][0] = s[0]; S.m[1][1] = s[1]; S.m[2][2] = s[2];
    return mul(T, mul(R, S));
}

struct Glb {
    Json json;
    std::vector<uint8_t> file;
    const uint8_t* bin = nullptr; size_t binLen = 0;
    std::vector<int> parent;
    explicit Glb(const std::string& path) : file(readFile(path)) {
        if (file.size() < 12 || rd<uint32_t>(file.data()) != 0x46546C67u) fail("OnlySupportsGlbInput: not a GLB file");
        size_t pos = 12;
        while (pos + 8 <= file.size()) {
            const uint32_t clen = rd<uint32_t>(&file[pos]), ctype = rd<uint32_t>(&file[pos + 4]);
            if (pos + 8 + clen > file.size()) fail("GLB: truncated chunk");
            if (ctype == 0x4E4F534Au) { JsonParser jp{(const char*)&file[pos + 8], (const char*)&file[pos + 8] + clen}; json = jp.parse(); }
            else if (ctype == 0x004E4942u) { bin = &file[pos + 8]; binLen = clen; }
            pos += 8 + (size_t)clen;
        }
        const Json& nodes = json.get("nodes");
        parent.assign(nodes.arr.size(), -1);
        for (size_t i = 0; i < nodes.arr.size(); i++)
            for (const Json& c : nodes.arr[i].get("children").arr) parent.at((size_t)c.num) = (int)i;
    }
    void view(size_t idx, const uint8_t*& p, size_t& n) const {
        const Json& bv = json.at("bufferViews").at(idx);
        const size_t o = (size_t)bv.number("byteOffset", 0);
        n = (size_t)bv.at("byteLength").num;
        if (o + n > binLen) fail("GLB: buffer view out of range");
        p = bin + o;
    }
    // accessor -> rows of `nc` components converted to T
    template <class T>
    std::vector<T> accessor(size_t idx, size_t& count, size_t& nc) const {
        const Json& a = json.at("accessors").at(idx);
        const Json& bv = json.at("bufferViews").at(a.index("bufferView"));
        const int ct = (int)a.at("componentType").num;
        const std::string& type = a.at("type").str;
        nc = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : type == "MAT4" ? 16 : 0;
        const size_t isz = ct == 5120 || ct == 5121 ? 1 : ct == 5122 || ct == 5123 ? 2 : ct == 5125 || ct == 5126 ? 4 : 0;
        if (!nc || !isz) fail("glTF: unsupported accessor type");
        count = (size_t)a.at("count").num;
        const size_t start = (size_t)bv.number("byteOffset", 0) + (size_t)a.number("byteOffset", 0);
        size_t stride = (size_t)bv.number("byteStride", 0);
        if (!stride) stride = isz * nc;
        if (count && start + (count - 1) * stride + isz * nc > binLen) fail("glTF: accessor out of range");
        std::vector<T> out(count * nc);
        for (size_t i = 0; i < count; i++)
            for (size_t c = 0; c < nc; c++) {
                const uint8_t* p = bin + start + i * stride + c * isz;
                switch (ct) {
                    case 5120: out[i * nc + c] = (T)rd<int8_t>(p); break;
                    case 5121: out[i * nc + c] = (T)rd<uint8_t>(p); break;
                    case 5122: out[i * nc + c] = (T)rd<int16_t>(p); break;
                    case 5123: out[i * nc + c] = (T)rd<uint16_t>(p); break;
                    case 5125: out[i * nc + c] = (T)rd<uint32_t>(p); break;
                    default: out[i * nc + c] = (T)rd<float>(p); break;
                }
            }
        return out;
    }
    M4 local(size_t i) const {   // main.zig:245-262
        const Json& n = json.at("nodes").at(i);
        if (const Json* m = n.find("matrix")) {
            M4 r{};
            for (int k = 0; k < 16; k++) r.m[k / 4][k % 4] = (float)m->at((size_t)k).num;
            return r;
        }
        float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        if (const Json* j = n.find("translation")) for (int k = 0; k < 3; k++) t[k] = (float)j->at((size_t)k).num;
        if (const Json* j = n.find("rotation")) for (int k = 0; k < 4; k++) q[k] = (float)j->at((size_t)k).num;
        if (const Json* j = n.find("scale")) for (int k = 0; k < 3; k++) s[k] = (float)j->at((size_t)k).num;
        return recompose(t, q, s);
    }
    M4 global(size_t i) const {   // main.zig:264-276
        M4 m = local(i);
        for (int p = parent[i]; p >= 0; p = parent[(size_t)p]) m = mul(local((size_t)p), m);
        return m;
    }
    // PNG -> RGB8 rows
    std::vector<uint8_t> pngRgb(size_t textureIndex, uint32_t& w, uint32_t& h) const {
        const Json& img = json.at("images").at(json.at("textures").at(textureIndex).index("source"));
        if (img.get("mimeType").str != "image/png") fail("only image/png textures are supported (World.zig:50)");
        const uint8_t* p; size_t n;
        view(img.index("bufferView"), p, n);
        static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        if (n < 8 || memcmp(p, sig, 8)) fail("PNG: bad signature");
        auto be32 = [](const uint8_t* q) { return ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]; };
        std::vector<uint8_t> idat;
        w = h = 0;
        for (size_t pos = 8; pos + 12 <= n;) {
            const uint32_t len = be32(p + pos);
            const uint8_t* type = p + pos + 4;
            if (pos + 12 + len > n) fail("PNG: truncated chunk");
            if (!memcmp(type, "IHDR", 4)) {
                w = be32(p + pos + 8); h = be32(p + pos + 12);
                if (p[pos + 16] != 8 || p[pos + 17] != 2) fail("PNG textures must be RGB8 (World.zig:59: img.pixels.rgb24)");
                if (p[pos + 20] != 0) fail("PNG: interlaced images are not supported");
            } else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), p + pos + 8, p + pos + 8 + len);
            else if (!memcmp(type, "IEND", 4)) break;
            pos += 12 + (size_t)len;
        }
        if (!w || !h) fail("PNG: no IHDR");
        const size_t stride = (size_t)w * 3;
        std::vector<uint8_t> raw = inflateAll(idat.data(), idat.size(), (stride + 1) * h);
        std::vector<uint8_t> out(stride * h);
        for (uint32_t y = 0; y < h; y++) {
            const uint8_t ft = raw[(stride + 1) * y];
            const uint8_t* s = &raw[(stride + 1) * y + 1];
            uint8_t* d = &out[stride * y];
            const uint8_t* up = y ? &out[stride * (y - 1)] : nullptr;
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= 3 ? d[i - 3] : 0, b = up ? up[i] : 0, c = (up && i >= 3) ? up[i - 3] : 0;
                int pred = 0;
                switch (ft) {
                    case 0: pred = 0; break;
                    case 1: pred = a; break;
                    case 2: pred = b; break;
                    case 3: pred = (a + b) / 2; break;
                    case 4: { const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                    default: fail("PNG: unknown filter type");
                }
                d[i] = (uint8_t)(s[i] + pred);
            }
        }
        return out;
    }
};

// World.zig:341-345 / Camera.zig:33-41: rows (x, z, y) of the column-major global matrix
void zup(const M4& m, float out[12]) {
    const int rows[3] = {0, 2, 1};
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out[4 * r + c] = m.m[c][rows[r]];
}

Scene loadGlb(const std::string& path) {
    const Glb g(path);
    const Json& J = g.json;
    Scene sc;
    // camera: first node with a camera (Camera.zig:26-51)
    const Json& nodes = J.get("nodes");
    int camNode = -1;
    for (size_t i = 0; i < nodes.arr.size() && camNode < 0; i++) if (nodes.arr[i].has("camera")) camNode = (int)i;
    if (camNode < 0) fail("NoCameraInGlb");
    float T[12];
    zup(g.global((size_t)camNode), T);
    auto mulVec = [&](const float v[3], float o[3]) { for (int r = 0; r < 3; r++) o[r] = (T[4 * r] * v[0] + T[4 * r + 1] * v[1]) + T[4 * r + 2] * v[2]; };
    const float negZ[3] = {0, 0, -1}, posY[3] = {0, 1, 0};
    float fwd[3], up[3];
    mulVec(negZ, fwd); mulVec(posY, up);
    const float len = std::sqrt((fwd[0] * fwd[0] + fwd[1] * fwd[1]) + fwd[2] * fwd[2]);
    for (float& f : fwd) f = f / len;
    const float yfov = (float)J.at("cameras").at(nodes.arr[(size_t)camNode].index("camera")).at("perspective").at("yfov").num;
    sc.lens = Lens{{T[3], T[7], T[11]}, {fwd[0], fwd[1], fwd[2]}, {up[0], up[1], up[2]}, yfov, 0.0f, 1.0f};
    // materials (World.zig:44-228)
    auto rgbTexture = [&](size_t texIndex, int mode) -> uint32_t {   // mode 0: RG8 (first two channels), 1: RGBA8 sRGB (alpha 255), 2 / 3: R8 from channel 0 / 1
        uint32_t w, h;
        const std::vector<uint8_t> rgb = g.pngRgb(texIndex, w, h);
        Tex t; t.w = w; t.h = h;
        const size_t n = (size_t)w * h;
        if (mode == 0) { t.fmt = MSNE_TEX_RG8_UNORM; t.data.resize(2 * n); for (size_t i = 0; i < n; i++) { t.data[2 * i] = rgb[3 * i]; t.data[2 * i + 1] = rgb[3 * i + 1]; } }
        else if (mode == 1) { t.fmt = MSNE_TEX_RGBA8_SRGB; t.data.resize(4 * n); for (size_t i = 0; i < n; i++) { memcpy(&t.data[4 * i], &rgb[3 * i], 3); t.data[4 * i + 3] = 255; } }
        else { t.fmt = MSNE_TEX_R8_UNORM; t.data.resize(n); for (size_t i = 0; i < n; i++) t.data[i] = rgb[3 * i + (size_t)(mode - 2)]; }
        sc.tex.push_back(std::move(t));
        return (uint32_t)sc.tex.size() - 1;
    };
    for (const Json& m : J.get("materials").arr) {
        const Json& pbr = m.get("pbrMetallicRoughness");
        const Json& ext = m.get("extensions");
        Mat mat;
        mat.normal = m.has("normalTexture") ? rgbTexture(m.at("normalTexture").index("index"), 0) : sc.solid<float>({0.5f, 0.5f}, MSNE_TEX_RG32F);
        if (m.has("emissiveTexture")) mat.emissive = rgbTexture(m.at("emissiveTexture").index("index"), 1);
        else {
            float ef[3] = {0, 0, 0};
            if (const Json* j = m.find("emissiveFactor")) for (int k = 0; k < 3; k++) ef[k] = (float)j->at((size_t)k).num;
            const float strength = (float)ext.get("KHR_materials_emissive_strength").number("emissiveStrength", 1.0);
            mat.emissive = sc.solid<float>({ef[0] * strength, ef[1] * strength, ef[2] * strength, 1.0f}, MSNE_TEX_RGBA32F);
        }
        mat.ior = (float)ext.get("KHR_materials_ior").number("ior", 1.5);
        if ((float)ext.get("KHR_materials_transmission").number("transmissionFactor", 0.0) == 1.0f) {
            mat.type = MSNE_MAT_GLASS;
            sc.mat.push_back(mat);
            continue;
        }
        if (pbr.has("baseColorTexture")) mat.color = rgbTexture(pbr.at("baseColorTexture").index("index"), 1);
        else {
            float c[3] = {1, 1, 1};
            if (const Json* j = pbr.find("baseColorFactor")) for (int k = 0; k < 3; k++) c[k] = (float)j->at((size_t)k).num;
            mat.color = sc.solid<float>({c[0], c[1], c[2], 1.0f}, MSNE_TEX_RGBA32F);
        }
        const float metallic = (float)pbr.number("metallicFactor", 1.0), roughness = (float)pbr.number("roughnessFactor", 1.0);
        if (pbr.has("metallicRoughnessTexture")) {
            // sic: metalness <- R, roughness <- G (World.zig:172-174); glTF itself stores roughness in G, metalness in B
            const size_t ti = pbr.at("metallicRoughnessTexture").index("index");
            mat.metalness = rgbTexture(ti, 2);
            mat.roughness = rgbTexture(ti, 3);
            mat.type = MSNE_MAT_STANDARD_PBR;
        } else if (metallic == 0.0f && roughness == 1.0f) {
            mat.type = MSNE_MAT_LAMBERT; mat.ior = 1.5f;
        } else if (metallic == 1.0f && roughness == 0.0f) {
            mat.type = MSNE_MAT_PERFECT_MIRROR; mat.color = 0; mat.ior = 1.5f;
        } else {
            mat.type = MSNE_MAT_STANDARD_PBR;
            mat.metalness = sc.solid<float>({metallic}, MSNE_TEX_R32F);
            mat.roughness = sc.solid<float>({roughness}, MSNE_TEX_R32F);
        }
        sc.mat.push_back(mat);
    }
    // nodes with meshes (World.zig:262-349)
    for (size_t i = 0; i < nodes.arr.size(); i++) {
        const Json& n = nodes.arr[i];
        if (!n.has("mesh")) continue;
        Inst inst;
        for (const Json& prim : J.at("meshes").at(n.index("mesh")).at("primitives").arr) {
            Mesh mesh;
            size_t cnt, nc;
            mesh.idx = g.accessor<uint32_t>(prim.index("indices"), cnt, nc);
            for (const auto& kv : prim.at("attributes").obj) {
                const size_t acc = (size_t)kv.second.num;
                if (kv.first == "POSITION") mesh.pos = g.accessor<float>(acc, cnt, nc);
                else if (kv.first.rfind("TEXCOORD", 0) == 0) mesh.uv = g.accessor<float>(acc, cnt, nc);
                else if (kv.first == "NORMAL") mesh.nrm = g.accessor<float>(acc, cnt, nc);
                else fail("UnhandledAttribute: " + kv.first + " (World.zig:308-311)");
            }
            const uint32_t material = (uint32_t)prim.index("material");
            const std::string& name = J.at("materials").at(material).get("name").str;
            sc.mesh.push_back(std::move(mesh));
            inst.geoms.push_back(Geo{(uint32_t)sc.mesh.size() - 1, material, name.rfind("Emitter", 0) == 0});
        }
        zup(g.global(i), inst.xf);
        sc.inst.push_back(std::move(inst));
    }
    return sc;
}

// ---- EXR -----------------------------------------------------------------------------------------------------------------------------
float halfToFloat(uint16_t h) {
    const uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31u;
    uint32_t man = h & 1023u, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else { int e = -1; do { man <<= 1; e++; } while (!(man & 1024u)); bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 1023u) << 13); }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}
// scanline EXR -> RGBA32F like tinyexr's LoadEXR (missing A = 1); NONE / ZIPS / ZIP, HALF / FLOAT
std::vector<float> readExr(const std::string& path, uint32_t& W, uint32_t& H) {
    const std::vector<uint8_t> d = readFile(path);
    if (d.size() < 8 || rd<int32_t>(d.data()) != 20000630) fail("not an OpenEXR file");
    if (rd<int32_t>(&d[4]) & 0x200) fail("tiled EXR not supported");
    size_t pos = 8;
    std::map<std::string, std::pair<size_t, size_t>> attrs;   // name -> (offset, size)
    auto cstr = [&](size_t& p) { std::string s; while (p < d.size() && d[p]) s += (char)d[p++]; p++; return s; };
    while (pos < d.size() && d[pos] != 0) {
        const std::string name = cstr(pos);
        cstr(pos);
        const int32_t size = rd<int32_t>(&d[pos]); pos += 4;
        attrs[name] = {pos, (size_t)size};
        pos += (size_t)size;
    }
    pos++;
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    if (!attrs.count("channels") || !attrs.count("compression") || !attrs.count("dataWindow")) fail("EXR: missing header attribute");
    for (size_t cp = attrs["channels"].first; d[cp] != 0;) {
        Chan c; c.name = cstr(cp);
        c.type = rd<int32_t>(&d[cp]);
        const int32_t xs = rd<int32_t>(&d[cp + 8]), ys = rd<int32_t>(&d[cp + 12]);
        cp += 16;
        if (xs != 1 || ys != 1 || (c.type != 1 && c.type != 2)) fail("unsupported EXR channel layout");
        chans.push_back(c);
    }
    const int comp = d[attrs["compression"].first];
    if (comp != 0 && comp != 2 && comp != 3) fail("EXR compression not supported (NONE, ZIPS, ZIP only)");
    const uint8_t* dw = &d[attrs["dataWindow"].first];
    const int x0 = rd<int32_t>(dw), y0 = rd<int32_t>(dw + 4), x1 = rd<int32_t>(dw + 8), y1 = rd<int32_t>(dw + 12);
    W = (uint32_t)(x1 - x0 + 1); H = (uint32_t)(y1 - y0 + 1);
    const int linesPerBlock = comp == 3 ? 16 : 1;
    const size_t nBlocks = (H + (size_t)linesPerBlock - 1) / (size_t)linesPerBlock;
    size_t bytesPerPixel = 0;
    for (const Chan& c : chans) bytesPerPixel += c.type == 1 ? 2 : 4;
    std::map<std::string, std::vector<float>> planes;
    for (const Chan& c : chans) planes[c.name].assign((size_t)W * H, 0.0f);
    for (size_t b = 0; b < nBlocks; b++) {
        const size_t off = (size_t)rd<uint64_t>(&d[pos + 8 * b]);
        const int y = rd<int32_t>(&d[off]);
        const size_t size = (size_t)rd<int32_t>(&d[off + 4]);
        const int nl = std::min(linesPerBlock, y1 - y + 1);
        const size_t expect = (size_t)nl * W * bytesPerPixel;
        std::vector<uint8_t> raw(d.begin() + (long)off + 8, d.begin() + (long)off + 8 + (long)size);
        if (comp && size < expect) {
            std::vector<uint8_t> z = inflateAll(raw.data(), raw.size(), expect);
            for (size_t i = 1; i < z.size(); i++) z[i] = (uint8_t)(z[i - 1] + z[i] - 128);   // predictor
            raw.resize(expect);
            const size_t half = (expect + 1) / 2;
            for (size_t i = 0; i < expect; i++) raw[i] = (i & 1) ? z[half + i / 2] : z[i / 2];   // de-interleave
        }
        size_t p = 0;
        for (int ln = 0; ln < nl; ln++)
            for (const Chan& c : chans) {
                float* row = &planes[c.name][(size_t)(y - y0 + ln) * W];
                for (uint32_t x = 0; x < W; x++) row[x] = c.type == 1 ? halfToFloat(rd<uint16_t>(&raw[p + 2 * (size_t)x])) : rd<float>(&raw[p + 4 * (size_t)x]);
                p += (size_t)W * (c.type == 1 ? 2 : 4);
            }
    }
    std::vector<float> out((size_t)W * H * 4, 1.0f);
    const char* rgb[3] = {"R", "G", "B"};
    for (int i = 0; i < 3; i++) {
        const std::vector<float>* src = planes.count(rgb[i]) ? &planes[rgb[i]] : (planes.count("Y") ? &planes["Y"] : nullptr);
        if (src) for (size_t k = 0; k < (size_t)W * H; k++) out[4 * k + (size_t)i] = (*src)[k];
    }
    if (planes.count("A")) for (size_t k = 0; k < (size_t)W * H; k++) out[4 * k + 3] = planes["A"][k];
    return out;
}
// three FLOAT channels B, G, R (exr.zig:187-200), NO_COMPRESSION, increasing-Y scanlines
void writeExr(const std::string& path, const float* rgba, uint32_t W, uint32_t H) {
    std::vector<uint8_t> head;
    auto put = [&](const void* p, size_t n) { head.insert(head.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    auto putStr = [&](const char* s) { put(s, strlen(s) + 1); };
    auto attr = [&](const char* name, const char* type, const std::vector<uint8_t>& v) { putStr(name); putStr(type); const int32_t n = (int32_t)v.size(); put(&n, 4); put(v.data(), v.size()); };
    auto bytes = [](std::initializer_list<int32_t> v) { std::vector<uint8_t> b(v.size() * 4); memcpy(b.data(), v.begin(), b.size()); return b; };
    auto fbytes = [](std::initializer_list<float> v) { std::vector<uint8_t> b(v.size() * 4); memcpy(b.data(), v.begin(), b.size()); return b; };
    const int32_t magic = 20000630, version = 2;
    put(&magic, 4); put(&version, 4);
    std::vector<uint8_t> chlist;
    for (const char* n : {"B", "G", "R"}) {
        chlist.push_back((uint8_t)n[0]); chlist.push_back(0);
        const int32_t rec[4] = {2, 0, 1, 1};   // FLOAT, pLinear 0 (+3 pad), sampling 1 x 1
        chlist.insert(chlist.end(), (const uint8_t*)rec, (const uint8_t*)rec + 16);
    }
    chlist.push_back(0);
    attr("channels", "chlist", chlist);
    attr("compression", "compression", {0});
    attr("dataWindow", "box2i", bytes({0, 0, (int32_t)W - 1, (int32_t)H - 1}));
    attr("displayWindow", "box2i", bytes({0, 0, (int32_t)W - 1, (int32_t)H - 1}));
    attr("lineOrder", "lineOrder", {0});
    attr("pixelAspectRatio", "float", fbytes({1.0f}));
    attr("screenWindowCenter", "v2f", fbytes({0.0f, 0.0f}));
    attr("screenWindowWidth", "float", fbytes({1.0f}));
    head.push_back(0);
    const size_t lineBytes = (size_t)3 * W * 4;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) fail("cannot write " + path);
    fwrite(head.data(), 1, head.size(), f);
    for (uint32_t y = 0; y < H; y++) { const uint64_t off = head.size() + 8 * (size_t)H + y * (8 + lineBytes); fwrite(&off, 8, 1, f); }
    std::vector<float> line((size_t)3 * W);
    for (uint32_t y = 0; y < H; y++) {
        const int32_t hdr[2] = {(int32_t)y, (int32_t)lineBytes};
        fwrite(hdr, 4, 2, f);
        for (uint32_t x = 0; x < W; x++) {
            const float* px = rgba + 4 * ((size_t)y * W + x);
            line[x] = px[2]; line[(size_t)W + x] = px[1]; line[2 * (size_t)W + x] = px[0];   // B, G, R planes
        }
        fwrite(line.data(), 4, line.size(), f);
    }
    fclose(f);
}

// ---- digest (for the CPU test) -------------------------------------------------------------------------------------------------------
uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
void dump(const Scene& sc) {
    printf("textures %zu materials %zu meshes %zu instances %zu\n", sc.tex.size(), sc.mat.size(), sc.mesh.size(), sc.inst.size());
    printf("lens %016llx\n", (unsigned long long)fnv(&sc.lens, sizeof(Lens)));
    for (const Tex& t : sc.tex) printf("tex %u %u %u %016llx\n", t.w, t.h, t.fmt, (unsigned long long)fnv(t.data.data(), t.data.size()));
    for (const Mat& m : sc.mat) { uint32_t bits; memcpy(&bits, &m.ior, 4); printf("mat %u %u %u %u %u %u %08x\n", m.type, m.normal, m.emissive, m.color, m.metalness, m.roughness, bits); }
    for (const Mesh& m : sc.mesh)
        printf("mesh %zu %zu %016llx %016llx %016llx %016llx\n", m.pos.size() / 3, m.idx.size() / 3, (unsigned long long)fnv(m.pos.data(), m.pos.size() * 4),
               (unsigned long long)fnv(m.idx.data(), m.idx.size() * 4), (unsigned long long)fnv(m.nrm.data(), m.nrm.size() * 4), (unsigned long long)fnv(m.uv.data(), m.uv.size() * 4));
    for (const Inst& i : sc.inst) {
        printf("inst %016llx", (unsigned long long)fnv(i.xf, 48));
        for (const Geo& g : i.geoms) printf(" %u/%u/%d", g.mesh, g.material, g.sampled ? 1 : 0);
        printf("\n");
    }
}

#define CHECKED(call) do { if (!(call)) fail(std::string(#call) + ": " + HdMoonshineExtLastError()); } while (0)

int run(int argc, char** argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    Extent2D extent{1280, 720};   // offline/main.zig:48
    int device = 0;
    bool dumpOnly = false;
    for (size_t i = 0; i < args.size();) {
        if (args[i] == "--extent" && i + 1 < args.size()) {
            unsigned w = 0, h = 0;
            if (sscanf(args[i + 1].c_str(), "%ux%u", &w, &h) != 2 || !w || !h) fail("--extent expects WxH");
            extent = Extent2D{w, h};
            args.erase(args.begin() + (long)i, args.begin() + (long)i + 2);
        } else if (args[i] == "--device" && i + 1 < args.size()) { device = atoi(args[i + 1].c_str()); args.erase(args.begin() + (long)i, args.begin() + (long)i + 2); }
        else if (args[i] == "--dump") { dumpOnly = true; args.erase(args.begin() + (long)i); }
        else i++;
    }
    if (args.size() == 3 && args[0] == "--exr-copy") {   // test hook: read any supported EXR, write it back the way results are written
        uint32_t w = 0, h = 0;
        const std::vector<float> img = readExr(args[1], w, h);
        writeExr(args[2], img.data(), w, h);
        printf("%u %u %016llx\n", w, h, (unsigned long long)fnv(img.data(), img.size() * 4));
        return 0;
    }
    if (args.size() < (dumpOnly ? 1u : 3u)) {
        fprintf(stderr, "usage: offline <in.glb> <skybox.exr> <out.exr> [spp=16] [--extent WxH] [--device N] | offline <in.glb> --dump\n");
        return 2;   // error.BadArgs
    }
    auto endsWith = [](const std::string& s, const char* suffix) { const size_t n = strlen(suffix); return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0; };
    if (!endsWith(args[0], ".glb")) fail("OnlySupportsGlbInput");
    const auto t0 = std::chrono::steady_clock::now();
    const Scene sc = loadGlb(args[0]);
    if (dumpOnly) { dump(sc); return 0; }
    if (!endsWith(args[1], ".exr")) fail("OnlySupportsExrSkybox");
    if (!endsWith(args[2], ".exr")) fail("OnlySupportsExrOutput");
    const uint32_t spp = args.size() > 3 ? (uint32_t)atoi(args[3].c_str()) : 16u;
    uint32_t skyW = 0, skyH = 0;
    const std::vector<float> sky = readExr(args[1], skyW, skyH);

    CHECKED(HdMoonshineExtSetDevice(device));
    HdMoonshine* m = HdMoonshineCreate();
    if (!m) fail(std::string("HdMoonshineCreate: ") + HdMoonshineExtLastError());
    const MsneSpecConstants offline = {1, 1024, 1, 1, 1, 1, 1};   // offline/main.zig:106-111
    CHECKED(HdMoonshineExtSetSpecConstants(m, &offline));
    std::vector<ImageHandle> tex;
    for (const Tex& t : sc.tex) tex.push_back(HdMoonshineExtCreateTexture(m, t.data.data(), Extent2D{t.w, t.h}, (MsneTexFormat)t.fmt, ""));
    std::vector<MaterialHandle> mats;
    for (const Mat& mt : sc.mat) {
        if (tex.empty()) fail("material without textures");
        const MsneMaterialInfo info = {mt.type, tex.at(mt.normal), tex.at(mt.emissive), tex.at(mt.color), tex.at(mt.metalness), tex.at(mt.roughness), mt.ior};
        mats.push_back(HdMoonshineExtCreateMaterial(m, &info));
    }
    std::vector<MeshHandle> meshes;
    for (const Mesh& me : sc.mesh)
        meshes.push_back(HdMoonshineExtCreateMeshIndexed(m, (const F32x3*)me.pos.data(), me.nrm.empty() ? nullptr : (const F32x3*)me.nrm.data(),
                                                         me.uv.empty() ? nullptr : (const F32x2*)me.uv.data(), me.pos.size() / 3, (const U32x3*)me.idx.data(), me.idx.size() / 3));
    for (const Inst& in : sc.inst) {
        std::vector<Geometry> geoms;
        for (const Geo& g : in.geoms) geoms.push_back(Geometry{meshes.at(g.mesh), mats.at(g.material), g.sampled});
        Mat3x4 t;
        memcpy(&t, in.xf, 48);
        HdMoonshineCreateInstance(m, t, geoms.data(), geoms.size(), true);
    }
    CHECKED(HdMoonshineExtSetBackground(m, sky.data(), Extent2D{skyW, skyH}));
    CHECKED(HdMoonshineExtBuildEmitterTable(m));
    const SensorHandle sensor = HdMoonshineCreateSensor(m, extent);
    const LensHandle lens = HdMoonshineCreateLens(m, sc.lens);
    printf("%.3f seconds to load world\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    const auto t1 = std::chrono::steady_clock::now();
    CHECKED(HdMoonshineExtRenderLaunches(m, sensor, lens, spp));
    CHECKED(HdMoonshineExtReadSensor(m, sensor));
    MsneStats st;
    HdMoonshineExtGetStats(m, &st);
    printf("%.3f seconds to render (%.1f Mrays/s)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(),
           (double)(st.rays_closest + st.rays_shadow) / std::max(st.device_ms, 1e-9) / 1e3);
    writeExr(args[2], HdMoonshineGetSensorData(m, sensor), extent.width, extent.height);
    HdMoonshineDestroy(m);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    try { return run(argc, argv); } catch (const std::exception& e) { fprintf(stderr, "offline: %s\n", e.what()); return 1; }
}
