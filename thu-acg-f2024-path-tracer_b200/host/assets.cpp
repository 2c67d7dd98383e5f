// Asset I/O for the host mirror: OBJ / baked .mesh loading (what tobj 4.0.2 provides to
// src/main.rs:408,433,458), RGB8 image loading (what image::to_rgb8 provides to src/texture.rs:62-69)
// from PNG or the baked raw container, and the 8-bit PNG writer behind Camera::render (camera.rs:118).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "pt_host.hpp"

namespace pt {

static std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// ---- meshes ---------------------------------------------------------------------------------------
ObjMesh ObjMesh::load(const std::string& path) {
    ObjMesh m;
    if (ends_with(path, ".mesh")) {  // tools/bake_assets.py container
        auto buf = read_file(path);
        if (buf.size() < 20 || memcmp(buf.data(), "PTM1", 4) != 0) throw std::runtime_error("bad .mesh file " + path);
        uint32_t n[4]; memcpy(n, buf.data() + 4, 16);
        size_t need = 20 + 4 * ((size_t)3 * n[0] + n[1] + (size_t)2 * n[2] + (size_t)3 * n[3]);
        if (buf.size() != need) throw std::runtime_error("truncated .mesh file " + path);
        const uint8_t* p = buf.data() + 20;
        m.positions.resize((size_t)3 * n[0]); memcpy(m.positions.data(), p, m.positions.size() * 4); p += m.positions.size() * 4;
        m.indices.resize(n[1]); memcpy(m.indices.data(), p, m.indices.size() * 4); p += m.indices.size() * 4;
        m.texcoords.resize((size_t)2 * n[2]); memcpy(m.texcoords.data(), p, m.texcoords.size() * 4); p += m.texcoords.size() * 4;
        m.normals.resize((size_t)3 * n[3]); memcpy(m.normals.data(), p, m.normals.size() * 4);
        return m;
    }
    // Wavefront OBJ, the subset tobj's OFFLINE_RENDERING_LOAD_OPTIONS exposes to the reference:
    // f32 attributes, fan triangulation, position indices only (single_index = false).
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line); std::string tag; ss >> tag;
        if (tag == "v") { std::string a, b, c; ss >> a >> b >> c; for (auto* s : {&a, &b, &c}) m.positions.push_back(strtof(s->c_str(), nullptr)); }
        else if (tag == "vt") { std::string a, b; ss >> a >> b; m.texcoords.push_back(strtof(a.c_str(), nullptr)); m.texcoords.push_back(strtof(b.c_str(), nullptr)); }
        else if (tag == "vn") { std::string a, b, c; ss >> a >> b >> c; for (auto* s : {&a, &b, &c}) m.normals.push_back(strtof(s->c_str(), nullptr)); }
        else if (tag == "f") {
            std::vector<uint32_t> vs; std::string tok;
            while (ss >> tok) {
                long i = strtol(tok.c_str(), nullptr, 10);
                long nv = (long)m.positions.size() / 3;
                vs.push_back((uint32_t)(i > 0 ? i - 1 : nv + i));
            }
            for (size_t k = 1; k + 1 < vs.size(); k++) { m.indices.push_back(vs[0]); m.indices.push_back(vs[k]); m.indices.push_back(vs[k + 1]); }
        }
    }
    return m;
}

// ---- PNG decode (8-bit RGB / RGBA / grey, non-interlaced) -------------------------------------------
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static ImagePtr decode_png(const std::vector<uint8_t>& buf, const std::string& path) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) throw std::runtime_error("not a PNG: " + path);
    uint32_t w = 0, h = 0; int ctype = -1, depth = 0, interlace = 0;
    std::vector<uint8_t> idat;
    size_t off = 8;
    while (off + 12 <= buf.size()) {
        uint32_t len = be32(&buf[off]); const uint8_t* type = &buf[off + 4]; const uint8_t* data = &buf[off + 8];
        if (off + 12 + len > buf.size()) break;
        if (!memcmp(type, "IHDR", 4)) { w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        off += 12 + len;
    }
    int ch = ctype == 2 ? 3 : ctype == 6 ? 4 : ctype == 0 ? 1 : ctype == 4 ? 2 : 0;
    if (depth != 8 || ch == 0 || interlace) throw std::runtime_error("unsupported PNG layout (need 8-bit non-interlaced): " + path);
    size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) throw std::runtime_error("PNG inflate failed: " + path);
    std::vector<uint8_t> px(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* in = &raw[(stride + 1) * y]; uint8_t ft = in[0]; in++;
        uint8_t* out = &px[stride * y]; const uint8_t* up = y ? &px[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= (size_t)ch ? out[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0, v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: throw std::runtime_error("bad PNG filter: " + path);
            }
            out[x] = (uint8_t)v;
        }
    }
    auto im = std::make_shared<Image>(); im->width = w; im->height = h; im->rgb.resize((size_t)3 * w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const uint8_t* p = &px[i * ch];
        if (ch >= 3) { im->rgb[3 * i] = p[0]; im->rgb[3 * i + 1] = p[1]; im->rgb[3 * i + 2] = p[2]; }  // to_rgb8 drops alpha
        else { im->rgb[3 * i] = im->rgb[3 * i + 1] = im->rgb[3 * i + 2] = p[0]; }
    }
    return im;
}
ImagePtr ImageTexture::load(const std::string& path) {
    auto buf = read_file(path);
    if (ends_with(path, ".png")) return decode_png(buf, path);
    if (ends_with(path, ".jpg") || ends_with(path, ".jpeg")) return decode_jpeg(buf);  // host/jpeg.cpp
    if (ends_with(path, ".hdr") || ends_with(path, ".pic")) return decode_hdr(buf);     // host/hdr.cpp
    if (ends_with(path, ".rgb8")) {  // 'PTI1', u32 w, u32 h, raw RGB (tools/bake_assets.py --raw)
        if (buf.size() < 12 || memcmp(buf.data(), "PTI1", 4) != 0) throw std::runtime_error("bad .rgb8 file " + path);
        uint32_t w, h; memcpy(&w, &buf[4], 4); memcpy(&h, &buf[8], 4);
        if (buf.size() != 12 + (size_t)3 * w * h) throw std::runtime_error("truncated .rgb8 file " + path);
        return from_rgb8(buf.data() + 12, w, h);
    }
    throw std::runtime_error("unsupported image format (.png, .jpg, .hdr or a baked .rgb8): " + path);
}

// ---- PNG encode (RGB8, filter 0) --------------------------------------------------------------------
static void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    out.insert(out.end(), hdr, hdr + 8);
    if (len) out.insert(out.end(), data, data + len);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    out.insert(out.end(), c, c + 4);
}
bool write_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw(((size_t)3 * w + 1) * h);
    for (uint32_t y = 0; y < h; y++) { raw[((size_t)3 * w + 1) * y] = 0; memcpy(&raw[((size_t)3 * w + 1) * y + 1], rgb + (size_t)3 * w * y, (size_t)3 * w); }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) return false;
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    put_chunk(out, "IHDR", ihdr, 13);
    put_chunk(out, "IDAT", comp.data(), clen);
    put_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

}  // namespace pt
