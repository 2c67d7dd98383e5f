// Radiance RGBE (.hdr / .pic) decode for the host mirror, so that scenes 4 and 6 can load the reference's
// assets/grace_probe_latlong.hdr (src/main.rs:270,528) without the Python bake.
//
// What the reference does with such a file (src/texture.rs:62-69): image 0.25.5 decodes it to Rgb32F
// (RGBE -> c * 2^(e-136), e == 0 -> black) and `.to_rgb8()` turns every channel into
// round(clamp(x, 0, 1) * 255) (radiance above 1 is clipped, the result is an 8-bit texture).
// Both steps are restated here; the only tie of that rounding (x = 0.5 -> 127.5) rounds away from zero
// like f32::round.  Handles the new adaptive RLE (2,2,hi,lo scanline prefix), the old repeat-pixel RLE
// (1,1,1,n) and flat pixels, and the orientation "-Y H +X W" (the only one image 0.25 accepts too).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pt_host.hpp"

namespace pt {

namespace {
struct Reader {
    const uint8_t* p; const uint8_t* end;
    uint8_t byte() { if (p >= end) throw std::runtime_error("truncated Radiance HDR file"); return *p++; }
    std::string line() {
        std::string s;
        while (p < end && *p != '\n') s.push_back((char)*p++);
        if (p >= end) throw std::runtime_error("truncated Radiance HDR header");
        p++;
        return s;
    }
};

inline uint8_t rgbe_channel_to_u8(uint8_t c, uint8_t e) {
    if (e == 0) return 0;
    float x = std::ldexp((float)c, (int)e - 136);
    x = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
    return (uint8_t)std::round(x * 255.f);
}
}  // namespace

ImagePtr decode_hdr(const std::vector<uint8_t>& buf) {
    Reader r{buf.data(), buf.data() + buf.size()};
    std::string magic = r.line();
    if (magic.rfind("#?", 0) != 0) throw std::runtime_error("not a Radiance HDR file (no #? signature)");
    bool rgbe = false;
    for (;;) {  // header: KEY=value lines up to an empty one
        std::string l = r.line();
        if (l.empty()) break;
        if (l.rfind("FORMAT=", 0) == 0) {
            if (l != "FORMAT=32-bit_rle_rgbe") throw std::runtime_error("unsupported Radiance HDR format: " + l);
            rgbe = true;
        }
    }
    if (!rgbe) throw std::runtime_error("Radiance HDR header carries no FORMAT=32-bit_rle_rgbe line");
    std::string dims = r.line();
    long h = 0, w = 0;
    if (sscanf(dims.c_str(), "-Y %ld +X %ld", &h, &w) != 2 || h <= 0 || w <= 0 || h > 65535 || w > 65535)
        throw std::runtime_error("unsupported Radiance HDR orientation / size: " + dims);

    auto im = std::make_shared<Image>();
    im->width = (uint32_t)w; im->height = (uint32_t)h; im->rgb.resize((size_t)3 * w * h);
    std::vector<uint8_t> scan((size_t)4 * w);  // planar for the new RLE: R.. G.. B.. E..
    for (long y = 0; y < h; y++) {
        uint8_t* out = &im->rgb[(size_t)3 * w * y];
        uint8_t a = r.byte(), b = r.byte(), c = r.byte(), d = r.byte();
        if (a == 2 && b == 2 && !(c & 0x80) && w >= 8 && w < 32768) {  // adaptive RLE, one channel after the other
            if (((long)c << 8 | d) != w) throw std::runtime_error("Radiance HDR scanline length mismatch");
            for (int ch = 0; ch < 4; ch++) {
                uint8_t* dst = &scan[(size_t)ch * w];
                long x = 0;
                while (x < w) {
                    uint8_t n = r.byte();
                    if (n > 128) {  // run
                        n -= 128;
                        if (x + n > w) throw std::runtime_error("Radiance HDR run overflows its scanline");
                        uint8_t v = r.byte();
                        memset(dst + x, v, n); x += n;
                    } else {        // literal
                        if (n == 0 || x + n > w) throw std::runtime_error("Radiance HDR literal overflows its scanline");
                        for (uint8_t k = 0; k < n; k++) dst[x++] = r.byte();
                    }
                }
            }
            for (long x = 0; x < w; x++) {
                uint8_t e = scan[(size_t)3 * w + x];
                out[3 * x] = rgbe_channel_to_u8(scan[x], e);
                out[3 * x + 1] = rgbe_channel_to_u8(scan[(size_t)w + x], e);
                out[3 * x + 2] = rgbe_channel_to_u8(scan[(size_t)2 * w + x], e);
            }
        } else {  // flat pixels with the old (1,1,1,n) repeat marker
            uint8_t px[4] = {a, b, c, d};
            uint8_t prev[4] = {0, 0, 0, 0};
            long x = 0; int shift = 0; bool have = true;
            while (x < w) {
                if (!have) { px[0] = r.byte(); px[1] = r.byte(); px[2] = r.byte(); px[3] = r.byte(); }
                have = false;
                if (px[0] == 1 && px[1] == 1 && px[2] == 1) {
                    long n = (long)px[3] << shift;
                    if (x == 0 || x + n > w) throw std::runtime_error("bad Radiance HDR repeat marker");
                    for (long k = 0; k < n; k++, x++) for (int ch = 0; ch < 3; ch++) out[3 * x + ch] = rgbe_channel_to_u8(prev[ch], prev[3]);
                    shift += 8;
                } else {
                    memcpy(prev, px, 4);
                    for (int ch = 0; ch < 3; ch++) out[3 * x + ch] = rgbe_channel_to_u8(px[ch], px[3]);
                    x++; shift = 0;
                }
            }
        }
    }
    return im;
}

}  // namespace pt
