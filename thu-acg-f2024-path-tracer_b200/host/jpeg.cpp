// JPEG decoder for the host scene pipeline (the reference decodes assets/envmap.jpg and earthmap.jpg with
// the `image` crate, src/texture.rs:62-69; this removes the Python bake step for them).
//
// Baseline (SOF0/SOF1) and progressive (SOF2) Huffman JPEG, 8-bit, 1 or 3 components, restart intervals.  Written from
// ITU-T T.81; the inverse DCT and the YCbCr -> RGB conversion follow the integer formulations of the IJG library
// (jidctint "islow", jdcolor) so that 4:4:4 files decode to the same bytes as libjpeg-based decoders.  Chroma of
// subsampled files is replicated (no "fancy" interpolation), which may differ from other decoders by a few LSB at edges.
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "pt_host.hpp"

namespace pt {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool defined = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build() {  // T.81 Annex C / F.2.2.3
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
        defined = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;  // allocated block grid (padded to whole MCUs)
    int w = 0, h_px = 0;             // component size in samples
    int dc_pred = 0;
    std::vector<int16_t> coef;       // blocks_w * blocks_h * 64, natural (de-zigzagged) order
};

struct Decoder {
    const uint8_t* p; const uint8_t* end;
    uint16_t qt[4][64]; bool qt_ok[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    bool progressive = false;
    Component comp[3];
    int restart_interval = 0;
    // bit reader
    uint32_t bitbuf = 0; int bitcnt = 0; int marker = 0;  // marker: a marker met inside entropy data (0 = none)
    int eobrun = 0;

    [[noreturn]] static void bad(const char* why) { throw std::runtime_error(std::string("jpeg: ") + why); }
    int u8() { if (p >= end) bad("truncated"); return *p++; }
    int u16() { int a = u8(); return (a << 8) | u8(); }

    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    int c = p < end ? *p++ : 0xD9;
                    while (c == 0xFF && p < end) c = *p++;  // fill bytes
                    if (c != 0) { marker = c; b = 0; }       // a marker: feed zeros from here on
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int get_bits(int n) {
        if (n == 0) return 0;
        if (bitcnt < n) fill();
        int v = (int)(bitbuf >> (32 - n));
        bitbuf <<= n; bitcnt -= n;
        return v;
    }
    int get_bit() { return get_bits(1); }
    static int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }  // T.81 F.2.2.1
    int decode(const Huff& h) {  // T.81 F.2.2.3, bit by bit
        if (!h.defined) bad("missing Huffman table");
        int code = 0;
        for (int l = 1; l <= 16; l++) {
            code = (code << 1) | get_bit();
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        bad("bad Huffman code");
    }
    void reset_entropy() { bitbuf = 0; bitcnt = 0; marker = 0; eobrun = 0; for (auto& c : comp) c.dc_pred = 0; }

    // ---- one block of one scan
    void block_baseline(Component& c, int16_t* b) {
        int t = decode(dc[c.td]);
        int diff = t ? extend(get_bits(t), t) : 0;
        c.dc_pred += diff;
        b[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            int rs = decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s == 0) { if (r == 15) { k += 16; continue; } break; }
            k += r;
            if (k > 63) bad("bad AC run");
            b[kZigzag[k++]] = (int16_t)extend(get_bits(s), s);
        }
    }
    void block_dc(Component& c, int16_t* b, int ah, int al) {  // T.81 G.1.2.1
        if (ah == 0) {
            int t = decode(dc[c.td]);
            int diff = t ? extend(get_bits(t), t) : 0;
            c.dc_pred += diff;
            b[0] = (int16_t)(c.dc_pred * (1 << al));
        } else if (get_bit()) b[0] = (int16_t)(b[0] | (1 << al));
    }
    void block_ac(Component& c, int16_t* b, int ss, int se, int ah, int al) {  // T.81 G.1.2.2 / G.1.2.3
        if (ah == 0) {
            if (eobrun) { eobrun--; return; }
            for (int k = ss; k <= se;) {
                int rs = decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += get_bits(r); break; }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) bad("bad AC run");
                    b[kZigzag[k++]] = (int16_t)(extend(get_bits(s), s) * (1 << al));
                }
            }
            return;
        }
        const int bit = 1 << al;
        auto refine = [&](int16_t* q) { if (get_bit() && (*q & bit) == 0) *q = (int16_t)(*q > 0 ? *q + bit : *q - bit); };
        if (eobrun) {
            eobrun--;
            for (int k = ss; k <= se; k++) { int16_t* q = &b[kZigzag[k]]; if (*q != 0) refine(q); }
            return;
        }
        int k = ss;
        do {
            int rs = decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += get_bits(r); r = 64; }  // rest of this block: refinement only
            } else {
                if (s != 1) bad("bad refinement code");
                s = get_bit() ? bit : -bit;
            }
            while (k <= se) {
                int16_t* q = &b[kZigzag[k++]];
                if (*q != 0) refine(q);
                else { if (r == 0) { *q = (int16_t)s; break; } r--; }
            }
        } while (k <= se);
    }

    // ---- a scan
    void scan() {
        int len = u16(); (void)len;
        int ns = u8();
        if (ns < 1 || ns > ncomp) bad("bad scan component count");
        Component* sc[3];
        for (int i = 0; i < ns; i++) {
            int id = u8(), tt = u8(); sc[i] = nullptr;
            for (int k = 0; k < ncomp; k++) if (comp[k].id == id) sc[i] = &comp[k];
            if (!sc[i]) bad("scan names an unknown component");
            sc[i]->td = tt >> 4; sc[i]->ta = tt & 15;
            if (sc[i]->td > 3 || sc[i]->ta > 3) bad("bad table index");
        }
        int ss = u8(), se = u8(), a = u8(), ah = a >> 4, al = a & 15;
        if (!progressive) { ss = 0; se = 63; ah = al = 0; }
        else if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1)) bad("bad spectral selection");
        reset_entropy();
        int todo = restart_interval ? restart_interval : 0x7FFFFFFF, next_rst = 0;
        auto do_block = [&](Component& c, int bx, int by) {
            int16_t* b = c.coef.data() + ((size_t)by * c.blocks_w + bx) * 64;
            if (!progressive) block_baseline(c, b);
            else if (ss == 0) block_dc(c, b, ah, al);
            else block_ac(c, b, ss, se, ah, al);
        };
        auto after_unit = [&](bool last) {
            if (--todo > 0 || last) return;
            // restart: the marker follows the byte-aligned end of the interval
            if (!marker) { bitcnt = 0; bitbuf = 0; fill(); }
            if (marker < 0xD0 || marker > 0xD7 || marker != 0xD0 + next_rst) bad("missing restart marker");
            next_rst = (next_rst + 1) & 7;
            reset_entropy();
            todo = restart_interval;
        };
        if (ns == 1) {  // non-interleaved: the component's own block grid (T.81 A.2.2)
            Component& c = *sc[0];
            int bw = (c.w + 7) / 8, bh = (c.h_px + 7) / 8;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) { do_block(c, bx, by); after_unit(by == bh - 1 && bx == bw - 1); }
        } else {
            for (int my = 0; my < mcus_y; my++)
                for (int mx = 0; mx < mcus_x; mx++) {
                    for (int i = 0; i < ns; i++)
                        for (int y = 0; y < sc[i]->v; y++)
                            for (int x = 0; x < sc[i]->h; x++) do_block(*sc[i], mx * sc[i]->h + x, my * sc[i]->v + y);
                    after_unit(my == mcus_y - 1 && mx == mcus_x - 1);
                }
        }
        // position after the entropy-coded segment: either we already met the next marker, or it lies ahead
        if (marker) { p -= 2; marker = 0; }
        else while (p + 1 < end && !(p[0] == 0xFF && p[1] != 0 && p[1] != 0xFF && !(p[1] >= 0xD0 && p[1] <= 0xD7))) p++;
        bitcnt = 0; bitbuf = 0;
    }

    // ---- IJG jidctint ("islow"): 13-bit constants, 2 extra bits between the passes
    static void idct(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
        constexpr int CB = 13, P1 = 2;
        constexpr long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
                       F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
        long ws[64];
        auto pass = [&](long i0, long i1, long i2, long i3, long i4, long i5, long i6, long i7, int shift, long* o, int ostep, bool first) {
            long z1 = (i2 + i6) * F0541, t2 = z1 - i6 * F1847, t3 = z1 + i2 * F0765;
            long t0 = (i0 + i4) * (1L << CB), t1 = (i0 - i4) * (1L << CB);
            long t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            long a0 = i7, a1 = i5, a2 = i3, a3 = i1;
            long y1 = a0 + a3, y2 = a1 + a2, y3 = a0 + a2, y4 = a1 + a3, y5 = (y3 + y4) * F1175;
            a0 *= F0298; a1 *= F2053; a2 *= F3072; a3 *= F1501;
            y1 *= -F0899; y2 *= -F2562; y3 *= -F1961; y4 *= -F0390;
            y3 += y5; y4 += y5;
            a0 += y1 + y3; a1 += y2 + y4; a2 += y2 + y3; a3 += y1 + y4;
            const long r = 1L << (shift - 1);
            (void)first;
            o[0 * ostep] = (t10 + a3 + r) >> shift; o[7 * ostep] = (t10 - a3 + r) >> shift;
            o[1 * ostep] = (t11 + a2 + r) >> shift; o[6 * ostep] = (t11 - a2 + r) >> shift;
            o[2 * ostep] = (t12 + a1 + r) >> shift; o[5 * ostep] = (t12 - a1 + r) >> shift;
            o[3 * ostep] = (t13 + a0 + r) >> shift; o[4 * ostep] = (t13 - a0 + r) >> shift;
        };
        for (int c = 0; c < 8; c++)  // columns
            pass((long)in[c] * q[c], (long)in[8 + c] * q[8 + c], (long)in[16 + c] * q[16 + c], (long)in[24 + c] * q[24 + c],
                 (long)in[32 + c] * q[32 + c], (long)in[40 + c] * q[40 + c], (long)in[48 + c] * q[48 + c], (long)in[56 + c] * q[56 + c],
                 CB - P1, ws + c, 8, true);
        for (int r = 0; r < 8; r++) {  // rows
            long o[8];
            const long* w = ws + 8 * r;
            pass(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], CB + P1 + 3, o, 1, false);
            for (int k = 0; k < 8; k++) { long v = o[k] + 128; out[r * stride + k] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
        }
    }

    std::shared_ptr<Image> run() {
        if (u8() != 0xFF || u8() != 0xD8) bad("not a JPEG file");
        bool have_frame = false, done = false;
        while (!done) {
            int b = u8();
            if (b != 0xFF) continue;
            int m = u8();
            while (m == 0xFF) m = u8();
            if (m == 0 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            switch (m) {
                case 0xD9: done = true; break;
                case 0xDB: {  // DQT
                    int len = u16() - 2;
                    while (len > 0) {
                        int pq = u8(), prec = pq >> 4, id = pq & 15;
                        if (id > 3) bad("bad quantisation table id");
                        for (int k = 0; k < 64; k++) qt[id][kZigzag[k]] = (uint16_t)(prec ? u16() : u8());
                        qt_ok[id] = true;
                        len -= 1 + (prec ? 128 : 64);
                    }
                    break;
                }
                case 0xC4: {  // DHT
                    int len = u16() - 2;
                    while (len > 0) {
                        int tc = u8(), cls = tc >> 4, id = tc & 15;
                        if (cls > 1 || id > 3) bad("bad Huffman table id");
                        Huff& h = cls ? ac[id] : dc[id];
                        int n = 0;
                        h.bits[0] = 0;
                        for (int l = 1; l <= 16; l++) { h.bits[l] = (uint8_t)u8(); n += h.bits[l]; }
                        if (n > 256) bad("bad Huffman table");
                        for (int k = 0; k < n; k++) h.vals[k] = (uint8_t)u8();
                        h.build();
                        len -= 17 + n;
                    }
                    break;
                }
                case 0xC0: case 0xC1: case 0xC2: {  // SOF0/1 (sequential), SOF2 (progressive)
                    u16();
                    if (u8() != 8) bad("only 8-bit samples are supported");
                    height = u16(); width = u16(); ncomp = u8();
                    if (width <= 0 || height <= 0 || (ncomp != 1 && ncomp != 3)) bad("unsupported frame (need 1 or 3 components)");
                    progressive = m == 0xC2;
                    for (int k = 0; k < ncomp; k++) {
                        comp[k].id = u8(); int hv = u8(); comp[k].h = hv >> 4; comp[k].v = hv & 15; comp[k].tq = u8();
                        if (comp[k].h < 1 || comp[k].h > 4 || comp[k].v < 1 || comp[k].v > 4 || comp[k].tq > 3) bad("bad sampling factors");
                        hmax = std::max(hmax, comp[k].h); vmax = std::max(vmax, comp[k].v);
                    }
                    mcus_x = (width + 8 * hmax - 1) / (8 * hmax); mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
                    for (int k = 0; k < ncomp; k++) {
                        Component& c = comp[k];
                        c.w = (width * c.h + hmax - 1) / hmax; c.h_px = (height * c.v + vmax - 1) / vmax;
                        c.blocks_w = mcus_x * c.h; c.blocks_h = mcus_y * c.v;
                        c.coef.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
                    }
                    have_frame = true;
                    break;
                }
                case 0xDD: u16(); restart_interval = u16(); break;
                case 0xDA:
                    if (!have_frame) bad("scan before frame header");
                    scan();
                    break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    bad("unsupported JPEG process (lossless / hierarchical / arithmetic)");
                default: {  // APPn, COM, ...: skip
                    int len = u16();
                    if (len < 2 || p + (len - 2) > end) bad("bad segment length");
                    p += len - 2;
                }
            }
            if (p >= end) done = true;
        }
        if (!have_frame) bad("no frame");
        // ---- dequantise + inverse DCT into component planes
        std::vector<std::vector<uint8_t>> plane(ncomp);
        for (int k = 0; k < ncomp; k++) {
            Component& c = comp[k];
            if (!qt_ok[c.tq]) bad("missing quantisation table");
            const int stride = c.blocks_w * 8;
            plane[k].resize((size_t)stride * c.blocks_h * 8);
            for (int by = 0; by < c.blocks_h; by++)
                for (int bx = 0; bx < c.blocks_w; bx++)
                    idct(c.coef.data() + ((size_t)by * c.blocks_w + bx) * 64, qt[c.tq], plane[k].data() + (size_t)by * 8 * stride + bx * 8, stride);
            std::vector<int16_t>().swap(c.coef);
        }
        // ---- colour conversion (IJG jdcolor, 16-bit fixed point); chroma replicated when subsampled
        auto im = std::make_shared<Image>();
        im->width = (uint32_t)width; im->height = (uint32_t)height; im->rgb.resize((size_t)width * height * 3);
        if (ncomp == 1) {
            const int stride = comp[0].blocks_w * 8;
            for (int y = 0; y < height; y++)
                for (int x = 0; x < width; x++) { uint8_t v = plane[0][(size_t)y * stride + x]; uint8_t* o = &im->rgb[((size_t)y * width + x) * 3]; o[0] = o[1] = o[2] = v; }
            return im;
        }
        int upsampled_stride[3] = {0, 0, 0};
        int cr_r[256], cb_b[256]; long cr_g[256], cb_g[256];
        for (int i = 0; i < 256; i++) {
            long x = i - 128;
            cr_r[i] = (int)((91881L * x + 32768L) >> 16);   // FIX(1.40200)
            cb_b[i] = (int)((116130L * x + 32768L) >> 16);  // FIX(1.77200)
            cr_g[i] = -46802L * x;                          // FIX(0.71414)
            cb_g[i] = -22554L * x + 32768L;                 // FIX(0.34414)
        }
        auto clamp8 = [](int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
        // chroma planes sampled 2:1 horizontally (and 1:1 or 2:1 vertically) against a full-resolution luma plane are brought
        // to full resolution with the IJG "fancy" triangle filter (jdsample h2v1 / h2v2); other ratios are replicated
        for (int k = 1; k < 3; k++) {
            Component& c = comp[k];
            if (!(comp[0].h == hmax && comp[0].v == vmax && c.h * 2 == hmax && (c.v == vmax || c.v * 2 == vmax))) continue;
            const int sw = c.w, sh = c.h_px, st = c.blocks_w * 8, ow = sw * 2, oh = c.v == vmax ? sh : sh * 2;
            std::vector<uint8_t> up((size_t)ow * oh);
            std::vector<int> col(sw);
            for (int oy = 0; oy < oh; oy++) {
                const bool v2 = c.v != vmax;
                const int iy = v2 ? oy / 2 : oy;
                int far = v2 ? ((oy & 1) ? iy + 1 : iy - 1) : iy;
                far = far < 0 ? 0 : (far >= sh ? sh - 1 : far);
                const uint8_t* near_row = plane[k].data() + (size_t)iy * st;
                const uint8_t* far_row = plane[k].data() + (size_t)far * st;
                uint8_t* o = up.data() + (size_t)oy * ow;
                if (v2) {
                    for (int x = 0; x < sw; x++) col[x] = 3 * near_row[x] + far_row[x];
                    if (sw == 1) { o[0] = (uint8_t)((col[0] * 4 + 8) >> 4); o[1] = (uint8_t)((col[0] * 4 + 7) >> 4); continue; }
                    o[0] = (uint8_t)((col[0] * 4 + 8) >> 4); o[1] = (uint8_t)((col[0] * 3 + col[1] + 7) >> 4);
                    for (int x = 1; x < sw - 1; x++) { o[2 * x] = (uint8_t)((col[x] * 3 + col[x - 1] + 8) >> 4); o[2 * x + 1] = (uint8_t)((col[x] * 3 + col[x + 1] + 7) >> 4); }
                    o[2 * (sw - 1)] = (uint8_t)((col[sw - 1] * 3 + col[sw - 2] + 8) >> 4); o[2 * sw - 1] = (uint8_t)((col[sw - 1] * 4 + 7) >> 4);
                } else {
                    if (sw == 1) { o[0] = o[1] = near_row[0]; continue; }
                    o[0] = near_row[0]; o[1] = (uint8_t)((near_row[0] * 3 + near_row[1] + 2) >> 2);
                    for (int x = 1; x < sw - 1; x++) { o[2 * x] = (uint8_t)((near_row[x] * 3 + near_row[x - 1] + 1) >> 2); o[2 * x + 1] = (uint8_t)((near_row[x] * 3 + near_row[x + 1] + 2) >> 2); }
                    o[2 * (sw - 1)] = (uint8_t)((near_row[sw - 1] * 3 + near_row[sw - 2] + 1) >> 2); o[2 * sw - 1] = near_row[sw - 1];
                }
            }
            plane[k].swap(up);
            c.blocks_w = (ow + 7) / 8; c.h = hmax; c.v = vmax;  // now a full-resolution plane of stride ow
            upsampled_stride[k] = ow;
        }
        const int s0 = comp[0].blocks_w * 8, s1 = upsampled_stride[1] ? upsampled_stride[1] : comp[1].blocks_w * 8,
                  s2 = upsampled_stride[2] ? upsampled_stride[2] : comp[2].blocks_w * 8;
        for (int y = 0; y < height; y++) {
            const int y0 = y * comp[0].v / vmax, y1 = y * comp[1].v / vmax, y2 = y * comp[2].v / vmax;
            for (int x = 0; x < width; x++) {
                const int Y = plane[0][(size_t)y0 * s0 + x * comp[0].h / hmax];
                const int cb = plane[1][(size_t)y1 * s1 + x * comp[1].h / hmax], cr = plane[2][(size_t)y2 * s2 + x * comp[2].h / hmax];
                uint8_t* o = &im->rgb[((size_t)y * width + x) * 3];
                o[0] = clamp8(Y + cr_r[cr]);
                o[1] = clamp8(Y + (int)((cb_g[cb] + cr_g[cr]) >> 16));
                o[2] = clamp8(Y + cb_b[cb]);
            }
        }
        return im;
    }
};

}  // namespace

ImagePtr decode_jpeg(const std::vector<uint8_t>& file) {
    Decoder d; d.p = file.data(); d.end = file.data() + file.size();
    memset(d.qt, 0, sizeof(d.qt));
    return d.run();
}

}  // namespace pt
