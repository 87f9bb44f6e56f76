// rr_jpeg.cpp — baseline JPEG decoder for RGB8 textures.
//
// The reference loads a material's texture with image::open(..) (render.rs:165-181, image 0.24.2 -> jpeg-decoder 0.2.6) and uses
// it only when the result is DynamicImage::ImageRgb8 (render.rs:251). A three-component YCbCr JPEG decodes to exactly that,
// a grey-scale one to ImageLuma8 (ignored by the path). This file covers what such a texture file normally is: Huffman-coded
// DCT frames, sequential (SOF0 / SOF1) or progressive (SOF2: spectral selection and successive approximation, T.81 annex G),
// 8 bits per sample, 1..4 samples per MCU axis, restart intervals, JFIF YCbCr or Adobe RGB.
// Arithmetic-coded, lossless, hierarchical and 12-bit files return nullptr = "not an RGB8 image", like every other failed load.
//
// Arithmetic follows ITU-T T.81 with the constants of the public-domain integer IDCT (12-bit fixed point, the one jpeg-decoder's
// idct.rs is also derived from), triangle-filter chroma upsampling for 2x1 and 2x2 subsampling (replication otherwise) and
// BT.601 full-range colour conversion. T.81 leaves the IDCT and the upsampling filter to the decoder within one level per
// sample, so texels may differ from another conforming decoder's by a level or two (tests/test_host_cpu.py measures the
// difference against libjpeg-turbo); the render path only ever sees the decoded RGB8 texels.
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "rr_host.hpp"

namespace rr {
namespace {

struct Huff {
    // canonical code tables (T.81 annex C/F): for each length, first code / first symbol index; 9-bit fast look-up
    uint8_t sym[256];
    int mincode[17], maxcode[18], valptr[17];
    int16_t fast[512];  // (len << 8) | symbol, or -1
    bool ok = false;
};

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;  // blocks per row / column (padded to whole MCUs)
    int pred = 0;
    std::vector<uint8_t> px;  // decoded samples, (bw*8) x (bh*8)
    std::vector<int> coef;    // progressive frames: the coefficients of all blocks (natural order), refined scan by scan
    int aw = 0, ah = 0;       // blocks that hold image samples (a non-interleaved scan visits only these)
};

struct BitReader {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool hit_marker = false;
    void fill() {
        while (n <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;          // stuffed zero
                    else { hit_marker = true; b = 0; }                 // a marker: feed zeros from here on
                } else ++p;
            }
            acc |= (uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; n -= k; }
    int get(int k) { if (k == 0) return 0; int v = peek(k); skip(k); return v; }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};

bool build_huff(Huff &h, const uint8_t *counts, const uint8_t *symbols, int nsym) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h.valptr[l] = k;
        h.mincode[l] = code;
        code += counts[l - 1];
        k += counts[l - 1];
        h.maxcode[l] = counts[l - 1] ? code - 1 : -1;
        if (code > (1 << l)) return false;
        code <<= 1;
    }
    h.maxcode[17] = 0x7fffffff;
    if (k != nsym || k > 256) return false;
    memcpy(h.sym, symbols, (size_t)nsym);
    for (int i = 0; i < 512; ++i) h.fast[i] = -1;
    for (int l = 1; l <= 9; ++l)
        for (int i = 0; i < counts[l - 1]; ++i) {
            const int c = h.mincode[l] + i, s = h.sym[h.valptr[l] + i];
            for (int pad = 0; pad < (1 << (9 - l)); ++pad) h.fast[(c << (9 - l)) | pad] = (int16_t)((l << 8) | s);
        }
    h.ok = true;
    return true;
}

int decode_sym(BitReader &br, const Huff &h) {
    const int f = h.fast[br.peek(9)];
    if (f >= 0) { br.skip(f >> 8); return f & 255; }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        const int c = code >> (16 - l);
        if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) {
            br.skip(l);
            return h.sym[h.valptr[l] + c - h.mincode[l]];
        }
    }
    return -1;
}

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }  // T.81 F.2.2.1

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Integer 8x8 inverse DCT, 12-bit fixed-point constants, column pass then row pass, output level-shifted and clamped.
// 64-bit intermediates: a hostile file can carry coefficients (clamped to +-2^20 by the caller) far outside what an
// encoder produces, and the 32-bit sums of the textbook version would overflow on them.
inline int f2f(double x) { return (int)(x * 4096 + 0.5); }
inline uint8_t clamp8(long long x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
#define RR_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                       \
    long long t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                        \
    p2 = s2; p3 = s6;                                                                    \
    p1 = (p2 + p3) * f2f(0.5411961);                                                     \
    t2 = p1 + p3 * f2f(-1.847759065);                                                    \
    t3 = p1 + p2 * f2f(0.765366865);                                                     \
    p2 = s0; p3 = s4;                                                                    \
    t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                        \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                              \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                  \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                              \
    p5 = (p3 + p4) * f2f(1.175875602);                                                   \
    t0 = t0 * f2f(0.298631336); t1 = t1 * f2f(2.053119869);                              \
    t2 = t2 * f2f(3.072711026); t3 = t3 * f2f(1.501321110);                              \
    p1 = p5 + p1 * f2f(-0.899976223); p2 = p5 + p2 * f2f(-2.562915447);                  \
    p3 = p3 * f2f(-1.961570560); p4 = p4 * f2f(-0.390180644);                            \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

void idct_block(const long long *in, uint8_t *out, int stride) {
    long long tmp[64];
    for (int i = 0; i < 8; ++i) {
        const long long *d = in + i;
        long long *v = tmp + i;
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            const long long dc = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
            continue;
        }
        RR_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
        x0 += 512; x1 += 512; x2 += 512; x3 += 512;
        v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
        v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
        v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
        v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
    }
    for (int i = 0; i < 8; ++i) {
        const long long *v = tmp + i * 8;
        uint8_t *o = out + i * stride;
        RR_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}

// ---- progressive scans (T.81 annex G) ----
inline int clampi(long long v) { return (int)(v < -(1ll << 28) ? -(1ll << 28) : (v > (1ll << 28) ? (1ll << 28) : v)); }
bool prog_dc(BitReader &br, Comp &c, int *blk, const Huff &h, int Ah, int Al) {
    if (Ah == 0) {
        const int t = decode_sym(br, h);
        if (t < 0 || t > 11) return false;
        c.pred = clampi((long long)c.pred + (t ? extend(br.get(t), t) : 0));
        blk[0] = clampi((long long)c.pred * (1 << Al));
    } else if (br.get(1)) {
        blk[0] = clampi((long long)blk[0] + (1 << Al));
    }
    return true;
}
bool prog_ac(BitReader &br, int *blk, const Huff &h, int Ss, int Se, int Ah, int Al, int &eobrun) {
    if (Ah == 0) {  // first pass over this band
        if (eobrun) { --eobrun; return true; }
        int k = Ss;
        do {
            const int rs = decode_sym(br, h);
            if (rs < 0) return false;
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r < 15) {
                    eobrun = (1 << r) + (r ? br.get(r) : 0) - 1;
                    break;
                }
                k += 16;
            } else {
                k += r;
                if (k > Se) return false;
                blk[ZIGZAG[k++]] = extend(br.get(sz), sz) * (1 << Al);
            }
        } while (k <= Se);
        return true;
    }
    // refinement: one more bit for the coefficients that are already non-zero, new +-1 coefficients in between
    const int bit = 1 << Al;
    auto refine = [&](int &v) {
        if (br.get(1) && (v & bit) == 0) v = clampi((long long)v + (v > 0 ? bit : -bit));
    };
    if (eobrun) {
        --eobrun;
        for (int k = Ss; k <= Se; ++k) {
            int &v = blk[ZIGZAG[k]];
            if (v != 0) refine(v);
        }
        return true;
    }
    int k = Ss;
    do {
        const int rs = decode_sym(br, h);
        if (rs < 0) return false;
        int r = rs >> 4, sz = rs & 15, val = 0;
        if (sz == 0) {
            if (r < 15) {
                eobrun = (1 << r) - 1 + (r ? br.get(r) : 0);
                r = 64;  // run to the end of the band, refining only
            }
        } else {
            if (sz != 1) return false;
            val = br.get(1) ? bit : -bit;
        }
        while (k <= Se) {
            int &v = blk[ZIGZAG[k++]];
            if (v != 0) {
                refine(v);
            } else {
                if (r == 0) { v = val; break; }
                --r;
            }
        }
    } while (k <= Se);
    return true;
}

inline unsigned be16(const uint8_t *p) { return ((unsigned)p[0] << 8) | p[1]; }

// one output row of `w` samples from a component plane, upsampled to full resolution
void upsample_row(const Comp &c, int hmax, int vmax, int y, int w, uint8_t *out) {
    const int sw = c.bw * 8, sh = c.bh * 8;
    const int hs = hmax / c.h, vs = vmax / c.v;
    const int cw = (w + hs - 1) / hs;  // meaningful source samples in a row
    const bool whole = hmax % c.h == 0 && vmax % c.v == 0;  // (hostile headers may carry ratios like 4:3)
    if (whole && hs == 1 && vs == 1) {
        memcpy(out, &c.px[(size_t)y * sw], (size_t)w);
        return;
    }
    if (whole && hs == 2 && (vs == 1 || vs == 2)) {
        // triangle filter: 3/4 nearer + 1/4 farther sample per axis (vertical blend first for 2x2), rounding as libjpeg's
        // "fancy" upsampling does: +8 >> 4 on even and +7 >> 4 on odd outputs for 2x2, +1 / +2 >> 2 for 2x1
        const uint8_t *near_row, *far_row = nullptr;
        if (vs == 2) {
            const int sy = y >> 1;
            int fy = (y & 1) ? sy + 1 : sy - 1;
            const int rows = (sh < 1 ? 1 : sh);
            if (fy < 0) fy = 0;
            if (fy >= rows) fy = rows - 1;
            near_row = &c.px[(size_t)sy * sw];
            far_row = &c.px[(size_t)fy * sw];
        } else {
            near_row = &c.px[(size_t)y * sw];
        }
        std::vector<int> v((size_t)cw);
        for (int i = 0; i < cw; ++i) v[i] = vs == 2 ? 3 * near_row[i] + far_row[i] : near_row[i];
        for (int x = 0; x < w; ++x) {
            const int i = x >> 1;
            int j = (x & 1) ? i + 1 : i - 1;
            if (j < 0) j = 0;
            if (j >= cw) j = cw - 1;
            if (vs == 2) out[x] = (uint8_t)((3 * v[i] + v[j] + ((x & 1) ? 7 : 8)) >> 4);
            else out[x] = (uint8_t)((3 * v[i] + v[j] + ((x & 1) ? 2 : 1)) >> 2);
        }
        return;
    }
    // any other ratio: sample replication
    const int sy = (int)((long long)y * c.v / vmax);
    const uint8_t *row = &c.px[(size_t)(sy < sh ? sy : sh - 1) * sw];
    for (int x = 0; x < w; ++x) {
        const int sx = (int)((long long)x * c.h / hmax);
        out[x] = row[sx < sw ? sx : sw - 1];
    }
}

}  // namespace

std::shared_ptr<TextureRgb8> load_jpeg_rgb8(const std::vector<uint8_t> &d) {
    if (d.size() < 4 || d[0] != 0xFF || d[1] != 0xD8) return nullptr;
    uint16_t qt[4][64];
    bool qt_ok[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    Comp comp[3];
    int ncomp = 0, W = 0, H = 0, hmax = 1, vmax = 1, restart = 0, mcux = 0, mcuy = 0, scans = 0;
    bool have_frame = false, progressive = false, adobe = false;
    int adobe_transform = -1;
    auto lim = [](long long v) { return v < -(1ll << 20) ? -(1ll << 20) : (v > (1ll << 20) ? (1ll << 20) : v); };

    // planes -> RGB8 (upsampling + colour conversion)
    auto assemble = [&]() {
        auto tex = std::make_shared<TextureRgb8>();
        tex->width = (uint32_t)W; tex->height = (uint32_t)H;
        tex->rgb8.resize((size_t)W * H * 3);
        // JFIF: YCbCr. Adobe marker with transform 0: the three components ARE R, G, B.
        const bool is_rgb = adobe && adobe_transform == 0;
        std::vector<uint8_t> r0((size_t)W), r1((size_t)W), r2((size_t)W);
        for (int y = 0; y < H; ++y) {
            upsample_row(comp[0], hmax, vmax, y, W, r0.data());
            upsample_row(comp[1], hmax, vmax, y, W, r1.data());
            upsample_row(comp[2], hmax, vmax, y, W, r2.data());
            uint8_t *o = &tex->rgb8[(size_t)y * W * 3];
            for (int x = 0; x < W; ++x) {
                if (is_rgb) { o[3 * x] = r0[x]; o[3 * x + 1] = r1[x]; o[3 * x + 2] = r2[x]; continue; }
                // BT.601 full range, 16-bit fixed point with rounding (ITU-T T.871)
                const int Y = r0[x] << 16, cb = r1[x] - 128, cr = r2[x] - 128;
                o[3 * x] = clamp8((Y + 91881 * cr + 32768) >> 16);
                o[3 * x + 1] = clamp8((Y - 22554 * cb - 46802 * cr + 32768) >> 16);
                o[3 * x + 2] = clamp8((Y + 116130 * cb + 32768) >> 16);
            }
        }
        return tex;
    };
    // progressive frames: all scans are in, dequantise and transform every block
    auto finish_progressive = [&]() -> std::shared_ptr<TextureRgb8> {
        long long blk[64];
        for (int c = 0; c < ncomp; ++c) {
            Comp &cc = comp[c];
            if (!qt_ok[cc.tq]) return nullptr;
            for (int by = 0; by < cc.bh; ++by)
                for (int bx = 0; bx < cc.bw; ++bx) {
                    const int *src = &cc.coef[((size_t)by * cc.bw + bx) * 64];
                    for (int k = 0; k < 64; ++k) blk[k] = lim((long long)src[k] * qt[cc.tq][k]);
                    idct_block(blk, &cc.px[(size_t)by * 8 * cc.bw * 8 + (size_t)bx * 8], cc.bw * 8);
                }
        }
        return assemble();
    };
    // entropy-coded segment ends at the next marker that is not a restart marker
    auto next_marker = [&](const uint8_t *from) {
        const uint8_t *q = from, *end = d.data() + d.size();
        while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
        return (size_t)(q - d.data());
    };
    // RSTn between restart intervals: byte-align and step over it
    auto take_restart = [&](BitReader &br, int &next_rst) {
        br.reset();
        const uint8_t *q = br.p;
        while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
        if (q + 1 >= br.end || q[1] != 0xD0 + next_rst) return false;
        br.p = q + 2;
        next_rst = (next_rst + 1) & 7;
        for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
        return true;
    };

    size_t p = 2;
    while (p + 4 <= d.size()) {
        if (d[p] != 0xFF) { ++p; continue; }
        const int m = d[p + 1];
        if (m == 0xFF) { ++p; continue; }
        p += 2;
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) break;
        if (p + 2 > d.size()) return nullptr;
        const size_t len = be16(&d[p]);
        if (len < 2 || p + len > d.size()) return nullptr;
        const uint8_t *b = &d[p + 2];
        const size_t n = len - 2;
        if (m == 0xDB) {  // DQT
            for (size_t i = 0; i < n;) {
                const int pq = b[i] >> 4, tq = b[i] & 15;
                if (tq > 3 || pq > 1) return nullptr;
                ++i;
                if (i + (pq ? 128 : 64) > n) return nullptr;
                for (int k = 0; k < 64; ++k) {
                    qt[tq][ZIGZAG[k]] = pq ? (uint16_t)be16(&b[i + 2 * k]) : b[i + k];
                }
                i += pq ? 128 : 64;
                qt_ok[tq] = true;
            }
        } else if (m == 0xC4) {  // DHT
            for (size_t i = 0; i + 17 <= n;) {
                const int tc = b[i] >> 4, th = b[i] & 15;
                if (tc > 1 || th > 3) return nullptr;
                int total = 0;
                for (int k = 0; k < 16; ++k) total += b[i + 1 + k];
                if (i + 17 + (size_t)total > n) return nullptr;
                if (!build_huff(tc ? ac[th] : dc[th], &b[i + 1], &b[i + 17], total)) return nullptr;
                i += 17 + (size_t)total;
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // SOF0 / SOF1: sequential, SOF2: progressive; Huffman
            if (have_frame || n < 6 || b[0] != 8) return nullptr;
            progressive = m == 0xC2;
            H = (int)be16(&b[1]); W = (int)be16(&b[3]); ncomp = b[5];
            if (W <= 0 || H <= 0) return nullptr;
            if (ncomp != 3) return nullptr;  // 1 component = ImageLuma8 (not Rgb8), 4 = CMYK: both ignored by the path
            if (n < 6 + 3 * (size_t)ncomp) return nullptr;
            for (int c = 0; c < ncomp; ++c) {
                comp[c].id = b[6 + 3 * c];
                comp[c].h = b[7 + 3 * c] >> 4; comp[c].v = b[7 + 3 * c] & 15;
                comp[c].tq = b[8 + 3 * c];
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return nullptr;
                hmax = comp[c].h > hmax ? comp[c].h : hmax;
                vmax = comp[c].v > vmax ? comp[c].v : vmax;
            }
            // untrusted header: refuse absurd sizes before allocating (at least ~1 bit per 8x8 block must follow; a
            // progressive frame keeps all its coefficients in memory, 4 bytes each)
            const unsigned long long blocks = ((unsigned long long)W + 7) / 8 * (((unsigned long long)H + 7) / 8);
            if ((unsigned long long)W * H > (progressive ? 1ull << 25 : 1ull << 28) || blocks / 8 > d.size()) return nullptr;
            mcux = (W + 8 * hmax - 1) / (8 * hmax); mcuy = (H + 8 * vmax - 1) / (8 * vmax);
            for (int c = 0; c < ncomp; ++c) {
                Comp &cc = comp[c];
                cc.bw = mcux * cc.h; cc.bh = mcuy * cc.v; cc.pred = 0;
                cc.aw = (int)((((long long)W * cc.h + hmax - 1) / hmax + 7) / 8);
                cc.ah = (int)((((long long)H * cc.v + vmax - 1) / vmax + 7) / 8);
                cc.px.assign((size_t)cc.bw * 8 * cc.bh * 8, 0);
                if (progressive) cc.coef.assign((size_t)cc.bw * cc.bh * 64, 0);
            }
            have_frame = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return nullptr;  // lossless / arithmetic / differential frames
        } else if (m == 0xDD) {
            if (n < 2) return nullptr;
            restart = (int)be16(b);
        } else if (m == 0xEE) {
            if (n >= 12 && memcmp(b, "Adobe", 5) == 0) { adobe = true; adobe_transform = b[11]; }
        } else if (m == 0xDA && progressive) {  // one of the scans of a progressive frame
            if (!have_frame || n < 1) return nullptr;
            const int ns = b[0];
            if (ns < 1 || ns > ncomp || n < 1 + 2 * (size_t)ns + 3) return nullptr;
            int sc[3];
            for (int s2 = 0; s2 < ns; ++s2) {
                const int cid = b[1 + 2 * s2];
                int c = -1;
                for (int k = 0; k < ncomp; ++k) if (comp[k].id == cid) c = k;
                if (c < 0 || (s2 > 0 && c <= sc[s2 - 1])) return nullptr;
                sc[s2] = c;
                comp[c].td = b[2 + 2 * s2] >> 4; comp[c].ta = b[2 + 2 * s2] & 15;
                if (comp[c].td > 3 || comp[c].ta > 3) return nullptr;
            }
            const int Ss = b[1 + 2 * ns], Se = b[2 + 2 * ns], Ah = b[3 + 2 * ns] >> 4, Al = b[3 + 2 * ns] & 15;
            if (Ss > Se || Se > 63 || Ah > 13 || Al > 13 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1)) return nullptr;
            const bool is_dc = Ss == 0;
            for (int s2 = 0; s2 < ns; ++s2) {
                const Comp &cc = comp[sc[s2]];
                if (is_dc ? (Ah == 0 && !dc[cc.td].ok) : !ac[cc.ta].ok) return nullptr;
            }
            BitReader br{&d[p + len], d.data() + d.size()};
            int todo = restart, next_rst = 0, eobrun = 0;
            for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
            bool ok = true;
            if (ns == 1) {  // non-interleaved: the component's own blocks in raster order
                Comp &cc = comp[sc[0]];
                for (int by = 0; ok && by < cc.ah; ++by)
                    for (int bx = 0; ok && bx < cc.aw; ++bx) {
                        if (restart && todo == 0) {
                            if (!take_restart(br, next_rst)) return nullptr;
                            todo = restart; eobrun = 0;
                        }
                        int *blk = &cc.coef[((size_t)by * cc.bw + bx) * 64];
                        ok = is_dc ? prog_dc(br, cc, blk, dc[cc.td], Ah, Al) : prog_ac(br, blk, ac[cc.ta], Ss, Se, Ah, Al, eobrun);
                        if (restart) --todo;
                    }
            } else {        // interleaved DC scan: MCU order
                for (int my = 0; ok && my < mcuy; ++my)
                    for (int mx = 0; ok && mx < mcux; ++mx) {
                        if (restart && todo == 0) {
                            if (!take_restart(br, next_rst)) return nullptr;
                            todo = restart;
                        }
                        for (int s2 = 0; ok && s2 < ns; ++s2) {
                            Comp &cc = comp[sc[s2]];
                            for (int by = 0; ok && by < cc.v; ++by)
                                for (int bx = 0; ok && bx < cc.h; ++bx)
                                    ok = prog_dc(br, cc, &cc.coef[((size_t)(my * cc.v + by) * cc.bw + (mx * cc.h + bx)) * 64], dc[cc.td], Ah, Al);
                        }
                        if (restart) --todo;
                    }
            }
            if (!ok) return nullptr;
            ++scans;
            p = next_marker(br.p);
            continue;
        } else if (m == 0xDA) {  // SOS: the one scan of a sequential frame (interleaved)
            if (!have_frame || n < 1 || b[0] != ncomp || n < 1 + 2 * (size_t)ncomp + 3) return nullptr;
            for (int s = 0; s < ncomp; ++s) {
                const int cid = b[1 + 2 * s];
                int c = -1;
                for (int k = 0; k < ncomp; ++k) if (comp[k].id == cid) c = k;
                if (c != s) return nullptr;  // components in frame order
                comp[c].td = b[2 + 2 * s] >> 4; comp[c].ta = b[2 + 2 * s] & 15;
                if (comp[c].td > 3 || comp[c].ta > 3 || !dc[comp[c].td].ok || !ac[comp[c].ta].ok || !qt_ok[comp[c].tq]) return nullptr;
            }
            for (int c = 0; c < ncomp; ++c) comp[c].pred = 0;
            BitReader br{&d[p + len], d.data() + d.size()};
            long long coef[64];
            int todo = restart, next_rst = 0;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart && todo == 0) {
                        if (!take_restart(br, next_rst)) return nullptr;
                        todo = restart;
                    }
                    for (int c = 0; c < ncomp; ++c) {
                        Comp &cc = comp[c];
                        for (int by = 0; by < cc.v; ++by)
                            for (int bx = 0; bx < cc.h; ++bx) {
                                memset(coef, 0, sizeof coef);
                                int t = decode_sym(br, dc[cc.td]);
                                if (t < 0 || t > 11) return nullptr;
                                cc.pred = (int)lim((long long)cc.pred + (t ? extend(br.get(t), t) : 0));
                                coef[0] = lim((long long)cc.pred * qt[cc.tq][0]);
                                for (int k = 1; k < 64;) {
                                    const int rs = decode_sym(br, ac[cc.ta]);
                                    if (rs < 0) return nullptr;
                                    const int r = rs >> 4, s = rs & 15;
                                    if (s == 0) {
                                        if (r != 15) break;  // EOB
                                        k += 16;
                                        continue;
                                    }
                                    k += r;
                                    if (k > 63) return nullptr;
                                    coef[ZIGZAG[k]] = lim((long long)extend(br.get(s), s) * qt[cc.tq][ZIGZAG[k]]);
                                    ++k;
                                }
                                const size_t ox = (size_t)(mx * cc.h + bx) * 8, oy = (size_t)(my * cc.v + by) * 8;
                                idct_block(coef, &cc.px[oy * (size_t)cc.bw * 8 + ox], cc.bw * 8);
                            }
                    }
                    if (restart) --todo;
                }
            return assemble();
        }
        p += len;
    }
    if (progressive && have_frame && scans > 0) return finish_progressive();
    return nullptr;
}

// image::open() equivalent for the formats a texture normally comes in: PNG and baseline JPEG, told apart by their signatures.
std::shared_ptr<TextureRgb8> load_image_rgb8(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return nullptr;
    uint8_t sig[2] = {0, 0};
    const size_t got = fread(sig, 1, 2, f);
    if (got == 2 && sig[0] == 0xFF && sig[1] == 0xD8) {
        std::vector<uint8_t> d(sig, sig + 2);
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) {
            d.insert(d.end(), buf, buf + n);
            if (d.size() > (1u << 30)) { fclose(f); return nullptr; }
        }
        fclose(f);
        return load_jpeg_rgb8(d);
    }
    fclose(f);
    return load_png_rgb8(path);
}

}  // namespace rr
