// rr_png.cpp — PNG output (image::save_buffer(.., Rgb8), main.rs:325-340) and texture input
// (image::open for RenderMaterial.texture, render.rs:165-181, 215) on top of zlib.
// Writer: 8-bit RGB, non-interlaced, filter 1 (Sub), one zlib stream deflated in parallel stripes. Reader: returns an image only
// when it decodes to 8-bit RGB (colour type 2, or a palette without transparency, which the image
// crate expands to Rgb8); everything else yields nullptr because the path honours only
// DynamicImage::ImageRgb8 (render.rs:251).
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "rr_host.hpp"

namespace rr {
namespace {

void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
void chunk(std::vector<uint8_t> &out, const char *type, const uint8_t *data, size_t n) {
    put32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}
uint32_t get32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

std::vector<uint8_t> encode_png_rgb8(const uint8_t *rgb, uint32_t w, uint32_t h) {
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13];
    ihdr[0] = w >> 24; ihdr[1] = w >> 16; ihdr[2] = w >> 8; ihdr[3] = w;
    ihdr[4] = h >> 24; ihdr[5] = h >> 16; ihdr[6] = h >> 8; ihdr[7] = h;
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk(out, "IHDR", ihdr, 13);
    // The frame is cut into row stripes that are deflated concurrently (pigz scheme): every stripe is
    // a raw-deflate stream ended with Z_SYNC_FLUSH (byte aligned, not final), the last one with
    // Z_FINISH; concatenated behind one zlib header and closed with the combined Adler-32 they form a
    // single valid zlib stream. Once the render takes < 1 ms the serial deflate of a 25-100 MB frame is
    // what the reference's "Rendering time" (main.rs:316-348 includes save_buffer) would consist of.
    const size_t row = (size_t)w * 3;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)32, (size_t)std::max<uint32_t>(1, h / 64)});
    if (const char *e = getenv("RR_PNG_THREADS")) nt = std::max(1, atoi(e));
    // four stripes per thread, handed out by a counter: sky rows deflate several times faster than rows full of spheres,
    // equal row counts per thread would leave most threads waiting for the slowest stripe
    const int ns = (int)std::min<uint32_t>((uint32_t)nt * 4u, std::max<uint32_t>(1, h / 16));
    struct Stripe {
        std::vector<uint8_t> z;
        uLong adler = 1, len = 0;
        bool ok = false;
    };
    std::vector<Stripe> stripes(ns);
    std::atomic<int> next{0};
    auto work = [&]() {
        std::unique_ptr<uint8_t[]> raw;  // not zero-filled
        size_t raw_cap = 0;
        for (int k; (k = next.fetch_add(1)) < ns;) {
            const uint32_t y0 = (uint32_t)((uint64_t)h * k / ns), y1 = (uint32_t)((uint64_t)h * (k + 1) / ns);
            const size_t raw_len = (row + 1) * (y1 - y0);
            if (raw_len > raw_cap) { raw.reset(new uint8_t[raw_len]); raw_cap = raw_len; }
            for (uint32_t y = y0; y < y1; ++y) {
                // filter type 1 (Sub: byte - byte of the pixel to the left). On rendered frames zlib level 1 then runs 1.5x
                // faster and the file is 40 % smaller than with filter 0 (measured: 4K default scene, 143 vs 96 MB/s per
                // thread, 9.7 % vs 16.1 % of the raw size); the filter itself is one pass fused into the copy.
                uint8_t *o = &raw[(row + 1) * (y - y0)];
                const uint8_t *in = rgb + row * y;
                o[0] = 1;
                for (size_t i = 0; i < row && i < 3; ++i) o[1 + i] = in[i];
                for (size_t i = 3; i < row; ++i) o[1 + i] = (uint8_t)(in[i] - in[i - 3]);
            }
            Stripe &s = stripes[k];
            s.len = (uLong)raw_len;
            s.adler = adler32(1L, raw.get(), (uInt)raw_len);
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) continue;
            s.z.resize(deflateBound(&zs, (uLong)raw_len) + 16);
            zs.next_in = raw.get(); zs.avail_in = (uInt)raw_len;
            zs.next_out = s.z.data(); zs.avail_out = (uInt)s.z.size();
            const int rc = deflate(&zs, k == ns - 1 ? Z_FINISH : Z_SYNC_FLUSH);
            s.ok = (k == ns - 1) ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0);
            s.z.resize(zs.total_out);
            deflateEnd(&zs);
        }
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nt; ++k) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    const int nt_total = ns;
    std::vector<uint8_t> z = {0x78, 0x01};  // zlib header: deflate, 32K window, fastest
    uLong adler = 1;
    size_t zbytes = 6;
    for (int k = 0; k < nt_total; ++k) zbytes += stripes[k].z.size();
    z.reserve(zbytes);
    for (int k = 0; k < nt_total; ++k) {
        if (!stripes[k].ok) throw std::runtime_error("png: deflate failed");
        z.insert(z.end(), stripes[k].z.begin(), stripes[k].z.end());
        adler = k == 0 ? stripes[k].adler : adler32_combine(adler, stripes[k].adler, (z_off_t)stripes[k].len);
    }
    z.push_back(adler >> 24); z.push_back(adler >> 16); z.push_back(adler >> 8); z.push_back(adler);
    chunk(out, "IDAT", z.data(), z.size());
    chunk(out, "IEND", nullptr, 0);
    return out;
}

void save_png_rgb8(const std::string &path, const uint8_t *rgb, uint32_t w, uint32_t h) {
    std::vector<uint8_t> png = encode_png_rgb8(rgb, w, h);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    size_t n = fwrite(png.data(), 1, png.size(), f);
    fclose(f);
    if (n != png.size()) throw std::runtime_error("short write to " + path);
}

std::shared_ptr<TextureRgb8> load_png_rgb8(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return nullptr;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (d.size() < 8 || memcmp(d.data(), sig, 8) != 0) return nullptr;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool trns = false;
    for (size_t p = 8; p + 12 <= d.size();) {
        uint32_t len = get32(&d[p]);
        if (p + 12 + len > d.size()) return nullptr;
        const char *t = (const char *)&d[p + 4];
        const uint8_t *body = &d[p + 8];
        if (!memcmp(t, "IHDR", 4) && len >= 13) {
            w = get32(body); h = get32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!memcmp(t, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(t, "tRNS", 4)) trns = true;
        else if (!memcmp(t, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(t, "IEND", 4)) break;
        p += 12 + len;
    }
    if (w == 0 || h == 0 || interlace != 0) return nullptr;
    int channels;
    if (ctype == 2 && depth == 8 && !trns) channels = 3;
    else if (ctype == 3 && depth == 8 && !trns && !plte.empty()) channels = 1;
    else return nullptr;  // Luma / RGBA / 16-bit: not ImageRgb8, the path would ignore it
    const size_t stride = (size_t)w * channels;
    // The header comes from a file named by an untrusted scene YAML: refuse sizes the IDAT data cannot possibly inflate
    // to (deflate expands at most ~1032:1) or that are absurd for a texture, BEFORE allocating (nullptr = "not an RGB8
    // image", like every other failed load).
    const unsigned long long need = ((unsigned long long)stride + 1) * h;
    if (need > (1ull << 31) || need > (unsigned long long)idat.size() * 1040ull + 65536ull) return nullptr;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return nullptr;
    std::vector<uint8_t> px(stride * h);
    const int bpp = channels;
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *in = &raw[(stride + 1) * y + 1];
        uint8_t *cur = &px[stride * y];
        const uint8_t *up = y ? &px[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
            int v = in[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: return nullptr;
            }
            cur[i] = (uint8_t)v;
        }
    }
    auto tex = std::make_shared<TextureRgb8>();
    tex->width = w; tex->height = h;
    if (channels == 3) {
        tex->rgb8 = std::move(px);
    } else {
        tex->rgb8.resize((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; ++i) {
            size_t k = (size_t)px[i] * 3;
            if (k + 2 >= plte.size()) return nullptr;
            tex->rgb8[3 * i] = plte[k]; tex->rgb8[3 * i + 1] = plte[k + 1]; tex->rgb8[3 * i + 2] = plte[k + 2];
        }
    }
    return tex;
}

}  // namespace rr
