// Mutation fuzzer of load_image_rgb8 (PNG reader + JPEG decoder behind the signature dispatch); build like fuzz_jpeg.cpp.
// Writes each mutated file to /tmp/fz/m.bin (create /tmp/fz first).
#include "rr_host.hpp"
#include <cstdio>
#include <random>
int main(int argc, char **argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc; ++i) {
        FILE *f = fopen(argv[i], "rb"); if (!f) continue;
        std::vector<uint8_t> d; uint8_t b[65536]; size_t n;
        while ((n = fread(b, 1, sizeof b, f)) > 0) d.insert(d.end(), b, b + n);
        fclose(f); seeds.push_back(d);
    }
    std::mt19937 rng(777);
    long ok = 0, total = 0;
    for (int it = 0; it < 6000; ++it) {
        std::vector<uint8_t> d = seeds[rng() % seeds.size()];
        int kind = rng() % 4;
        if (kind == 0) d.resize(rng() % (d.size() + 1));
        int flips = 1 + rng() % 6;
        for (int k = 0; k < flips && !d.empty(); ++k) {
            size_t pos = (rng() % 2 == 0) ? rng() % std::min<size_t>(d.size(), 64) : rng() % d.size();
            d[pos] = (kind == 2) ? (uint8_t)rng() : d[pos] ^ (1u << (rng() % 8));
        }
        FILE *f = fopen("/tmp/fz/m.bin", "wb"); fwrite(d.data(), 1, d.size(), f); fclose(f);
        auto t = rr::load_image_rgb8("/tmp/fz/m.bin");
        ++total; if (t) { ++ok; if (t->rgb8.size() != (size_t)t->width * t->height * 3) { printf("size mismatch\n"); return 1; } }
    }
    printf("fuzz png/dispatch: %ld of %ld mutated files still decoded, no crash\n", ok, total);
}
