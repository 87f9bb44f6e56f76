// Mutation fuzzer of the baseline JPEG texture decoder (host/rr_jpeg.cpp). Build with -fsanitize=address,undefined:
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=all -Iray-rust_b200/host -Iinclude tools/fuzz/fuzz_jpeg.cpp ray-rust_b200/host/rr_jpeg.cpp ray-rust_b200/host/rr_png.cpp -lz -pthread -o /tmp/fuzz_jpeg && /tmp/fuzz_jpeg seed1.jpg seed2.jpg ...
// FUZZ_SEED / FUZZ_N choose the RNG seed and the number of mutated files (default 30 000). Runs over baseline and progressive
// PIL-written seeds found (and the decoder now avoids) a 32-bit IDCT overflow and a row overrun with non-integer sampling ratios.
#include "rr_host.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char **argv) {
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 1; i < argc; ++i) {
        FILE *f = fopen(argv[i], "rb"); if (!f) continue;
        std::vector<uint8_t> d; uint8_t b[65536]; size_t n;
        while ((n = fread(b, 1, sizeof b, f)) > 0) d.insert(d.end(), b, b + n);
        fclose(f); seeds.push_back(d);
    }
    std::mt19937 rng(getenv("FUZZ_SEED") ? (unsigned)atol(getenv("FUZZ_SEED")) : 12345u);
    const int iters = getenv("FUZZ_N") ? atoi(getenv("FUZZ_N")) : 30000;
    long ok = 0, total = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> d = seeds[rng() % seeds.size()];
        int kind = rng() % 4;
        if (kind == 0) d.resize(rng() % (d.size() + 1));
        int flips = 1 + rng() % 8;
        for (int k = 0; k < flips && !d.empty(); ++k) {
            size_t pos = (rng() % 3 == 0) ? rng() % std::min<size_t>(d.size(), 700) : rng() % d.size();
            d[pos] = (kind == 2) ? (uint8_t)rng() : d[pos] ^ (1u << (rng() % 8));
        }
        auto t = rr::load_jpeg_rgb8(d);
        ++total; if (t) { ++ok; if (t->rgb8.size() != (size_t)t->width * t->height * 3) { printf("size mismatch\n"); return 1; } }
    }
    printf("fuzz: %ld of %ld mutated files still decoded, no crash\n", ok, total);
    return 0;
}
