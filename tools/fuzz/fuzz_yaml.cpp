// Mutation fuzzer of RenderEnv::deserialize (host/rr_yaml.cpp): 40 000 mutated scene documents under ASan/UBSan, no report.
// Build: g++ -O1 -g -std=c++17 -fsanitize=address,undefined -ffp-contract=off -Iray-rust_b200/host -Iinclude tools/fuzz/fuzz_yaml.cpp ray-rust_b200/host/rr_{host,yaml,jpeg,png,web}.cpp -Lray-rust_b200 -lray_rust_b200 -Wl,-rpath,$PWD/ray-rust_b200 -lz -pthread
#include "rr_host.hpp"
#include <cstdio>
#include <random>
#include <string>
int main() {
    rr::RenderEnv env = rr::default_scene(64, 48, false, false, 0.0f);
    std::string base = env.serialize();
    std::mt19937 rng(4242);
    long ok = 0, total = 0;
    for (int it = 0; it < 40000; ++it) {
        std::string d = base;
        int kind = rng() % 5;
        if (kind == 0) d.resize(rng() % (d.size() + 1));
        int flips = 1 + rng() % 6;
        for (int k = 0; k < flips && !d.empty(); ++k) {
            size_t pos = rng() % d.size();
            if (kind == 1) d[pos] = (char)(rng() % 96 + 32);
            else if (kind == 2) d.insert(pos, std::string(1 + rng() % 4, " -:[\n#'\"0e"[rng() % 10]));
            else if (kind == 3) d.erase(pos, 1 + rng() % 12);
            else d[pos] ^= (char)(1u << (rng() % 7));
        }
        rr::RenderEnv e2 = rr::default_scene(8, 8, false, false, 0.0f);
        try { e2.deserialize(d); ++ok; (void)e2.serialize(); } catch (const std::exception &) {}
        ++total;
    }
    printf("fuzz yaml: %ld of %ld mutated documents still parsed, no crash\n", ok, total);
}
