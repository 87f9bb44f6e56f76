// main.cpp — the `ray-rust` command line of main.rs:31-350 on the device path.
//   ray-rust <width> <height> [-t N] [-o out.png] [-m] [-g G] [-s scene.yaml] [-d scene.yaml] [--gpu D]
// Same positional arguments, flags, defaults and console output as the reference binary. `-t` is
// parsed and echoed but has no effect (the GPU grid replaces the row-scheduler threads). `-w/-p`
// start the web front-end of webserver.rs on the resident device scene (rr_web.cpp).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "rr_host.hpp"

static void usage() {
    fprintf(stderr,
            "ray-rust (B200 device path)\n\nUSAGE:\n    ray-rust [OPTIONS] <width> <height>\n\nOPTIONS:\n"
            "    -t, --threads <threads>                   thread count [default: 8] (ignored by the device path)\n"
            "    -o, --output <output>                     Output file name [default: foo.png]\n"
            "    -m, --raymarch                            Use ray marching\n"
            "    -g, --gloweffect <gloweffect>             Enable glow effect and set its strength when ray marching method is used\n"
            "    -s, --serialize_file <serialize_file>     File name for serialized scene output\n"
            "    -d, --deserialize_file <deserialize_file> File name for deserialized scene input\n"
            "    -w, --webserver                           Launch a web server that responds with rendered images\n"
            "    -p, --port_no <port_no>                   [default: 3000]\n"
            "        --gpu <index>                         CUDA device [default: 0]\n");
}

int main(int argc, char **argv) {
    std::vector<std::string> pos;
    std::string threads = "8", output = "foo.png", glow, ser, deser, port = "3000", gpu = "0";
    bool raymarch = false, webserver = false, have_glow = false, have_ser = false, have_deser = false, have_gpu = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&](std::string &dst) {
            if (i + 1 >= argc) { usage(); exit(2); }
            dst = argv[++i];
        };
        if (a == "-t" || a == "--threads") val(threads);
        else if (a == "-o" || a == "--output") val(output);
        else if (a == "-m" || a == "--raymarch") raymarch = true;
        else if (a == "-g" || a == "--gloweffect") { val(glow); have_glow = true; }
        else if (a == "-s" || a == "--serialize_file") { val(ser); have_ser = true; }
        else if (a == "-d" || a == "--deserialize_file") { val(deser); have_deser = true; }
        else if (a == "-w" || a == "--webserver") webserver = true;
        else if (a == "-p" || a == "--port_no") val(port);
        else if (a == "--gpu") { val(gpu); have_gpu = true; }
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { usage(); return 2; }
    // parser(): "Value for <name>: <value>" (main.rs:95-105)
    const long width = atol(pos[0].c_str()), height = atol(pos[1].c_str());
    if (width < 0 || height < 0) { fprintf(stderr, "Parsing width/height failed\n"); return 101; }
    printf("Value for width: %ld\n", width);
    printf("Value for height: %ld\n", height);
    const int thread_count = atoi(threads.c_str());
    printf("Value for threads: %d\n", thread_count);
    printf("Value for output: %s\n", output.c_str());
    float glow_value = 0.0f;
    if (have_glow) {
        char *end = nullptr;
        glow_value = strtof(glow.c_str(), &end);
        if (end == glow.c_str() || *end) have_glow = false;  // parser_opt: a failed parse is None
        else printf("Value for gloweffect: %g\n", glow_value);
    }
    if (have_ser) printf("Value for serialize_file: %s\n", ser.c_str());
    if (have_deser) printf("Value for deserialize_file: %s\n", deser.c_str());

    try {
        rr::RenderEnv ren = rr::default_scene((int)width, (int)height, raymarch, have_glow, glow_value);
        if (have_deser) {  // main.rs:278-295
            std::ifstream f(deser);
            if (!f) { fprintf(stderr, "Error: No such file or directory (os error 2)\n"); return 1; }
            std::stringstream ss;
            ss << f.rdbuf();
            ren.deserialize(ss.str());
        }
        if (webserver) {  // main.rs:297-309
            printf("Value for port_no: %s\n", port.c_str());
            setvbuf(stdout, nullptr, _IOLBF, 0);
            return rr::run_webserver(ren, (int)width, (int)height, atoi(port.c_str()), atoi(gpu.c_str()));
        }
        if (have_ser) {  // main.rs:311-314
            std::ofstream f(ser, std::ios::binary);
            if (!f) { fprintf(stderr, "Error: cannot create %s\n", ser.c_str()); return 1; }
            f << ren.serialize();
        }
        const int device = atoi(gpu.c_str());
        // Like the reference, the scene exists before the clock starts (main.rs:154-316 builds RenderEnv first): here that
        // includes its device copy (CUDA context + upload) and the page-locked frame. Reported on stderr, stdout stays
        // the reference's.
        auto setup0 = std::chrono::steady_clock::now();
        rr::PinnedFrame data;
        if (ren.camera_motion.empty()) {
            rr::upload_scene(ren, device);
            data.resize((size_t)3 * width * height);
        }
        fprintf(stderr, "device setup: %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - setup0).count());
        auto start = std::chrono::steady_clock::now();  // main.rs:316: the timed region includes the PNG encode
        if (!ren.camera_motion.empty()) {
            // frames are dealt to every visible GPU unless --gpu pins one (render.rs:926-989, pipelined)
            rr::render_frames(ren, (size_t)width, (size_t)height, [&](int i, const uint8_t *data, size_t) {
                try { rr::save_png_rgb8(output + std::to_string(i) + ".png", data, (uint32_t)width, (uint32_t)height); } catch (...) {}
            }, thread_count, have_gpu ? std::vector<int>{device} : std::vector<int>{});
        } else {
            rr::render_rgb8(ren, data.data(), device);  // `data` is page-locked: asynchronous D2H at full PCIe rate
            rr::save_png_rgb8(output, data.data(), (uint32_t)width, (uint32_t)height);
        }
        auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - start).count();
        printf("Rendering time: %lld.%06lld\n", (long long)(us / 1000000), (long long)(us % 1000000));
    } catch (const rr::DeserializeError &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
