// rr_web.cpp — the interactive web path of webserver.rs:22-333 on the device: `ray-rust W H -w [-p PORT]`.
//   GET /                                   -> HTML page with key controls (own page, same behaviour: it re-requests /render)
//   GET /render?x=&y=&z=&yaw=&pitch=        -> PNG of the scene from that camera (webserver.rs:222-299)
//   GET /image                              -> barb.png if it exists, else the text "image" (webserver.rs:209-221)
//   anything else                           -> 404 "empty"
// The reference clones the whole RenderEnv per request (webserver.rs:268) and renders on a tokio worker; here every
// request only patches the camera fields of rr_frame_params and renders on the ONE resident scene handle, which is
// safe to call from concurrent connection threads: the C ABI runs up to four renders of one handle at a time on separate
// lanes (stream pair + device frame), and each request renders into a page-locked frame from a small pool.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>

#include "rr_host.hpp"

namespace rr {
namespace {

// Page with the reference's controls (webserver.rs:66-208): W/S forward/back, A/D strafe, Q/Z up/down, arrows turn by 5
// degrees; keys held down keep stepping, one /render request in flight at a time; start pose = the scene's camera.
std::string page_html(const RenderEnv &ren) {
    char init[256];
    snprintf(init, sizeof init, "var cam={x:%.9g,y:%.9g,z:%.9g,yaw:%.9g,pitch:%.9g};\n", ren.camera.position.x, ren.camera.position.y,
             ren.camera.position.z, ren.camera.pyr.y * 180.0f / PI, ren.camera.pyr.x * 180.0f / PI);
    std::string s =
        "<html><head><title>ray-rust</title><script>\n";
    s += init;
    s +=
        "var held={},busy=false;\n"
        "var moves={w:[1,0,0,0,0],s:[-1,0,0,0,0],a:[0,1,0,0,0],d:[0,-1,0,0,0],q:[0,0,10,0,0],z:[0,0,-10,0,0],\n"
        " ArrowRight:[0,0,0,5,0],ArrowLeft:[0,0,0,-5,0],ArrowUp:[0,0,0,0,-5],ArrowDown:[0,0,0,0,5]};\n"
        "function show(){document.getElementById('label').innerHTML='x='+cam.x+'<br>y='+cam.y+'<br>z='+cam.z+'<br>yaw='+cam.yaw+'<br>pitch='+cam.pitch;}\n"
        "function request(){busy=true;show();\n"
        " fetch('/render?x='+cam.x+'&y='+cam.y+'&z='+cam.z+'&yaw='+cam.yaw+'&pitch='+cam.pitch)\n"
        "  .then(function(r){if(r.ok)return r.blob();throw new Error(r.status);})\n"
        "  .then(function(b){document.getElementById('render').src=URL.createObjectURL(b);busy=false;step();})\n"
        "  .catch(function(e){busy=false;console.log('render request failed: '+e.message);});}\n"
        "function step(){if(busy)return;var moved=false,c=Math.cos(cam.yaw*Math.PI/180),n=Math.sin(cam.yaw*Math.PI/180);\n"
        " for(var k in moves){if(!held[k])continue;var m=moves[k];moved=true;\n"
        "  cam.x+=10*(m[0]*c+m[1]*n);cam.z+=10*(m[1]*c-m[0]*n);cam.y+=m[2];cam.yaw+=m[3];cam.pitch+=m[4];}\n"
        " if(moved)request();}\n"
        "window.onload=function(){request();\n"
        " window.onkeydown=function(e){if(e.key in moves){if(!held[e.key]){held[e.key]=true;step();}e.preventDefault();}};\n"
        " window.onkeyup=function(e){if(e.key in moves){held[e.key]=false;e.preventDefault();}};};\n"
        "</script><style>table{border-collapse:collapse;border:solid;}</style></head><body>\n"
        "<h1>ray-rust web interface</h1><img id='render'><hr><h2>Controls</h2><table border='1'>\n"
        "<tr><td>W</td><td>forward</td></tr><tr><td>S</td><td>backward</td></tr><tr><td>A</td><td>left</td></tr>\n"
        "<tr><td>D</td><td>right</td></tr><tr><td>Q</td><td>up</td></tr><tr><td>Z</td><td>down</td></tr>\n"
        "<tr><td>Left arrow</td><td>Turn left</td></tr><tr><td>Right arrow</td><td>Turn right</td></tr>\n"
        "<tr><td>Up arrow</td><td>Turn up</td></tr><tr><td>Down arrow</td><td>Turn down</td></tr></table>\n"
        "<hr><h2>Debug</h2><div id='label'></div></body></html>\n";
    return s;
}

void send_all(int fd, const void *buf, size_t n) {
    const char *p = (const char *)buf;
    while (n > 0) {
        ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
        if (k <= 0) return;
        p += k;
        n -= (size_t)k;
    }
}
void respond(int fd, int code, const char *status, const char *ctype, const void *body, size_t n, bool no_cache = false) {
    char head[256];
    int h = snprintf(head, sizeof head, "HTTP/1.1 %d %s\r\nContent-Type: %s\r\nContent-Length: %zu\r\n%sConnection: close\r\n\r\n", code, status,
                     ctype, n, no_cache ? "Cache-Control: no-cache\r\n" : "");
    send_all(fd, head, (size_t)h);
    send_all(fd, body, n);
}

// Page-locked frames shared by the connection threads (at most as many as requests are in flight at once).
struct FramePool {
    std::mutex mu;
    std::vector<std::unique_ptr<PinnedFrame>> free_;
    std::unique_ptr<PinnedFrame> take(size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!free_.empty()) {
                auto f = std::move(free_.back());
                free_.pop_back();
                return f;
            }
        }
        try { return std::unique_ptr<PinnedFrame>(new PinnedFrame(bytes ? bytes : 1)); } catch (...) { return nullptr; }
    }
    void give(std::unique_ptr<PinnedFrame> f) {
        std::lock_guard<std::mutex> lk(mu);
        if (free_.size() < 8) free_.push_back(std::move(f));
    }
};

struct Server {
    const RenderEnv *ren;
    int width, height, device;
    rr_scene *handle;
    FramePool *pool;
};

void handle_conn(int fd, const Server *srv) {
    char buf[4096];
    ssize_t n = ::recv(fd, buf, sizeof buf - 1, 0);
    if (n <= 0) { ::close(fd); return; }
    buf[n] = 0;
    char method[8] = {0}, target[2048] = {0};
    if (sscanf(buf, "%7s %2047s", method, target) != 2) { ::close(fd); return; }
    std::string uri(target), path = uri, query;
    size_t qpos = uri.find('?');
    if (qpos != std::string::npos) { path = uri.substr(0, qpos); query = uri.substr(qpos + 1); }
    printf("Got request at %s\n", target);
    if (uri == "/") {
        const std::string page = page_html(*srv->ren);
        respond(fd, 200, "OK", "text/html", page.data(), page.size());
    } else if (uri == "/image") {
        std::ifstream f("barb.png", std::ios::binary);
        if (f) {
            std::stringstream ss;
            ss << f.rdbuf();
            const std::string s = ss.str();
            printf("Responding with image %zu\n", s.size());
            respond(fd, 200, "OK", "image/png", s.data(), s.size());
        } else {
            respond(fd, 200, "OK", "text/plain", "image", 5);
        }
    } else if (path == "/render") {
        printf("GET /render, query = %s\n", query.c_str());
        float xpos = 0, ypos = 0, zpos = 0, yaw = 0, pitch = 0;  // webserver.rs:224-260: unparsable values stay 0
        std::stringstream qs(query);
        std::string kv;
        while (std::getline(qs, kv, '&')) {
            size_t eq = kv.find('=');
            if (eq == std::string::npos || kv.find('=', eq + 1) != std::string::npos) continue;
            const std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
            char *end = nullptr;
            const float f = strtof(v.c_str(), &end);
            if (end == v.c_str() || *end) continue;
            if (k == "x") xpos = f; else if (k == "y") ypos = f; else if (k == "z") zpos = f;
            else if (k == "yaw") yaw = f; else if (k == "pitch") pitch = f;
        }
        printf("Rendering with xpos=%g, ypos=%g, zpos=%g, yaw=%g pitch=%g\n", xpos, ypos, zpos, yaw, pitch);
        // webserver.rs:269-274: position, pyr.y = yaw*PI/180, pyr.x = pitch*PI/180, rotation = from_pyr(pyr)
        Vec3 pyr = srv->ren->camera.pyr;
        pyr.y = yaw * PI / 180.0f;
        pyr.x = pitch * PI / 180.0f;
        const Quat rot = Quat::from_pyr(pyr);
        rr_frame_params p = srv->ren->frame_params();
        p.xres = srv->width; p.yres = srv->height;
        p.cam_position[0] = xpos; p.cam_position[1] = ypos; p.cam_position[2] = zpos;
        p.cam_rotation[0] = rot.x; p.cam_rotation[1] = rot.y; p.cam_rotation[2] = rot.z; p.cam_rotation[3] = rot.w;
        // a page-locked frame from the server's pool (reused across requests): asynchronous D2H at the full PCIe rate
        std::unique_ptr<PinnedFrame> data = srv->pool->take((size_t)3 * srv->width * srv->height);
        const bool ok = data && rr_render_rgb8(srv->handle, &p, data->data(), 0) == RR_OK;
        if (ok) {
            std::vector<uint8_t> png = encode_png_rgb8(data->data(), (uint32_t)srv->width, (uint32_t)srv->height);
            srv->pool->give(std::move(data));
            respond(fd, 200, "OK", "image/png", png.data(), png.size(), true);
        } else {
            if (data) srv->pool->give(std::move(data));
            const std::string msg = std::string("fail to render: ") + rr_last_error();
            respond(fd, 500, "Internal Server Error", "text/plain", msg.data(), msg.size());
        }
    } else {
        respond(fd, 404, "Not Found", "text/plain", "empty", 5);
    }
    ::shutdown(fd, SHUT_RDWR);
    ::close(fd);
}

}  // namespace

// run_webserver, webserver.rs:324-333. Blocks forever (until the process is killed), like the reference.
int run_webserver(const RenderEnv &ren, int width, int height, int port, int device) {
    FramePool pool;
    Server srv{&ren, width, height, device, nullptr, &pool};
    FlatScene flat = flatten(ren);
    rr_scene_desc desc = flat.desc();
    if (rr_scene_create(&desc, device, &srv.handle) != RR_OK)
        fprintf(stderr, "warning: no usable CUDA device (%s); /render will answer 500\n", rr_last_error());
    int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) { perror("socket"); return 1; }
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in addr;
    memset(&addr, 0, sizeof addr);
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);  // 0.0.0.0, webserver.rs:325
    addr.sin_port = htons((uint16_t)port);
    if (::bind(ls, (sockaddr *)&addr, sizeof addr) < 0 || ::listen(ls, 64) < 0) { perror("bind/listen"); ::close(ls); return 1; }
    printf("Listening on http://0.0.0.0:%d\n", port);
    fflush(stdout);
    for (;;) {
        int fd = ::accept(ls, nullptr, nullptr);
        if (fd < 0) continue;
        std::thread(handle_conn, fd, &srv).detach();
    }
}

}  // namespace rr
