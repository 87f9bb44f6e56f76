"""`ray-rust W H -w -p PORT` (host/rr_web.cpp), the device-backed mirror of webserver.rs:22-333.
CPU: routes, status codes, query parsing rules, and that /render has no CPU fallback. GPU: /render frames equal the oracle's
render of the same camera (webserver.rs:268-274 camera patch), including concurrent requests on the one resident scene."""
import io
import os
import socket
import subprocess
import time
import urllib.error
import urllib.request
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "ray-rust_b200", "host")
CLI = os.path.join(HOST, "ray-rust")
W, H = 160, 120


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.fixture(scope="module")
def server(tmp_path_factory):
    subprocess.check_call(["make", "-C", HOST, "-s"])
    port = _free_port()
    cwd = tmp_path_factory.mktemp("web")
    log = open(cwd / "server.log", "w")
    proc = subprocess.Popen([CLI, str(W), str(H), "-w", "-p", str(port)], stdout=log, stderr=subprocess.STDOUT, cwd=cwd)
    base = f"http://127.0.0.1:{port}"
    for _ in range(600):
        try:
            urllib.request.urlopen(base + "/nothing", timeout=1)
        except urllib.error.HTTPError:
            break
        except OSError:
            assert proc.poll() is None, open(cwd / "server.log").read()
            time.sleep(0.1)
    yield base, cwd
    proc.kill()
    proc.wait()
    log.close()


def _get(url):
    try:
        r = urllib.request.urlopen(url, timeout=60)
        return r.status, dict(r.headers), r.read()
    except urllib.error.HTTPError as e:
        return e.code, dict(e.headers), e.read()


def test_routes(server):
    base, cwd = server
    st, hd, body = _get(base + "/")
    assert st == 200 and hd["Content-Type"] == "text/html"
    page = body.decode()
    # start pose = the scene's camera (webserver.rs:71-83): default camera (0,-150,-300), pyr = (-pi/2 rolled into yaw) ...
    assert "x:0,y:-150,z:-300" in page and "/render?x=" in page
    for key in ("ArrowRight", "ArrowLeft", "ArrowUp", "ArrowDown", "w:", "s:", "a:", "d:", "q:", "z:"):
        assert key in page
    assert _get(base + "/image")[2] == b"image"                      # no barb.png in the cwd (webserver.rs:219-221)
    (cwd / "barb.png").write_bytes(b"\x89PNG fake")
    assert _get(base + "/image")[2] == b"\x89PNG fake"
    (cwd / "barb.png").unlink()
    st, _, body = _get(base + "/nosuch")
    assert st == 404 and body == b"empty"                             # webserver.rs:300-304
    log = open(cwd / "server.log").read()
    assert "Listening on http://0.0.0.0:" in log and "Got request at /image" in log


def test_render_without_gpu_fails_loudly(server):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    st, _, body = _get(server[0] + "/render?x=0&y=-150&z=-300&yaw=-90&pitch=0")
    assert st == 500 and body.startswith(b"fail to render")


def _camera_env(rr, x, y, z, yaw, pitch):
    """What webserver.rs:268-274 does to its clone of the RenderEnv."""
    from ray_rust_b200.scene import Camera

    ren = rr.default_scene(W, H)
    f32 = np.float32
    pyr = list(ren.camera.pyr)
    pyr[1] = f32(yaw) * f32(np.pi) / f32(180.0)
    pyr[0] = f32(pitch) * f32(np.pi) / f32(180.0)
    ren.camera = Camera((x, y, z), pyr)
    return ren


@pytest.mark.gpu
def test_render_matches_oracle(server, rr, oracle):
    from PIL import Image

    base, cwd = server
    cams = [(0.0, -150.0, -300.0, -90.0, 0.0), (40.0, -120.0, -250.0, -75.0, 10.0), (0.0, 0.0, 0.0, 0.0, 0.0), (-200.0, -50.0, 100.0, 30.0, -20.0)]
    urls = [base + f"/render?x={c[0]}&y={c[1]}&z={c[2]}&yaw={c[3]}&pitch={c[4]}" for c in cams]
    urls.append(base + "/render?x=abc&y=-150&z=-300&yaw=-90&bogus=1&pitch=0=0")  # unparsable / malformed pairs stay 0
    cams.append((0.0, -150.0, -300.0, -90.0, 0.0))
    with ThreadPoolExecutor(8) as ex:                                 # concurrent requests share the resident scene
        got = list(ex.map(_get, urls * 2))
    for (st, hd, body), cam in zip(got, cams * 2):
        assert st == 200 and hd["Content-Type"] == "image/png" and hd["Cache-Control"] == "no-cache"
        dev = np.asarray(Image.open(io.BytesIO(body)).convert("RGB"))
        assert dev.shape == (H, W, 3)
        ref = oracle.render(_camera_env(rr, *cam))["u8"]
        d = np.abs(dev.astype(int) - ref.astype(int)).max(axis=2)
        assert (d <= 1).mean() >= 0.9995 and (d == 0).mean() >= 0.998, (cam, (d == 0).mean(), d.max())
    assert "Rendering with xpos=40, ypos=-120, zpos=-250, yaw=-75 pitch=10" in open(cwd / "server.log").read()
