"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/rr_ffi.h
declares, the ctypes mirror matches the header, and the host-side scene logic is right.
No compute call is made here (there is no GPU in the authoring container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "rr_ffi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(rr):
    lib = rr.ffi.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in rr_ffi.h but not exported"
    assert sorted(rr.ffi.PROTOTYPES) == syms
    assert lib.rr_abi_version() == 2


def test_integration_sys_block_lists_every_symbol():
    """INTEGRATION.md's Rust `-sys` block (the binding a maintainer would add) declares every entry point of the header,
    and its #[repr(C)] structs carry every field of the C structs in the same order."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rust = sorted(set(re.findall(r"pub fn (rr_[a-z0-9_]+)\(", doc)))
    assert rust == _header_symbols()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rr_ffi.h")).read(), flags=re.S)
    for st in ["rr_material", "rr_object", "rr_texture", "rr_scene_desc", "rr_frame_params", "rr_ray_counts"]:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (st, st), hdr, flags=re.S).group(1)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
            c_fields += [re.sub(r"[\*\s]|\[.*?\]", "", n) for n in names.split(",")]
        rbody = re.search(r"pub struct %s \{(.*?)\n?\}" % st, doc, flags=re.S).group(1)
        r_fields = re.findall(r"pub ([a-z0-9_]+):", rbody)
        assert r_fields == c_fields, st


def test_struct_layout_matches_header(rr, tmp_path):
    """Compile a C probe against include/rr_ffi.h and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess

    f = rr.ffi
    structs = ["rr_material", "rr_object", "rr_texture", "rr_scene_desc", "rr_frame_params", "rr_ray_counts"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rr_ffi.h"', "int main(void){"]
    for st in structs:
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for fld, _ in getattr(f, st)._fields_:
            lines.append(f'printf("{st}.{fld} %zu\\n", offsetof({st}, {fld}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for st in structs:
        cls = getattr(f, st)
        assert int(got[st]) == C.sizeof(cls), st
        for fld, _ in cls._fields_:
            assert int(got[f"{st}.{fld}"]) == getattr(cls, fld).offset, f"{st}.{fld}"


def test_missing_library_fails_loudly(rr, monkeypatch):
    f = rr.ffi
    monkeypatch.setattr(f, "_lib", None)
    monkeypatch.setattr(f, "LIB_PATH", "/nonexistent/libray_rust_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        f.load()


def test_library_is_built_from_the_sources_next_to_it(rr):
    """Built files are not in git history but travel to the GPU box: the hash of csrc/ + include/rr_ffi.h + the nvcc flags
    is compiled into the library, and the loader rebuilds a binary whose hash differs (ray-rust_b200/build.py)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("rr_build", os.path.join(ROOT, "ray-rust_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib = rr.ffi.load()
    info = lib.rr_build_info().decode()
    assert info.startswith("rr_src_hash=") and "sm_100a" in info and "fmad=false" in info
    assert info.split()[0] == "rr_src_hash=" + b.source_hash() == "rr_src_hash=" + b.embedded_hash()
    assert b.source_hash(["-DX"]) != b.source_hash()


def test_error_convention_without_device(rr):
    """Without a GPU every entry point must fail with a status code and a message, not crash."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = rr.ffi.load()
    n = C.c_int(-1)
    rc = lib.rr_device_count(C.byref(n))
    assert rc in (rr.ffi.RR_ERR_CUDA, rr.ffi.RR_OK)
    ren = rr.default_scene(8, 8)
    with pytest.raises(rr.ffi.RrError):
        rr.DeviceScene(ren, 0)
    assert lib.rr_scene_create(None, 0, None) == rr.ffi.RR_ERR_BAD_ARG
    assert b"null" in lib.rr_last_error()


def test_default_scene_matches_main_rs(rr):
    ren = rr.default_scene(640, 480)
    flat = ren.flatten()
    d = flat.desc
    assert d.n_objects == 5 and d.n_materials == 4 and d.n_textures == 0  # bar.png is absent
    assert d.objects[0].kind == rr.ffi.RR_FLOOR and d.objects[0].uvmap == rr.ffi.RR_UV_ZX
    assert [d.objects[i].r for i in range(1, 5)] == [80.0, 80.0, 80.0, 100.0]
    assert d.objects[1].material == d.objects[2].material  # the shared mirror material
    p = ren.frame_params()
    assert p.yfov == np.float32(480) / np.float32(640) and p.xfov == 1.0
    light = np.array(list(p.light), dtype=np.float32)
    assert abs(float(np.sqrt((light.astype(np.float64) ** 2).sum())) - 1.0) < 1e-6
    q = np.array(list(p.cam_rotation), dtype=np.float64)
    assert abs((q ** 2).sum() - 1.0) < 1e-6


def test_from_pyr_matches_oracle(rr, oracle):
    lib = oracle.load()
    rng = np.random.default_rng(7)
    for _ in range(50):
        pyr = rng.uniform(-3.2, 3.2, 3).astype(np.float32)
        out = (C.c_float * 4)()
        lib.oracle_quat_from_pyr(oracle.fa(*pyr), out)
        q = rr.Quat.from_pyr(tuple(pyr)).as_tuple()
        assert [np.float32(x) for x in out] == list(q)


def test_yaml_roundtrip(rr):
    ren = rr.synthetic_scene(64, 48, n_spheres=20)
    text = ren.serialize()
    ren2 = rr.default_scene(64, 48)
    ren2.deserialize(text)
    a, b = ren.flatten().desc, ren2.flatten().desc
    assert a.n_objects == b.n_objects == 21
    for i in range(a.n_objects):
        oa, ob = a.objects[i], b.objects[i]
        assert (oa.kind, oa.uvmap, oa.r, list(oa.org), list(oa.face_normal)) == (ob.kind, ob.uvmap, ob.r, list(ob.org), list(ob.face_normal))
        ma, mb = a.materials[oa.material], b.materials[ob.material]
        for fld, _ in rr.ffi.rr_material._fields_:
            va, vb = getattr(ma, fld), getattr(mb, fld)
            assert (list(va) == list(vb)) if hasattr(va, "__len__") else (va == vb)
    # the serializer writes the constants, not the env's values (render.rs:742-743)
    assert "max_reflections: 3" in text and "max_refractions: 10" in text and "camera_motion: []" in text


def test_yaml_missing_material_error(rr):
    ren = rr.default_scene(8, 8)
    text = ren.serialize().replace("material: red", "material: nosuch")
    with pytest.raises(rr.DeserializeError, match="RenderSphere couldn't find material nosuch"):
        rr.default_scene(8, 8).deserialize(text)
    with pytest.raises(rr.DeserializeError, match="serde_yaml::Error"):
        rr.default_scene(8, 8).deserialize("camera: 3")


def test_synthetic_scene_is_deterministic(rr):
    a = rr.synthetic_scene(32, 32).flatten().desc
    b = rr.synthetic_scene(32, 32).flatten().desc
    assert a.n_objects == 1025 and a.n_materials == 17
    assert bytes(C.string_at(a.objects, C.sizeof(rr.ffi.rr_object) * 1025)) == bytes(
        C.string_at(b.objects, C.sizeof(rr.ffi.rr_object) * 1025))
    assert a.objects[0].kind == rr.ffi.RR_FLOOR  # index 0 ends the bounce loop (render.rs:1187)


def test_frame_rows(rr):
    ren = rr.default_scene(16, 70)
    tot = 0
    for k in range(3):
        p = ren.frame_params(band_rows=8, band_index=k, band_count=3)
        n = C.c_int32()
        assert rr.ffi.load().rr_frame_rows(C.byref(p), C.byref(n)) == 0
        assert n.value == rr.frame_rows(p)
        tot += n.value
    assert tot == 70
    tot = 0
    for idx, span in ((0, 3), (3, 2), (5, 2)):  # unequal spans of a 7-slot period
        p = ren.frame_params(band_rows=4, band_index=idx, band_count=7, band_span=span)
        n = C.c_int32()
        assert rr.ffi.load().rr_frame_rows(C.byref(p), C.byref(n)) == 0
        assert n.value == rr.frame_rows(p)
        tot += n.value
    assert tot == 70
    bad = ren.frame_params(band_rows=4, band_index=5, band_count=7, band_span=3)  # span runs past the period
    assert rr.ffi.load().rr_frame_rows(C.byref(bad), C.byref(n)) == rr.ffi.RR_ERR_BAD_ARG
