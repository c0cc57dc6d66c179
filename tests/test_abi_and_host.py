"""CPU-side checks of the boundary: the shared library loads and exports every symbol include/t3d_particles.h
declares, fails loudly without a GPU, and its pure-host helpers (the scalar logic of the path) agree with the
oracle.  No compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cpu_harness as H
from tractor3d_b200 import abi, presets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "t3d_particles.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    names = declared_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/t3d_particles.h but not exported"
    assert set(names) == set(abi.PROTOTYPES), set(names) ^ set(abi.PROTOTYPES)
    assert lib.t3d_abi_version() == 4


def test_params_struct_layout_matches_header_defaults():
    lib = abi.load()
    p = abi.EmitterParams()
    lib.t3d_emitter_params_default(C.byref(p))
    assert p.as_dict() == abi.default_params().as_dict()
    assert C.sizeof(abi.EmitterParams) == 264  # sizeof(t3d_emitter_params) as compiled by gcc


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    h = C.c_void_p()
    rc = lib.t3d_ctx_create(0, None, C.byref(h))
    assert rc == abi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.t3d_last_error()


def test_device_rng_spec_matches_oracle_restatement():
    lib, O = abi.load(), H.oracle()
    rng = np.random.default_rng(0)
    for _ in range(2000):
        seed, serial, j = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**40)), int(rng.integers(0, 64))
        a, b = lib.t3d_device_rng_draw(seed, serial, j), O.device_rng_draw(seed, serial, j)
        assert a == b and 0 <= a < 2**31


def test_rate_cap_and_emission_count_follow_the_oracle():
    # Drive the oracle's update() with awkward dts and replay the same scalar logic through the host helpers.
    lib = abi.load()
    for rate, dts in ((300, [16.6667] * 50), (1000, [16.6667] * 50), (1200, [16.6667] * 40), (37, [3.0, 5.5, 2.1, 9.9] * 30),
                      (999, [7.9, 0.2, 33.3] * 30)):
        O = H.oracle(fresh=True)
        p, _ = presets.smoke()
        p.emission_rate = rate
        p.particle_count_max = 1_000_000
        p.energy_min = p.energy_max = 1e9
        eo = O.emitter(p)
        eo.start()
        running, emit_time = C.c_double(0.0), C.c_float(0.0)
        tpe = np.float32(1000.0) / np.float32(rate)
        total = 0
        for dt in dts:
            eo.update(dt)
            ms = C.c_float()
            if lib.t3d_host_rate_cap(C.byref(running), dt, C.byref(ms)):
                total += lib.t3d_host_emission_count(C.byref(emit_time), float(tpe), ms.value)
            assert total == eo.count(), (rate, dt)
            assert np.float32(emit_time.value) == np.float32(eo.emit_time())


def test_fuse_policy_follows_the_measurements():
    # profiles/r2_churn.txt: with moves folded into the stream the fused frame wins from 0 to 7 % of the particles replaced
    # per frame, so update() always waits for the draw() behind it
    lib = abi.load()
    for particles, spawned in ((64 << 20, 0), (38_500_000, 131_072), (19_300_000, 262_144), (14_300_000, 1 << 20), (0, 0), (0, 1)):
        assert lib.t3d_host_fuse_policy(particles, spawned) == 1


def test_sprite_frame_coords_follow_the_oracle():
    lib, O = abi.load(), H.oracle()
    for tw, th, n, w, h in ((1024, 256, 3, 256, 256), (256, 256, 16, 64, 64), (64, 64, 1, 64, 64), (300, 200, 7, 100, 50),
                            (512, 512, 40, 128, 128)):
        p = abi.default_params(texture_width=tw, texture_height=th)
        eo = O.emitter(p)
        eo.set_sprite_frame_coords(n, w, h)
        out = np.zeros(n * 4, dtype=np.float32)
        lib.t3d_host_sprite_frame_coords(n, w, h, tw, th, out.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(out.view(np.uint32), eo.sprite_tex_coords().reshape(-1).view(np.uint32))


def test_particle_draws_cuts_a_recorded_stream_like_the_oracle():
    lib, O = abi.load(), H.oracle(fresh=True)
    for ellipsoid, fro in ((0, 0), (0, 3), (1, 0), (1, 4)):
        p, _ = presets.fire()
        p.ellipsoid, p.sprite_frame_random_offset = ellipsoid, fro
        O.rand_seed(100 + ellipsoid * 10 + fro)
        eo = O.emitter(p)
        lens = []
        stream = []
        for k in range(200):
            O.rand_record(1)
            eo.emit_once(1)
            d = O.recorded()
            lens.append(len(d)); stream.append(d)
        allv = np.concatenate(stream).astype(np.int32)
        pos = 0
        for k in range(200):
            n = lib.t3d_host_particle_draws(allv[pos:].ctypes.data_as(C.POINTER(C.c_int32)), len(allv) - pos, ellipsoid, fro)
            assert n == lens[k]
            pos += n
        assert pos == len(allv)
        if not ellipsoid:
            assert set(lens) == {26 + (1 if fro else 0)}
        else:
            assert min(lens) == 26 + (1 if fro else 0) and max(lens) > min(lens)


PARTICLE_TEXT = """
// a test emitter
particle sparks
{
    sprite
    {
        path = sheet.png
        width = 32
        height = 16
        blendMode = MULTIPLIED
        animated = true
        looped = true
        frameCount = 6
        frameRandomOffset = 9
        frameDuration = 33.7
    }

    particleCountMax = 777
    emissionRate = 55
    ellipsoid = true
    orbitPosition = false
    orbitVelocity = true
    sizeStartMin = 0.25
    sizeStartMax = 1e1
    energyMin = 250
    energyMax = 1250
    colorStart = 1, 0.5, 0.25, 1
    colorEnd = 0, 0, 0, 0.125
    position = 1, 2, 3
    velocityVar = 0.5, 0.5, 0.5
    /* block
       comment */
    rotationSpeedMin = -2
    rotationSpeedMax = 2.5
    rotationAxis = 0, 1, 0
    rotationPerParticleSpeedMax = 3
}
"""


def _write_png(path, w, h):
    import struct
    import zlib
    raw = b"".join(b"\x00" + b"\x00" * (w * 4) for _ in range(h))
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    open(path, "wb").write(png)


def _parse(lib, path):
    p = abi.EmitterParams()
    fc, sw, sh = C.c_uint32(), C.c_int(), C.c_int()
    tex = C.create_string_buffer(1024)
    rc = lib.t3d_particle_file_parse(path.encode(), C.byref(p), C.byref(fc), C.byref(sw), C.byref(sh), tex, 1024)
    return rc, p, fc.value, sw.value, sh.value, tex.value.decode()


def test_particle_file_parse_and_save_round_trip(tmp_path):
    lib = abi.load()
    _write_png(str(tmp_path / "sheet.png"), 96, 32)
    f = tmp_path / "sparks.particle"
    f.write_text(PARTICLE_TEXT)
    rc, p, fc, sw, sh, tex = _parse(lib, str(f))
    assert rc == abi.OK
    assert (fc, sw, sh) == (6, 32, 16) and tex.endswith("sheet.png")
    assert (p.texture_width, p.texture_height) == (96.0, 32.0)
    assert p.particle_count_max == 777 and p.emission_rate == 55 and p.ellipsoid == 1
    assert p.orbit_position == 0 and p.orbit_velocity == 1 and p.orbit_acceleration == 0
    assert p.sprite_animated == 1 and p.sprite_looped == 1
    assert p.sprite_frame_random_offset == 6          # min(frameRandomOffset, frameCount), PE.cpp:128
    assert p.sprite_frame_duration == 33              # float -> long, PE.cpp:129/206
    assert p.blend_mode == abi.BLEND_MULTIPLIED
    assert (p.size_start_min, p.size_start_max, p.size_end_min) == (0.25, 10.0, 0.0)   # missing key -> 0
    assert (p.energy_min, p.energy_max) == (250.0, 1250.0)
    assert list(p.color_start) == [1, 0.5, 0.25, 1] and list(p.color_end) == [0, 0, 0, 0.125]
    assert list(p.position) == [1, 2, 3] and list(p.velocity_var) == [0.5, 0.5, 0.5] and list(p.velocity) == [0, 0, 0]
    assert (p.rotation_speed_min, p.rotation_speed_max) == (-2.0, 2.5) and list(p.rotation_axis) == [0, 1, 0]
    assert p.rotation_per_particle_speed_max == 3.0
    # save (the sample's writer omits rotationSpeed*/rotationAxis*, ParticlesSample.cpp:340-424) and re-parse
    out = tmp_path / "saved.particle"
    assert lib.t3d_particle_file_save(str(out).encode(), b"sparks", C.byref(p), fc, sw, sh, b"sheet.png") == abi.OK
    rc, q, fc2, sw2, sh2, _ = _parse(lib, str(out))
    assert rc == abi.OK and (fc2, sw2, sh2) == (fc, sw, sh)
    dp, dq = p.as_dict(), q.as_dict()
    for k in ("rotation_speed_min", "rotation_speed_max", "rotation_axis", "rotation_axis_var"):
        dp.pop(k); dq.pop(k)
    assert dp == dq


def test_particle_file_errors(tmp_path):
    lib = abi.load()
    rc, *_ = _parse(lib, str(tmp_path / "missing.particle"))
    assert rc == abi.ERR_IO
    bad = tmp_path / "bad.particle"
    bad.write_text("material m\n{\n}\n")
    assert _parse(lib, str(bad))[0] == abi.ERR_IO          # namespace must be `particle` (PE.cpp:94)
    nosprite = tmp_path / "nosprite.particle"
    nosprite.write_text("particle p\n{\n    emissionRate = 3\n}\n")
    assert _parse(lib, str(nosprite))[0] == abi.ERR_IO     # PE.cpp:101-106


@pytest.mark.skipif(not (H.ref_available() and os.path.isdir("/root/reference/samples/browser/res/common/particles")),
                    reason="needs the reference tree and oracle/_ref")
@pytest.mark.parametrize("name", ["fire", "smoke", "explosion"])
def test_stock_particle_files_parse_like_the_reference(name):
    # The reference's own Properties parser (running inside oracle/_ref) and ours must agree on the stock files,
    # and both must agree with the presets the benchmarks use.
    lib, R = abi.load(), H.ref()
    path = f"/root/reference/samples/browser/res/common/particles/{name}.particle"
    rc, p, fc, sw, sh, _ = _parse(lib, path)
    assert rc == abi.OK
    h = R.emitter_create_from_file(path.encode())
    er = H.CpuEmitter(R, handle=h)
    assert p.as_dict() == er.params().as_dict()
    preset, sprite = getattr(presets, name)()
    assert preset.as_dict() == p.as_dict() and sprite == (fc, sw, sh)
    table = np.zeros(fc * 4, dtype=np.float32)
    lib.t3d_host_sprite_frame_coords(fc, sw, sh, p.texture_width, p.texture_height, table.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(table.view(np.uint32), er.sprite_tex_coords().reshape(-1).view(np.uint32))


def test_properties_tree_is_what_create_from_properties_reads(tmp_path):
    # t3d_properties: the slice of tractor::Properties the path reads.  A tree loaded from a file and the same tree built
    # by hand (what a caller holding its own parsed scene file would do, SceneLoader.cpp:318-324) give the parameters the
    # file parser gives -- and those are pinned to the reference's own Properties class by test_oracle_vs_ref.
    lib = abi.load()
    _write_png(str(tmp_path / "sheet.png"), 96, 32)
    f = tmp_path / "sparks.particle"
    f.write_text(PARTICLE_TEXT)
    rc, want, fc, sw, sh, tex = _parse(lib, str(f))
    assert rc == abi.OK

    def parse_tree(node):
        p = abi.EmitterParams()
        n, w, h = C.c_uint32(), C.c_int(), C.c_int()
        buf = C.create_string_buffer(1024)
        rc = lib.t3d_particle_properties_parse(node, C.byref(p), C.byref(n), C.byref(w), C.byref(h), buf, 1024)
        return rc, p, n.value, w.value, h.value, buf.value.decode()

    root = C.c_void_p()
    assert lib.t3d_properties_load(str(f).encode(), C.byref(root)) == abi.OK
    try:
        assert lib.t3d_properties_namespace(root) == b"" and lib.t3d_properties_namespace_count(root) == 1
        particle = lib.t3d_properties_get_namespace(root, 0)
        assert lib.t3d_properties_namespace(particle) == b"particle" and lib.t3d_properties_id(particle) == b"sparks"
        sprite = lib.t3d_properties_get_namespace(particle, 0)
        assert lib.t3d_properties_namespace(sprite) == b"sprite"
        # typed getters, Properties.cpp's conversions
        assert lib.t3d_properties_get_int(sprite, b"frameCount") == 6 and lib.t3d_properties_get_int(sprite, b"missing") == 0
        assert lib.t3d_properties_get_float(sprite, b"frameDuration") == np.float32(33.7)
        assert lib.t3d_properties_get_bool(sprite, b"animated", 0) == 1 and lib.t3d_properties_get_bool(sprite, b"missing", 1) == 1
        assert lib.t3d_properties_get_long(particle, b"energyMax") == 1250
        assert lib.t3d_properties_get_float(particle, b"sizeStartMax") == 10.0
        v = (C.c_float * 4)()
        assert lib.t3d_properties_get_vector(particle, b"colorStart", v, 4) == 1 and list(v) == [1, 0.5, 0.25, 1]
        assert lib.t3d_properties_get_vector(particle, b"colorStart", v, 3) == 1 and list(v)[:3] == [1, 0.5, 0.25]
        assert lib.t3d_properties_get_vector(particle, b"nothing", v, 3) == 0 and list(v)[:3] == [0, 0, 0]
        buf = C.create_string_buffer(1024)
        assert lib.t3d_properties_get_path(sprite, b"path", buf, 1024) == abi.OK and buf.value.decode() == tex
        assert lib.t3d_properties_get_path(sprite, b"width", buf, 1024) == abi.ERR_IO
        keys = [lib.t3d_properties_key(sprite, i) for i in range(lib.t3d_properties_count(sprite))]
        assert keys[:3] == [b"path", b"width", b"height"] and lib.t3d_properties_key(sprite, 99) is None
        rc, got, fc2, sw2, sh2, tex2 = parse_tree(particle)
        assert rc == abi.OK and (fc2, sw2, sh2, tex2) == (fc, sw, sh, tex) and got.as_dict() == want.as_dict()
        # the root (an unnamed namespace) is not a particle
        assert parse_tree(root)[0] == abi.ERR_IO
        # ... and the same tree built by hand, key by key
        mine = C.c_void_p()
        assert lib.t3d_properties_new(b"particle", b"copy", C.byref(mine)) == abi.OK
        try:
            child = C.c_void_p()
            assert lib.t3d_properties_add_namespace(mine, b"sprite", None, C.byref(child)) == abi.OK
            for src, dst in ((sprite, child), (particle, mine)):
                for i in range(lib.t3d_properties_count(src)):
                    k, val = lib.t3d_properties_key(src, i), lib.t3d_properties_value(src, i)
                    if k == b"path":
                        val = tex.encode()          # a hand-built tree has no directory to resolve against
                    assert lib.t3d_properties_set(dst, k, val) == abi.OK
            assert lib.t3d_properties_set(mine, b"emissionRate", b"1") == abi.OK and lib.t3d_properties_set(mine, b"emissionRate", b"55") == abi.OK
            rc, got, fc3, sw3, sh3, _ = parse_tree(mine)
            assert rc == abi.OK and (fc3, sw3, sh3) == (fc, sw, sh) and got.as_dict() == want.as_dict()
        finally:
            lib.t3d_properties_free(mine)
    finally:
        lib.t3d_properties_free(root)


def test_ellipsoid_acceptance_needs_no_square_root():
    # The spawn kernel tests `x*x + y*y + z*z > 1 + 2^-23` where the reference tests `sqrtf(...) > 1.0f`
    # (Vector3::length() in generateVectorInEllipsoid, PE.cpp:629-650): identical for every float.
    thr = np.uint32(0x3F800001).view(np.float32)
    near_one = np.arange(0x3F000000, 0x40000000, dtype=np.uint32).view(np.float32)       # every float in [0.5, 2)
    assert np.array_equal(np.sqrt(near_one) > np.float32(1.0), near_one > thr)
    sweep = np.arange(0, 0x7F800000, 61, dtype=np.uint32).view(np.float32)              # and a sweep of all positive floats
    assert np.array_equal(np.sqrt(sweep) > np.float32(1.0), sweep > thr)


def _measure(lib, g, font_size, spacing, text, size):
    w, h = C.c_uint32(), C.c_uint32()
    lib.t3d_host_font_measure_text(g.ctypes.data, g.size, font_size, spacing, text, len(text), size, C.byref(w), C.byref(h))
    return w.value, h.value


def test_font_measure_text_known_answers():
    # Font::measureText (Font.cpp:675-731): the widest line, `size` per line; '\n' alone ends a line, ' ' and '\t' advance by
    # the first glyph's advance (unscaled), a glyph by floor(advance * scale + spacing), everything else by nothing
    lib = abi.load()
    g = np.zeros(95, dtype=abi.GLYPH_DTYPE)
    g["advance"] = 10
    g["advance"][ord("W") - 32] = 17
    assert _measure(lib, g, 20, 0.0, b"", 0) == (0, 0)
    assert _measure(lib, g, 20, 0.0, b"abc", 0) == (30, 20)
    assert _measure(lib, g, 20, 0.0, b"a c\tW", 0) == (10 + 10 + 10 + 40 + 17, 20)
    assert _measure(lib, g, 20, 0.0, b"abc\nabcdW\n", 0) == (57, 60)               # a trailing '\n' counts a line
    assert _measure(lib, g, 20, 0.0, b"ab\rcd", 0) == (40, 20)                     # '\r' is skipped, not a line break
    assert _measure(lib, g, 20, 0.0, b"ab\x00cdefgh", 0) == (20, 20)               # c_str(): the text ends at its first NUL
    assert _measure(lib, g, 20, 0.0, b"a\xe9b", 0) == (20, 20)                     # bytes >= 128: signed char, no glyph
    assert _measure(lib, g, 20, 0.0, b"abc", 30) == (45, 30)                       # scale 1.5
    assert _measure(lib, g, 20, 0.1, b"abc", 30) == (3 * (15 + 3), 30)             # spacing = (int)(30 * 0.1) pixels per glyph
    assert _measure(lib, g, 20, 0.0, b"WW", 7) == (2 * 5, 7)                       # floor(17 * 0.35) = 5


@pytest.mark.skipif(not H.ref_available(), reason="needs oracle/_ref")
def test_font_measure_text_matches_the_reference():
    lib, R = abi.load(), H.ref(fresh=True)
    for glyph_count, font_size, size, spacing, n, seed in ((95, 24, 0, 0.0, 1, 3), (95, 24, 31, 0.125, 300, 4), (60, 18, 7, -0.05, 5000, 5),
                                                         (95, 32, 64, 0.3, 70000, 6), (10, 12, 12, 0.0, 999, 7)):
        g = H.random_font(seed, glyph_count, font_size)
        h = R.font_create(font_size, g.ctypes.data, glyph_count, 512, 512)
        R.font_set_character_spacing(h, spacing)
        for text in (H.random_text(n, seed), H.random_text(n, seed + 100).replace(b"\x00", b"\x01"), b"", b"\n\n", b"   \t"):
            w, hh = C.c_uint32(), C.c_uint32()
            R.font_measure_text(h, text, len(text), size, C.byref(w), C.byref(hh))
            assert _measure(lib, g, font_size, spacing, text, size) == (w.value, hh.value), (glyph_count, size, spacing, len(text))
        R.font_destroy(h)
