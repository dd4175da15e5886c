"""The .config instance file: our reader/writer against the reference's InstanceConfig compiled
in place (oracle/_ref) -- same values, same accept/reject decisions (rapidjson's IsFloat quirk), same text."""
import json

import pytest

from adypt_b200 import AdyptError, host

GOOD = {
    "width": 1920, "height": 1080,
    "scene": {"filename": "models/city.obj"},
    "pathTracer": {"invocationSize": 8, "stackSize": 12, "maxBounce": 5, "subpixel": 8, "tmpLifetime": 16,
                   "rayTMin": 0.0001, "clamp": 4.0, "sun": [1.0, 0.9, 0.8]},
    "bvh": {"filename": "models/city.bvh", "maxSpatialDepth": 48, "triangleSAH": 0.3, "nodeSAH": 1.0},
    "camera": {"speed": 1.0, "mouseSensitive": 0.3, "fov": 45.0, "yaw": 7.0, "pitch": -31.0,
               "position": [91.87, 24.98, 189.76]},
}


def write(tmp_path, obj, name="a.config"):
    p = tmp_path / name
    p.write_text(obj if isinstance(obj, str) else json.dumps(obj))
    return str(p)


def fields(c):
    return (c.width, c.height, c.pt.invocation_size, c.pt.stack_size, c.pt.max_bounce, c.pt.subpixel, c.pt.tmp_lifetime,
            c.pt.ray_tmin, c.pt.clamp, tuple(c.pt.sun), c.bvh.max_spatial_depth, c.bvh.triangle_sah, c.bvh.node_sah,
            c.cam.speed, c.cam.mouse_sensitive, c.cam.fov, c.cam.yaw, c.cam.pitch, tuple(c.cam.position),
            c.obj_filename, c.bvh_filename)


def ref_fields(r):
    return (r.width, r.height, r.invocation_size, r.stack_size, r.max_bounce, r.subpixel, r.tmp_lifetime, r.ray_tmin, r.clamp,
            tuple(r.sun), r.max_spatial_depth, r.triangle_sah, r.node_sah, r.speed, r.mouse_sensitive, r.fov, r.yaw, r.pitch,
            tuple(r.position), r.obj_filename, r.bvh_filename)


def test_load_matches_reference(refmod, tmp_path):
    p = write(tmp_path, GOOD)
    assert fields(host.InstanceConfig.load(p)) == ref_fields(refmod.config_load(p))


def test_written_text_is_identical_to_the_reference(refmod, tmp_path):
    odd = json.loads(json.dumps(GOOD))
    odd["pathTracer"]["rayTMin"] = 1e-7
    odd["pathTracer"]["sun"] = [0.0, 123456789.0, 3.5e30]
    odd["camera"]["position"] = [-0.5, 1e21, 2.5e-5]
    odd["camera"]["yaw"] = 359.99
    odd["scene"]["filename"] = 'dir with "quotes"\\and\ttab/x.obj'
    for obj in (GOOD, odd):
        p = write(tmp_path, obj)
        ours = host.InstanceConfig.load(p).to_json()
        theirs = refmod.config_roundtrip_json(p)  # InstanceConfig::GetJson of the same file
        # same layout, key order, escapes and -- since the writer generates digits with rapidjson's own Grisu2 -- numbers
        assert ours == theirs
        # and what we write loads back, in both programs, to the same values
        q = str(tmp_path / "b.config")
        host.InstanceConfig.load(p).save(q)
        assert fields(host.InstanceConfig.load(q)) == ref_fields(refmod.config_load(q)) == fields(host.InstanceConfig.load(p))


@pytest.mark.parametrize("mutate", [
    lambda c: c["camera"].__setitem__("fov", 45),            # IsFloat quirk: an integer literal is NOT a Float
    lambda c: c.__setitem__("width", 1280.0),                 # ... and a double is not a Uint
    lambda c: c.__setitem__("width", -1),
    lambda c: c["pathTracer"].__setitem__("sun", [1.0, 1.0]),
    lambda c: c["pathTracer"].__setitem__("sun", [1, 1, 1]),
    lambda c: c["scene"].__setitem__("filename", 5),
    lambda c: c["pathTracer"].__setitem__("maxBounce", 4294967296),
])
def test_rejects_exactly_what_the_reference_rejects(refmod, tmp_path, mutate):
    bad = json.loads(json.dumps(GOOD))
    mutate(bad)
    p = write(tmp_path, bad)
    assert refmod.config_load(p) is None
    with pytest.raises(AdyptError) as e:
        host.InstanceConfig.load(p)
    assert "[PARSER]ERR" in str(e.value)


@pytest.mark.parametrize("mutate", [lambda c: c["bvh"].pop("nodeSAH"), lambda c: c.pop("camera")])
def test_missing_keys_are_rejected_where_the_reference_aborts(tmp_path, mutate):
    """Every key is mandatory. The reference indexes the missing member and dies in rapidjson's assertion
    (document.h operator[]); we return the error the CHECK macro would have printed."""
    bad = json.loads(json.dumps(GOOD))
    mutate(bad)
    with pytest.raises(AdyptError) as e:
        host.InstanceConfig.load(write(tmp_path, bad))
    assert "[PARSER]ERR: undefined" in str(e.value)


@pytest.mark.parametrize("text", ["", "[1,2]", '{"width": }', "{", "nonsense"])
def test_malformed_json_is_rejected(refmod, tmp_path, text):
    p = write(tmp_path, text)
    assert refmod.config_load(p) is None
    with pytest.raises(AdyptError):
        host.InstanceConfig.load(p)


def test_accepts_float_spellings_like_the_reference(refmod, tmp_path):
    txt = json.dumps(GOOD).replace('"fov": 45.0', '"fov": 4.5e1').replace('"clamp": 4.0', '"clamp": 4E0').replace('"yaw": 7.0', '"yaw": 700.0e-2')
    p = write(tmp_path, txt)
    assert fields(host.InstanceConfig.load(p)) == ref_fields(refmod.config_load(p))


def test_defaults_match_instanceconfig_hpp():
    c = host.InstanceConfig.default()
    assert (c.width, c.height) == (1280, 720)
    assert (c.pt.invocation_size, c.pt.stack_size, c.pt.max_bounce, c.pt.subpixel, c.pt.tmp_lifetime) == (8, 12, 5, 8, 16)
    assert abs(c.pt.ray_tmin - 1e-4) < 1e-10 and c.pt.clamp == 4.0
    assert (c.bvh.max_spatial_depth, c.bvh.node_sah) == (48, 1.0) and abs(c.bvh.triangle_sah - 0.3) < 1e-7
    assert c.cam.fov == 45.0 and c.cam.speed == 1.0
    with pytest.raises(AdyptError):
        host.InstanceConfig.load("/nonexistent/x.config")


@pytest.mark.parametrize("number", ["1e400", "1e309", "12345678901234567890123456789e290", "1" + "0" * 309, "0.1e310", "1e308", "0.001e311", "1e-400", "1e-99999999999"])
def test_numbers_too_big_for_a_double_reject_the_whole_file(refmod, tmp_path, number):
    """rapidjson's parser gives up on a number whose exponent or digit string overflows a double (reader.h:1232-1236,
    1313-1319) -- wherever it stands, even under a key nobody reads -- while huge NEGATIVE exponents quietly become 0."""
    txt = json.dumps(GOOD)[:-1] + ', "ignored": ' + number + "}"
    p = write(tmp_path, txt)
    r = refmod.config_load(p)
    if r is None:
        with pytest.raises(AdyptError):
            host.InstanceConfig.load(p)
    else:
        assert fields(host.InstanceConfig.load(p)) == ref_fields(r)


def test_config_text_fuzz_matches_reference(refmod, tmp_path):
    """300 random re-spellings of a valid instance file -- whitespace, key order, number formats (some making an int
    a float, which InstanceConfig refuses, or overflowing), escapes, duplicate and foreign keys, trailing junk, BOM,
    comments: both readers must agree on accept / reject and on every value."""
    import random
    rnd = random.Random(17)

    def fnum(v):
        if isinstance(v, int):
            return rnd.choice([str(v)] * 20 + [f"{v}.0", f"{v}e0", "-0" if v == 0 else str(v)])
        alts = [repr(float(v)), f"{v:.3f}", f"{v:e}", f"{v:E}", f"{v:.1f}", f"{int(v)}.5e-1", "-0.0", repr(float(v)), repr(float(v))]
        if rnd.random() < 0.03:
            alts.append(rnd.choice(["1e400", str(int(v)), "1e-400"]))
        return rnd.choice(alts)

    ws = lambda: rnd.choice(["", " ", "  ", "\n", "\t", "\r\n", " \n "])

    def emit(o):
        if isinstance(o, dict):
            items = list(o.items())
            if rnd.random() < 0.5:
                rnd.shuffle(items)
            parts = []
            for k, v in items:
                parts.append(ws() + json.dumps(k) + ws() + ":" + ws() + emit(v))
                if rnd.random() < 0.03:
                    parts.append(ws() + json.dumps(k) + ":" + emit(v))
                if rnd.random() < 0.08:
                    parts.append(ws() + '"extra%d":' % rnd.randrange(9) + rnd.choice(["1", "null", "[1,2]", '{"a":1}', '"x"', "true"]))
            return "{" + ",".join(parts) + ws() + "}"
        if isinstance(o, list):
            return "[" + ",".join(ws() + emit(v) + ws() for v in o) + "]"
        if isinstance(o, str):
            return rnd.choice([json.dumps(o), json.dumps(o).replace("/", "\\/"), json.dumps(o + "é"), json.dumps(o + "é", ensure_ascii=False), '"a\\u0041\\n"'])
        return fnum(o)

    accepted = 0
    for t in range(300):
        txt = emit(json.loads(json.dumps(GOOD)))
        if rnd.random() < 0.1:
            txt += rnd.choice([" ", "\n", "x", ",", "}"])
        if rnd.random() < 0.05:
            txt = "﻿" + txt
        if rnd.random() < 0.05:
            txt = "// c\n" + txt
        p = tmp_path / "f.config"
        p.write_text(txt, encoding="utf-8", newline="")
        r = refmod.config_load(str(p))
        try:
            o = host.InstanceConfig.load(str(p))
        except AdyptError:
            o = None
        assert (r is None) == (o is None), (t, txt)
        if r is not None:
            accepted += 1
            assert fields(o) == ref_fields(r), (t, txt)
    assert accepted >= 60


def test_number_text_equals_rapidjsons_writer(refmod):
    """adypt_config_format_double against rapidjson's internal::dtoa (what its Writer emits): Grisu2 digits -- not
    always the shortest -- and Prettify's fixed / exponent forms, on floats widened to double (what a .config holds),
    raw double bit patterns, subnormals, powers of two and ten."""
    import ctypes as C
    import numpy as np
    lib = host._lib()
    lib.adypt_config_format_double.argtypes = [C.c_double, C.c_char_p]
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.random(20000, dtype=np.float32).astype(np.float64),
        (rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 20000).astype(np.float32)).astype(np.float64),
        np.frombuffer(rng.integers(1, 0x7FEFFFFFFFFFFFFF, 20000, dtype=np.int64).tobytes(), dtype=np.float64),
        np.frombuffer(rng.integers(1, 0x000FFFFFFFFFFFFF, 5000, dtype=np.int64).tobytes(), dtype=np.float64),
        10.0 ** np.arange(-320, 308), 2.0 ** np.arange(-1074, 1023), np.arange(0, 300, dtype=np.float64),
        np.array([5e-324, 1.7976931348623157e308, 0.3, 0.1, 1 / 3, 1e21, 1e22, 9.999999999999999e20, 1e-6, 1e-7, 0.30000001192092896, -2.5, -0.0]),
    ])
    vals = vals[np.isfinite(vals)]
    theirs = refmod.dtoa(vals)
    buf = C.create_string_buffer(32)
    for v, t in zip(vals, theirs):
        assert lib.adypt_config_format_double(float(v), buf) == 0
        assert buf.value.decode() == t, (float(v), buf.value, t)
    assert lib.adypt_config_format_double(float("nan"), buf) != 0 and lib.adypt_config_format_double(float("inf"), buf) != 0
