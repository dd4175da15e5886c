"""The host-side file parsers (OBJ / MTL, .config, .bvh cache) under AddressSanitizer + UndefinedBehaviorSanitizer: random
OBJ / MTL pairs, and byte-mutated, truncated and padded copies of valid .config and .bvh files. These files come from
users, so whatever they hold the parsers may refuse them but must not read or write out of bounds; every OBJ that loads
also goes through the SBVH + wide-BVH builder with whatever coordinates it held (NaN, inf, huge). No GPU needed:
tests/parser_harness.cpp links the host sources only."""
import json
import os
import random
import subprocess

import numpy as np
import pytest

from adypt_b200 import host, workloads as W
from test_config import GOOD
from test_host_builder import _fuzz_obj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mutations(data: bytes, rng, n):
    for trial in range(n):
        b = bytearray(data)
        kind = trial % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif kind == 1:
            b = b[: int(rng.integers(1, len(b)))]
        elif kind == 2:
            i = int(rng.integers(0, len(b)))
            b[i:i] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
        else:
            i = int(rng.integers(0, len(b)))
            j = min(len(b), i + int(rng.integers(1, 16)))
            b[i:j] = bytes([255] * (j - i))
        yield bytes(b)


def test_parsers_are_clean_under_asan_and_ubsan(tmp_path):
    csrc = os.path.join(ROOT, "adypt_b200", "csrc")
    exe = str(tmp_path / "harness")
    sources = [os.path.join(ROOT, "tests", "parser_harness.cpp"), os.path.join(csrc, "hostmath.cpp")] + [
        os.path.join(csrc, "host", f) for f in ("obj_loader.cpp", "config.cpp", "bvh_build.cpp", "host_api.cpp", "image_decode.cpp",
                                                "image_decode_more.cpp", "jpeg_decode.cpp")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fwrapv", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-DADYPT_NO_FMAD",
           "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + csrc, "-I/usr/local/cuda/include"] + sources + ["-lz", "-lpthread", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available here: " + r.stderr[-300:])
    corpus = tmp_path / "corpus"
    corpus.mkdir()
    rng = np.random.default_rng(99)
    n = 0
    # OBJ / MTL: random pairs in odd spellings (the generator of the parity fuzz test), each next to its own m.mtl
    rnd = random.Random(7)
    for trial in range(100):
        d = tmp_path / f"obj{trial}"
        d.mkdir()
        p = _fuzz_obj(rnd, str(d))
        data = open(p, "rb").read()
        os.replace(os.path.join(str(d), "m.mtl"), str(corpus / "m.mtl"))  # the last one stays: every t*.obj names m.mtl
        (corpus / f"{n:05d}.obj").write_bytes(data)
        n += 1
        for m in _mutations(data, rng, 1):
            (corpus / f"{n:05d}.obj").write_bytes(m)
            n += 1
    # .config: the valid file and mutations of its text
    good = json.dumps(GOOD, indent=4).encode()
    (corpus / f"{n:05d}.config").write_bytes(good)
    n += 1
    for m in _mutations(good, rng, 300):
        (corpus / f"{n:05d}.config").write_bytes(m)
        n += 1
    # .bvh: a real cache file written by the library, then mutations (header, counts, node bytes)
    hs = host.build_scene(W.tiny_scene("deep"))
    bvh = str(tmp_path / "t.bvh")
    hs.save_bvh(bvh)
    data = open(bvh, "rb").read()
    (corpus / f"{n:05d}.bvh").write_bytes(data)
    n += 1
    for m in _mutations(data, rng, 300):
        (corpus / f"{n:05d}.bvh").write_bytes(m)
        n += 1
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1:max_allocation_size_mb=4096")
    r = subprocess.run([exe, str(corpus)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    accepted = int(r.stdout.split("accepted")[1].split()[0])
    built = int(r.stdout.split("built")[1].split()[0])
    assert accepted >= 100 and built >= 80 and n > 800, (r.stdout, n)
