#!/usr/bin/env python3
"""Writes tests/golden/texture_files.npz: image FILES in every format the reference's stb_image loads, with the pixels
stbi_load(file, ..., 3) of the reference's vendored stb_image (oracle/_ref) returns for them -- so that the decoders can be
checked against the reference's answers where /root/reference is absent. Run in the build container:
    python tests/golden/make_texture_golden.py
The files come from the generators of tests/test_textures.py (seeded), so the fixture is reproducible."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_textures as T  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    d = tempfile.mkdtemp()
    files = {}
    for name, data in T.rare_format_cases(d).items():
        if not name.startswith("CRASHES_STB") and len(data) <= 6000:
            files[name] = data
    for mk in (T.make_images, T.tga_cases, T.bmp_cases, T.jpeg_cases):
        for name, path in list(mk(d).items())[:6]:
            data = open(path, "rb").read()
            if len(data) <= 6000:
                files["common_" + name] = data
    out = {}
    names = []
    for i, (name, data) in enumerate(sorted(files.items())):
        p = os.path.join(d, "f.bin")
        open(p, "wb").write(data)
        img = ref.load_image_rgb8(p)
        names.append(name)
        out[f"file_{i}"] = np.frombuffer(data, dtype=np.uint8)
        out[f"shape_{i}"] = np.array(img.shape if img is not None else (0, 0, 0), dtype=np.int32)
        out[f"pixels_{i}"] = img.reshape(-1) if img is not None else np.zeros(0, dtype=np.uint8)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "texture_files.npz")
    np.savez_compressed(path, **out)
    print(path, len(names), "files", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
