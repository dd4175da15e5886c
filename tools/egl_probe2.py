#!/usr/bin/env python3
"""Second EGL probe (GPU box): the box has no libEGL.so (glvnd dispatcher) but the driver's vendor library
/usr/lib/libEGL_nvidia.so.0 is on disk. Can a headless OpenGL 4.5 context be created from it directly?"""
import ctypes as C, os, subprocess, sys
for path in ("/usr/lib/libEGL_nvidia.so.0", "/usr/lib/libGLX_nvidia.so.0", "/usr/lib/libnvidia-eglcore.so", "/usr/lib/libnvidia-glcore.so"):
    print("==", path, os.path.exists(path), os.path.realpath(path) if os.path.exists(path) else "")
    if os.path.exists(path):
        try:
            out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, timeout=30).stdout
            syms = [l.split()[-1] for l in out.splitlines() if l.strip()]
            print("   exported symbols:", len(syms), [s for s in syms if s.startswith(("egl", "__egl", "gl", "__glX"))][:40])
        except Exception as e:
            print("   nm failed:", e)
try:
    egl = C.CDLL("/usr/lib/libEGL_nvidia.so.0", mode=C.RTLD_GLOBAL)
    print("dlopen ok")
    for name in ("eglGetProcAddress", "eglGetDisplay", "eglInitialize", "eglQueryDevicesEXT", "eglGetPlatformDisplayEXT", "__egl_Main"):
        try:
            getattr(egl, name)
            print("  has", name)
        except AttributeError:
            print("  no ", name)
    try:
        gpa = egl.eglGetProcAddress
        gpa.restype = C.c_void_p
        gpa.argtypes = [C.c_char_p]
        for n in (b"eglQueryDevicesEXT", b"eglGetPlatformDisplayEXT", b"eglInitialize", b"glDispatchCompute"):
            print("  eglGetProcAddress(%s) = %s" % (n.decode(), gpa(n)))
        qd = C.CFUNCTYPE(C.c_uint, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int))(gpa(b"eglQueryDevicesEXT"))
        devs = (C.c_void_p * 16)()
        nd = C.c_int(0)
        print("  eglQueryDevicesEXT ->", qd(16, devs, C.byref(nd)), "devices:", nd.value)
        if nd.value > 0:
            gpd = C.CFUNCTYPE(C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p)(gpa(b"eglGetPlatformDisplayEXT"))
            dpy = gpd(0x313F, devs[0], None)  # EGL_PLATFORM_DEVICE_EXT
            print("  display:", dpy)
            init = egl.eglInitialize
            init.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
            ma, mi = C.c_int(0), C.c_int(0)
            print("  eglInitialize ->", init(dpy, C.byref(ma), C.byref(mi)), ma.value, mi.value)
    except Exception as e:
        print("  direct EGL use failed:", repr(e))
except OSError as e:
    print("dlopen failed:", e)
