#!/bin/bash
# Is there a headless EGL / OpenGL 4.5 stack on the GPU box, i.e. could the reference's own GL compute path
# (glDispatchCompute of the unmodified shaders, src/Tracer/OglPathTracer.cpp:60,95-119) be timed as a third baseline?
# north_star: "the original GL compute path is timed as well only if a headless EGL context exists".
echo "== ldconfig EGL/GL/GLX/OpenGL libraries"; ldconfig -p | grep -i -E "libEGL|libGL\.|libGLX|libOpenGL|libGLESv2|libnvidia-egl|libnvidia-gl" || echo "none"
echo "== NVIDIA EGL vendor files"; ls -l /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>&1
echo "== driver GL libraries on disk"; find / -xdev \( -name "libEGL_nvidia*" -o -name "libnvidia-eglcore*" -o -name "libnvidia-glcore*" -o -name "libGLX_nvidia*" \) 2>/dev/null | head -20 || true
echo "== tools"; for t in eglinfo glxinfo glslangValidator; do printf "%s: " $t; command -v $t || echo "absent"; done
echo "== headers"; ls /usr/include/EGL /usr/include/GL /usr/include/GLFW 2>&1 | head
echo "== NVIDIA_DRIVER_CAPABILITIES=${NVIDIA_DRIVER_CAPABILITIES:-unset}"
echo "== /dev/dri"; ls -l /dev/dri 2>&1
python3 - <<'PY'
import ctypes, ctypes.util
for name in ("EGL", "GL", "OpenGL", "GLESv2"):
    p = ctypes.util.find_library(name)
    print(f"find_library({name}) = {p}")
p = ctypes.util.find_library("EGL")
if p:
    egl = ctypes.CDLL(p)
    egl.eglGetDisplay.restype = ctypes.c_void_p
    egl.eglGetDisplay.argtypes = [ctypes.c_void_p]
    d = egl.eglGetDisplay(None)
    major, minor = ctypes.c_int(0), ctypes.c_int(0)
    ok = egl.eglInitialize(ctypes.c_void_p(d), ctypes.byref(major), ctypes.byref(minor)) if d else 0
    print(f"eglGetDisplay(EGL_DEFAULT_DISPLAY) = {d}, eglInitialize = {ok}, version {major.value}.{minor.value}")
else:
    print("no libEGL: a headless EGL context cannot be created on this box")
PY
