#!/usr/bin/env python3
"""Build the UNMODIFIED reference (Tencent/ncnn, CPU path) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by
the product (ncnn_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and only as the checker / baseline.

What this does
--------------
The reference's sources are compiled *where they lie* under /root/reference by a
recipe that lives here (a generated Makefile + `make -j`); the reference's own CMake
build system is NOT run.  The four tiny headers CMake would have configured are
written by this script into oracle/_ref/<isa>/gen/:

  platform.h           <- src/platform.h.in with every `#cmakedefine01 X` resolved
  ncnn_export.h        <- `#define NCNN_EXPORT` (what generate_export_header emits)
  layer_declaration.h  <- one include + DEFINE_LAYER_CREATOR per layer class
  layer_registry.h     <- the creator tables, in src/CMakeLists.txt order
  layer_type_enum.h    <- the LayerType enum, same order

(mechanism: cmake/ncnn_add_layer.cmake:103-166, src/CMakeLists.txt:66-191).
Runtime ISA dispatch (NCNN_RUNTIME_CPU, which needs CMake to clone sources) is
replaced by two whole-library builds with fixed ISA flags:

  oracle/_ref/libncnn_ref_avx2.so     -mavx2 -mfma -mf16c
  oracle/_ref/libncnn_ref_avx512.so   + -mavx512f/cd/bw/dq/vl

oracle/ref.py picks the widest one /proc/cpuinfo supports, so the same tree runs on
the GPU box whatever its host CPU is.  Release flags follow src/CMakeLists.txt:374-383
(-Ofast -ffast-math ... ) and OpenMP is linked (/usr/bin/g++; the image's $CXX has no
libgomp.spec).  Outputs go ONLY under oracle/_ref/ (git-ignored, shipped by gpurun).
No reference source is copied into the repository.
"""
import os
import re
import subprocess
import sys

REF = os.environ.get("NCNN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ISA_FLAGS = {
    "avx2": "-mavx2 -mfma -mf16c",
    "avx512": "-mavx2 -mfma -mf16c -mavx512f -mavx512cd -mavx512bw -mavx512dq -mavx512vl",
}

# values for platform.h.in's #cmakedefine01 switches (everything else -> 0)
PLATFORM_ON = {
    "NCNN_STDIO", "NCNN_STRING", "NCNN_THREADS", "NCNN_C_API", "NCNN_PLATFORM_API", "NCNN_BATCH",
    "NCNN_PIXEL", "NCNN_PIXEL_ROTATE", "NCNN_PIXEL_AFFINE", "NCNN_PIXEL_DRAWING",
    "NCNN_GNU_INLINE_ASM", "NCNN_AVX", "NCNN_FMA", "NCNN_F16C", "NCNN_AVX2",
    "NCNN_INT8", "NCNN_WEIGHT_QUANT", "NCNN_BF16", "NCNN_FORCE_INLINE",
}

CORE_SRCS = [
    "allocator.cpp", "benchmark.cpp", "blob.cpp", "c_api.cpp", "command.cpp", "cpu.cpp", "datareader.cpp",
    "expression.cpp", "gpu.cpp", "layer.cpp", "mat.cpp", "mat_pixel.cpp", "mat_pixel_affine.cpp",
    "mat_pixel_drawing.cpp", "mat_pixel_resize.cpp", "mat_pixel_rotate.cpp", "modelbin.cpp", "net.cpp",
    "option.cpp", "paramdict.cpp", "pipeline.cpp", "pipelinecache.cpp", "simpleocv.cpp", "simpleomp.cpp",
    "simplestl.cpp", "simplemath.cpp", "simplevk.cpp",
]


def layer_list():
    """(ClassName, enabled) in registry order, from src/CMakeLists.txt."""
    txt = open(os.path.join(REF, "src/CMakeLists.txt")).read()
    out = []
    for m in re.finditer(r"^ncnn_add_layer\((\w+)(?:\s+(\w+))?\)", txt, re.M):
        out.append((m.group(1), (m.group(2) or "ON").upper() != "OFF"))
    return out


def gen_headers(gen, isa):
    os.makedirs(gen, exist_ok=True)
    on = set(PLATFORM_ON)
    if isa == "avx512":
        on.add("NCNN_AVX512")
    src = open(os.path.join(REF, "src/platform.h.in")).read()

    def sub01(m):
        return "#define %s %d" % (m.group(1), 1 if m.group(1) in on else 0)

    src = re.sub(r"#cmakedefine01 (\w+)", sub01, src)
    src = src.replace('#cmakedefine NCNN_VERSION_STRING "@NCNN_VERSION_STRING@"', '#define NCNN_VERSION_STRING "1.0.oracle"')
    src = src.replace("#cmakedefine NCNN_VERSION_NUMBER @NCNN_VERSION_NUMBER@", "#define NCNN_VERSION_NUMBER 20260101")
    # NCNN_SIMPLEOCV stays off for the library proper; the two translation units that need the reference's OpenCV stand-in
    # (simpleocv.cpp and the YOLOv8 example driver) are compiled with -DNCNN_SIMPLEOCV=1
    src = src.replace("#define NCNN_SIMPLEOCV 0", "#ifndef NCNN_SIMPLEOCV\n#define NCNN_SIMPLEOCV 0\n#endif")
    old = open(os.path.join(gen, "platform.h")).read() if os.path.exists(os.path.join(gen, "platform.h")) else None
    if old != src:
        open(os.path.join(gen, "platform.h"), "w").write(src)
    open(os.path.join(gen, "ncnn_export.h"), "w").write(
        "#ifndef NCNN_EXPORT_H\n#define NCNN_EXPORT_H\n#define NCNN_EXPORT __attribute__((visibility(\"default\")))\n"
        "#define NCNN_NO_EXPORT\n#define NCNN_DEPRECATED\n#endif\n")

    decl, reg, reg_arch, enum = [], [], [], []
    for idx, (cls, enabled) in enumerate(layer_list()):
        name = cls.lower()
        has_arch = enabled and os.path.exists(os.path.join(REF, "src/layer/x86/%s_x86.cpp" % name))
        if enabled:
            decl.append('#include "layer/%s.h"\nnamespace ncnn { DEFINE_LAYER_CREATOR(%s) }\n' % (name, cls))
            reg.append('{"%s", %s_layer_creator},\n' % (cls, cls))
        else:
            reg.append('{"%s", 0},\n' % cls)
        if has_arch:
            decl.append('#include "layer/x86/%s_x86.h"\nnamespace ncnn { DEFINE_LAYER_CREATOR(%s_x86) }\n' % (name, cls))
            reg_arch.append('{"%s", %s_x86_layer_creator},\n' % (cls, cls))
        else:
            reg_arch.append('{"%s", 0},\n' % cls)
        enum.append("%s = %d,\n" % (cls, idx))
    open(os.path.join(gen, "layer_declaration.h"), "w").write("".join(decl))
    open(os.path.join(gen, "layer_registry.h"), "w").write(
        "static const layer_registry_entry layer_registry[] = {\n%s};\n"
        "static const layer_registry_entry layer_registry_arch[] = {\n%s};\n" % ("".join(reg), "".join(reg_arch)))
    open(os.path.join(gen, "layer_type_enum.h"), "w").write("".join(enum))
    for f in ("layer_shader_registry.h", "layer_shader_spv_data.h", "layer_shader_type_enum.h"):
        open(os.path.join(gen, f), "w").write("\n")


def sources():
    srcs = [os.path.join(REF, "src", s) for s in CORE_SRCS]
    for cls, enabled in layer_list():
        if not enabled:
            continue
        name = cls.lower()
        srcs.append(os.path.join(REF, "src/layer/%s.cpp" % name))
        for suffix in ("_x86.cpp",):  # the _x86_<isa>.cpp files only serve NCNN_RUNTIME_CPU dispatch
            p = os.path.join(REF, "src/layer/x86/%s%s" % (name, suffix))
            if os.path.exists(p):
                srcs.append(p)
    return srcs


def build(isa, jobs):
    bdir = os.path.join(OUT, isa)
    gen = os.path.join(bdir, "gen")
    obj = os.path.join(bdir, "obj")
    os.makedirs(obj, exist_ok=True)
    gen_headers(gen, isa)
    cflags = ("-std=c++11 -Ofast -ffast-math -DNDEBUG -fPIC -fopenmp -fvisibility=hidden -fvisibility-inlines-hidden "
              "-w %s -I%s -I%s/src -I%s/src/layer -I%s/src/layer/x86" % (ISA_FLAGS[isa], gen, REF, REF, REF))
    rules, objs = [], []
    for s in sources():
        o = os.path.join(obj, os.path.relpath(s, os.path.join(REF, "src")).replace("/", "__")[:-4] + ".o")
        extra = ""
        if os.path.basename(s) == "simpleocv.cpp":
            o = o[:-2] + "_on.o"  # (a new object name: the cached one was compiled with NCNN_SIMPLEOCV 0, i.e. empty)
            extra = " -DNCNN_SIMPLEOCV=1"
        objs.append(o)
        rules.append("%s: %s\n\t@echo CXX %s\n\t@%s %s%s -c $< -o $@\n" % (o, s, os.path.basename(s), CXX, cflags, extra))
    # the reference's YOLOv8 post-processing (file-static functions of examples/yolov8.cpp), made callable by a driver TU of
    # OURS that includes the example where it lies (oracle/yolov8_example_driver.cpp)
    ydrv = os.path.join(HERE, "yolov8_example_driver.cpp")
    ydrv_o = os.path.join(obj, "yolov8_example_driver.o")
    objs.append(ydrv_o)
    rules.append("%s: %s %s\n\t@echo CXX yolov8_example_driver.cpp\n\t@%s %s -DUSE_NCNN_SIMPLEOCV -DNCNN_SIMPLEOCV=1 -DNCNN_REFERENCE_YOLOV8_CPP='\"%s\"' -c $< -o $@\n"
                 % (ydrv_o, ydrv, os.path.join(REF, "examples/yolov8.cpp"), CXX, cflags, os.path.join(REF, "examples/yolov8.cpp")))
    # the driver is OUR code (oracle/ref_driver.cpp): extra C entry points next to the reference's own c_api
    drv = os.path.join(HERE, "ref_driver.cpp")
    drv_o = os.path.join(obj, "ref_driver.o")
    objs.append(drv_o)
    rules.append("%s: %s\n\t@echo CXX ref_driver.cpp\n\t@%s %s -fvisibility=default -c $< -o $@\n" % (drv_o, drv, CXX, cflags))
    lib = os.path.join(OUT, "libncnn_ref_%s.so" % isa)
    mk = os.path.join(bdir, "Makefile")
    with open(mk, "w") as f:
        f.write("all: %s\n%s: %s\n\t@echo LINK $@\n\t@%s -shared -fopenmp -o $@ $^ -lpthread\n\n" % (lib, lib, " ".join(objs), CXX))
        f.write("\n".join(rules))
    subprocess.check_call(["make", "-f", mk, "-j%d" % jobs], cwd=bdir)
    return lib


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print("reference tree %s not present; keeping prebuilt oracle/_ref as is" % REF)
        return 0
    isas = sys.argv[1:] or ["avx2", "avx512"]
    jobs = int(os.environ.get("JOBS", os.cpu_count() or 4))
    for isa in isas:
        print("== building reference for", isa)
        print(build(isa, jobs))
    return 0


if __name__ == "__main__":
    sys.exit(main())
