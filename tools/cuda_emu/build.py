"""Build the host emulation of the library: every source under rtlsdr_wsprd_b200/csrc is transpiled (transpile.py) into <outdir>,
compiled by g++ against the stand-in CUDA headers of this directory and linked with the fiber runtime into
<outdir>/libwsprd_b200_emu.so -- the same C ABI as the CUDA build (select it with WSPR_B200_LIB).
-ffp-contract=off plays the part of nvcc's -fmad=false: no multiply-add is fused that the sources do not fuse themselves.
Usage: python tools/cuda_emu/build.py <outdir>"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from transpile import transpile  # noqa: E402

SOURCES = ["wspr_kernels", "wspr_decode", "wspr_frontend", "wspr_abi"]


def build(outdir, opt="-O2", defs=(), csrc=None):
    """defs: extra -D options (the build-time variants of the kernels); csrc: another copy of the csrc directory (a patched one)"""
    csrc = csrc or os.path.join(ROOT, "rtlsdr_wsprd_b200", "csrc")
    dst = os.path.join(outdir, "rtlsdr_wsprd_b200", "csrc")
    os.makedirs(dst, exist_ok=True)
    os.makedirs(os.path.join(outdir, "include"), exist_ok=True)
    with open(os.path.join(ROOT, "include", "wspr_b200.h")) as f:
        text = f.read()
    with open(os.path.join(outdir, "include", "wspr_b200.h"), "w") as f:
        f.write(text)
    for name in sorted(os.listdir(csrc)):
        path = os.path.join(csrc, name)
        if name.endswith((".cuh", ".h")):
            with open(path) as f:
                text = transpile(f.read()) if name.endswith(".cuh") else f.read()
            with open(os.path.join(dst, name), "w") as f:
                f.write(text)
        elif name.endswith(".cu"):
            with open(path) as f:
                text = transpile(f.read())
            with open(os.path.join(dst, name[:-3] + ".cpp"), "w") as f:
                f.write(text)
    flags = ["g++", opt, "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fvisibility=default", "-w", "-I" + HERE, "-I" + dst] + list(defs)
    objs, procs = [], []
    for name in SOURCES:
        obj = os.path.join(dst, name + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(flags + ["-c", os.path.join(dst, name + ".cpp"), "-o", obj], stderr=subprocess.PIPE, text=True))
    obj = os.path.join(dst, "emu_runtime.o")
    objs.append(obj)
    procs.append(subprocess.Popen(flags + ["-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", obj], stderr=subprocess.PIPE, text=True))
    errors = ""
    for p in procs:
        _, err = p.communicate()
        if p.returncode:
            errors += err
    if errors:
        raise RuntimeError("host compilation of the transpiled sources failed:\n" + errors[-6000:])
    lib = os.path.join(outdir, "libwsprd_b200_emu.so")
    subprocess.run(["g++", "-shared", "-o", lib] + objs + ["-lpthread"], check=True)
    return lib


if __name__ == "__main__":
    # python tools/cuda_emu/build.py <outdir> [-DNAME=VALUE ...] [--csrc DIR]
    args = sys.argv[1:]
    other = None
    if "--csrc" in args:
        k = args.index("--csrc")
        other = args[k + 1]
        del args[k:k + 2]
    print(build(args[0] if args and not args[0].startswith("-D") else "/tmp/wspr_b200_emu", defs=[a for a in args if a.startswith("-D")], csrc=other))
