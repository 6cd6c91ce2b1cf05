"""TEST INFRASTRUCTURE ONLY: compiles the kernels a second time with g++ -DSYG_EMU (CUDA-on-CPU fibers, tests/emu/syg_emu.h)
so index arithmetic, barriers and host logic can be exercised without a GPU.  Never loaded by the sygnals_b200 package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sygnals_b200", "csrc")
OUT = os.path.join(HERE, "libsygb200_emu.so")


def build_emu(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("syg_emu.h", "syg_emu.cpp")]
    deps.append(os.path.join(ROOT, "include", "sygb200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(p) <= os.path.getmtime(OUT) for p in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-DSYG_EMU", "-fPIC", "-shared", "-pthread", "-I", HERE, "-I", CSRC,
           "-Wno-unused-value", "-o", OUT + ".tmp", "-x", "c++", os.path.join(CSRC, "syg_api.cu"), os.path.join(HERE, "syg_emu.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulator build) failed:\n" + r.stdout + r.stderr[-6000:])
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build_emu(force=True))
