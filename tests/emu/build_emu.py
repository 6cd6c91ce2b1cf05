"""TEST INFRASTRUCTURE ONLY: compiles the kernels a second time with g++ -DSYG_EMU (CUDA-on-CPU fibers, tests/emu/syg_emu.h)
so index arithmetic, barriers and host logic can be exercised without a GPU.  Never loaded by the sygnals_b200 package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sygnals_b200", "csrc")
# SYG_EMU_ASAN=1: AddressSanitizer build (run the emu tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so) and
# ASAN_OPTIONS=detect_leaks=0) -- the memory-safety check for the kernels' global / shared indexing on this GPU pool, where
# compute-sanitizer is closed.
ASAN = os.environ.get("SYG_EMU_ASAN", "") == "1"
OUT = os.path.join(HERE, "libsygb200_emu_asan.so" if ASAN else "libsygb200_emu.so")


def build_emu(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("syg_emu.h", "syg_emu.cpp")]
    deps.append(os.path.join(ROOT, "include", "sygb200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(p) <= os.path.getmtime(OUT) for p in deps):
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(ROOT, "build", "emu_obj_asan" if ASAN else "emu_obj")
    san = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g", "-O1"] if ASAN else []
    os.makedirs(objdir, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")] + [os.path.join(HERE, "syg_emu.cpp")]

    def cc(src):
        obj = os.path.join(objdir, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        cmd = ["g++", "-O2", "-std=c++17", "-DSYG_EMU", "-fPIC", "-pthread", "-I", HERE, "-I", CSRC, "-Wno-unused-value"] + san + [
               "-c", "-o", obj, "-x", "c++", src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ (emulator build) failed on {src}:\n" + r.stdout + r.stderr[-6000:])
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        objs = list(ex.map(cc, srcs))
    r = subprocess.run(["g++", "-shared", "-pthread"] + (["-fsanitize=address"] if ASAN else []) + ["-o", OUT + ".tmp"] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulator link) failed:\n" + r.stdout + r.stderr[-6000:])
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build_emu(force=True))
