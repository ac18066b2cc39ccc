"""TEST INFRASTRUCTURE ONLY: compile quaff_b200/csrc/quaffgpu.cu with g++ against the CUDA-on-threads
shim (cuda_emu.h) into tests/emu/libquaffgpu_emu.so, so the kernels' logic can be checked against the
oracle on a GPU-less machine.  Never loaded by the quaff_b200 package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "quaff_b200", "csrc")
LIB = os.path.join(HERE, "libquaffgpu_emu.so")


def build(force=False):
    if os.environ.get("QG_EMU_LIB"):                      # e.g. an AddressSanitizer build of the same sources
        return os.environ["QG_EMU_LIB"]
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"),
                                                                 os.path.join(ROOT, "include", "quaffgpu.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DQG_EMU", "-ffp-contract=off", "-x", "c++",
           "-I", HERE, "-I", CSRC, os.path.join(CSRC, "quaffgpu.cu"), os.path.join(HERE, "cuda_emu.cpp"),
           "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL", "-o", LIB, "-lpthread", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emu build failed:\n" + (res.stdout + res.stderr)[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
