"""Builds tests/_build/libemdee_cusim.so: the product's abi.cpp + engine.cu (+ engine_*.cuh), unchanged except for
the two syntactic rewrites below, compiled by g++ against the execution-model emulator tests/cusim/cusim.h.
TEST INFRASTRUCTURE (see cusim.h): lets the kernel logic run on a machine without a GPU.

  kernel<<<grid, block, smem, stream>>>(args)   ->  cusim::launch([&] { kernel(args); }, grid, block, smem, stream)
  extern __shared__ <attrs> T name[];           ->  T* name = reinterpret_cast<T*>(cusim::dyn_smem());
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "emdee_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build")
GEN = os.path.join(OUT, "cusim_src")
# CUSIM_FMA=1: a second build that lets g++ contract a*b+c into fused multiply-adds (as nvcc does by default), to check on
# the CPU that the tolerances of the GPU tests survive the different rounding of the device build
FMA = os.environ.get("CUSIM_FMA") == "1"
LIB = os.path.join(OUT, "libemdee_cusim_fma.so" if FMA else "libemdee_cusim.so")
PARTS = ["engine.cu", "engine_common.cuh", "engine_list.cuh", "engine_force.cuh",
         "engine_dynamics.cuh", "engine_bodies.cuh", "engine_bonded.cuh", "engine_ewald.cuh", "engine_dist.cuh", "engine_extra.cuh"]


def _match_back(text, i, open_ch, close_ch):
    """text[i] == close_ch: index of the matching open_ch."""
    depth = 0
    while i >= 0:
        c = text[i]
        if c == close_ch:
            depth += 1
        elif c == open_ch:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced")


def _match_fwd(text, i, open_ch, close_ch):
    depth = 0
    while i < len(text):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            return "".join(out)
        j = k - 1
        while text[j].isspace():
            j -= 1
        if text[j] == ">":                      # template arguments of the kernel
            j = _match_back(text, j, "<", ">") - 1
            while text[j].isspace():
                j -= 1
        end_name = j + 1
        while j >= 0 and (text[j].isalnum() or text[j] in "_:"):
            j -= 1
        start = j + 1
        callee = text[start:k].strip()
        assert re.match(r"[A-Za-z_]", text[start]) and end_name > start, text[max(0, k - 80):k + 20]
        e = text.find(">>>", k)
        cfg = text[k + 3:e].strip()
        p = e + 3
        while text[p].isspace():
            p += 1
        assert text[p] == "(", text[k:k + 120]
        q = _match_fwd(text, p, "(", ")")
        args = text[p + 1:q]
        out.append(text[pos:start])
        out.append(f"cusim::launch([&] {{ {callee}({args}); }}, {cfg})")
        pos = q + 1


def rewrite_dyn_smem(text):
    pat = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:<> ]+?)\s+(\w+)\[\];")
    return pat.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(cusim::dyn_smem());", text)


def generate():
    os.makedirs(GEN, exist_ok=True)
    for name in PARTS:
        text = open(os.path.join(CSRC, name)).read()
        text = rewrite_dyn_smem(rewrite_launches(text))
        path = os.path.join(GEN, name if name.endswith(".cuh") else name.replace(".cu", ".cpp"))
        if not os.path.exists(path) or open(path).read() != text:
            open(path, "w").write(text)


def build(force=False):
    deps = [os.path.join(CSRC, n) for n in PARTS + ["abi.cpp", "engine.h", "nb_math.h"]] + \
           [os.path.join(HERE, "cusim.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    generate()
    fp = ["-O2", "-march=x86-64-v3", "-ffp-contract=fast"] if FMA else ["-O1", "-g", "-ffp-contract=off"]
    cmd = ["/usr/bin/g++", "-std=c++17", *fp, "-fPIC", "-shared", "-Wl,-Bsymbolic",
           "-Wl,--no-undefined", "-Wno-unused-function", "-Wno-unused-variable",
           "-I" + os.path.join(HERE, "include"), "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
           "-x", "c++", os.path.join(CSRC, "abi.cpp"), os.path.join(GEN, "engine.cpp"), "-o", LIB + f".tmp{os.getpid()}", "-ldl"]
    subprocess.check_call(cmd)
    os.replace(LIB + f".tmp{os.getpid()}", LIB)   # atomic: concurrent builders (several ranks) never expose a half-written file
    return LIB


FAKE_NCCL = os.path.join(OUT, "libnccl_fake.so")


def build_fake_nccl(force=False):
    src = os.path.join(HERE, "fake_nccl.cpp")
    if not force and os.path.exists(FAKE_NCCL) and os.path.getmtime(src) <= os.path.getmtime(FAKE_NCCL):
        return FAKE_NCCL
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + HERE, src, "-o", FAKE_NCCL, "-lrt", "-pthread"])
    return FAKE_NCCL


if __name__ == "__main__":
    print(build(force=True))
    print(build_fake_nccl(force=True))
