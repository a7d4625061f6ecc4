"""TEST INFRASTRUCTURE: builds a CPU-emulated copy of selected product CUDA sources (see cuda_emu.h).
Two textual transformations are applied to each .cu file before it goes through g++:
  kernel<<<grid, block, smem, stream>>>(args);   ->  emu_launch(dim3(grid), dim3(block), smem, [=]() { kernel(args); });
  extern __shared__ [__align__(n)] T name[];      ->  T *name = (T*)emu_dyn_smem();
Everything else is the product source, compiled unchanged with tests/cuda_emu/cuda_emu.h force-included."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "minimap2_b200", "csrc")
BUILD = os.path.join(HERE, "_build")


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(src):
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1 *\2 = (\1*)emu_dyn_smem();", src)
    out, i = "", 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out += src[i:]; break
        k = j  # walk back over the kernel expression: identifier, optional template argument list
        depth = 0
        while k > i:
            c = src[k - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            k -= 1
        name = src[k:j]
        e = src.find(">>>", j)
        cfg = split_top(src[j + 3:e])
        assert src[e + 3] == "(", src[e:e + 40]
        depth, p = 0, e + 3
        while True:
            if src[p] == "(":
                depth += 1
            elif src[p] == ")":
                depth -= 1
                if depth == 0:
                    break
            p += 1
        args = src[e + 4:p]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[i:k] + "(emu_trace(\"%s\"), emu_launch(dim3(%s), dim3(%s), (size_t)(%s), [=]() { %s(%s); }))" % (name.replace('"', ''), cfg[0], cfg[1], smem, name, args)
        i = p + 1
    return out


def build(name, files, extra=("emu_stubs.cc",)):
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "lib%s.so" % name)
    srcs = [os.path.join(CSRC, f) for f in files] + [os.path.join(HERE, "emu_runtime.cc"), os.path.join(HERE, "cuda_emu.h"), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs):
        return lib
    gen = []
    for f in files:
        o = os.path.join(BUILD, f.rsplit(".", 1)[0] + ".emu.cc")
        open(o, "w").write("#line 1 \"%s\"\n" % os.path.join(CSRC, f) + transform(open(os.path.join(CSRC, f)).read()))
        gen.append(o)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-w", "-include", os.path.join(HERE, "cuda_emu.h"),
           "-I" + os.path.join(HERE, "fake_include"), "-I" + HERE, "-I" + CSRC, "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"] + gen + [os.path.join(HERE, "emu_runtime.cc")] + [os.path.join(HERE, x) for x in extra] + ["-Wl,-Bsymbolic", "-lz", "-o", lib]  # -Bsymbolic: the emulated cuda* entry points win even if a real libcudart is already loaded (torch)
    subprocess.check_call(cmd)
    return lib


ALL = ["mmb_ctx.cu", "ksw_fast.cu", "ksw_extd2.cu", "sketch.cu", "scan.cu", "seed.cu", "chain.cu", "index.cu", "map.cu", "synth.cu", "finalize.cu",
       "align.cc", "hits.cc", "format.cc", "options.cc", "fastx.cc"]

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "k3"
    if which == "k3":
        print(build("mmb_emu_k3", ["mmb_ctx.cu", "ksw_fast.cu", "ksw_extd2.cu"]))
    else:
        print(build("mmb_emu_all", ALL, extra=()))
