"""TEST INFRASTRUCTURE.  Rewrites a .cu file of da_sac_b200/csrc into C++ that g++ compiles against cuda_emul.h, WITHOUT
touching the product source:

    kernel<targs><<<grid, block, smem, stream>>>(args);   ->   cuda_emul::run_grid(grid, block, smem, [&]() { kernel<targs>(args); });

Everything else (kernel bodies, entry points, argument checks, launch geometry) is compiled as written; the CUDA keywords are
macros in cuda_emul.h.  Usage: python translate.py in.cu out.cpp
"""
import re
import sys


def matching(s, i, open_ch, close_ch):
    """index just past the bracket that closes s[i] (which must be open_ch)"""
    assert s[i] == open_ch
    depth = 0
    while True:
        c = s[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def split_top(s):
    """split at top-level commas"""
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


KEYWORDS = {"if", "for", "while", "switch", "catch", "return", "sizeof", "defined", "do", "else"}
PRIMITIVES = re.compile(r"\b(__syncthreads|__syncwarp|__shfl\w*)\b")


def syncing_functions(src):
    """names of the functions of this file that reach a barrier / warp primitive, directly or through a call (by name)"""
    bodies = {}
    for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\(", src):
        name = m.group(1)
        if name in KEYWORDS:
            continue
        try:
            e = matching(src, m.end() - 1, "(", ")")
        except IndexError:
            continue
        rest = re.match(r"\s*(?:const\s*)?\{", src[e:])
        if not rest:
            continue
        b0 = e + rest.end() - 1
        bodies.setdefault(name, "")
        bodies[name] += src[b0:matching(src, b0, "{", "}")]
    sync = {n for n, b in bodies.items() if PRIMITIVES.search(b)}
    changed = True
    while changed:
        changed = False
        for n, b in bodies.items():
            if n not in sync and any(re.search(r"\b%s\s*[<(]" % re.escape(c), b) for c in sync):
                sync.add(n); changed = True
    return sync, set(bodies)


STRIP_ALSO = {"mbar_wait", "smem_u32"}      # no asm inside, but it spins without yielding: cuda_emul_tc.h has a fiber-aware one


def strip_ptx_wrappers(src):
    """remove the SACB_DEVINL functions whose body is inline PTX (cuda_emul_tc.h defines functions of the same names)"""
    removed, out, pos = [], [], 0
    for m in re.finditer(r"(?:^template\s*<[^>\n]*>\s*\n)?^SACB_DEVINL[^\n(]*?\b([A-Za-z_]\w*)\s*\(", src, flags=re.M):
        if m.start() < pos:
            continue
        e = matching(src, m.end() - 1, "(", ")")
        rest = re.match(r"\s*\{", src[e:])
        if not rest:
            continue
        b1 = matching(src, e + rest.end() - 1, "{", "}")
        body = src[e:b1]
        if re.search(r"\basm\b", body) or m.group(1) in STRIP_ALSO:
            out.append(src[pos:m.start()])
            out.append("// [emulation] %s: see cuda_emul_tc.h" % m.group(1) + "\n" * src[m.start():b1].count("\n"))
            removed.append(m.group(1))
            pos = b1
    out.append(src[pos:])
    return "".join(out), removed


def translate_tc(src):
    """tensor-core translation units (sacb_common.cuh, sacb_gemm.cu): PTX wrappers out, dynamic shared memory from the emulator"""
    src, removed = strip_ptx_wrappers(src)
    src = re.sub(r"^#include\s+<(cuda_runtime|cuda_bf16|cuda)\.h>", r"// (\1.h: see cuda_emul.h / cuda_emul_tc.h)", src, flags=re.M)
    src = re.sub(r"^#define SACB_DEVINL .*$", "// (SACB_DEVINL: see cuda_emul.h)", src, flags=re.M)
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)cuda_emul::dyn_smem();", src)
    src = src.replace('#include "sacb_common.cuh"', '#include "sacb_common_emul.h"')
    return src, removed


def translate(src):
    sync, known = syncing_functions(src)
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        # kernel expression: identifier, optionally followed by <template args>, immediately before <<<
        j = i
        if src[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        m = re.search(r"[A-Za-z_][A-Za-z_0-9:]*$", src[:j])
        assert m, "no kernel name before <<< at offset %d" % i
        k0 = m.start()
        kernel = src[k0:i]
        e = src.index(">>>", i)
        cfg = split_top(src[i + 3:e])
        assert len(cfg) == 4, "launch configuration needs grid, block, smem, stream: " + src[i:e + 3]
        a0 = e + 3
        assert src[a0] == "(", src[a0:a0 + 20]
        a1 = matching(src, a0, "(", ")")
        assert src[a1] == ";"
        out.append(src[pos:k0])
        kname = re.match(r"[A-Za-z_]\w*", kernel.split("::")[-1]).group(0)
        assert kname in known, "launch of a kernel not defined in this file: " + kname
        out.append("cuda_emul::run_grid(\"%s\", %s, %s, %s, %s, [&]() { %s%s; });"
                   % (kname, cfg[0], cfg[1], cfg[2], "true" if kname in sync else "false", kernel, src[a0:a1]))
        out.append("\n" * (src[k0:a1 + 1].count("\n") - out[-1].count("\n")))      # keep the line numbers of the original
        pos = a1 + 1
    text = "".join(out)
    # cuda_emul.h (force-included) stands in for the CUDA headers and for sacb_common.cuh
    return re.sub(r'^#include\s+"sacb_common\.cuh"', "// (sacb_common.cuh: see cuda_emul.h)", text, flags=re.M)


if __name__ == "__main__":
    tc = "--tc" in sys.argv
    args = [a for a in sys.argv[1:] if a != "--tc"]
    with open(args[0]) as f:
        text = f.read()
    if tc:
        text, removed = translate_tc(text)
        sys.stderr.write("translate.py --tc %s: %d PTX wrappers replaced (%s)\n" % (args[0], len(removed), ", ".join(removed)))
    if "<<<" in text:
        text = translate(text)
    with open(args[1], "w") as f:
        f.write('#line 1 "%s"\n' % args[0])
        f.write(text)
