"""Proves that additions to csrc/sacb_gemm.cu (new template instantiations behind switches that are off by default) leave the
GPU-verified kernels untouched: compiles the file as of a given commit and as of the working tree, then compares every default
instantiation instruction by instruction (kernel-parameter offsets c[0x0][..] normalised, since added parameters shift them).

    python profiles/sass_identity.py 289435f          # 289435f = last commit whose library passed pytest -m gpu on a B200
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRCS = ("da_sac_b200/csrc/sacb_gemm.cu", "da_sac_b200/csrc/sacb_p2p.cu")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c"]


def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    d, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append(re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[P]", m.group(1)))
    return d


def norm(name):
    """map an instantiation of the working tree to the name it had in the verified commit (template bools added since)"""
    n = re.sub(r"conv_gemm_pair_kernel(I(Lb0E)+EEv|E)14CUtensorMap_st(S\d_)+NS_8GemmArgsE", "conv_gemm_pair_kernel<default>", name)
    n = re.sub(r"conv_wgrad_pair_kernel(ILb0EEEv|E)14CUtensorMap_st(S\d_)+NS_9WgradArgsE", "conv_wgrad_pair_kernel<default>", n)
    n = re.sub(r"allreduce_sgd_kernel(ILb0EEEv|E)NS_7P2PArgsE", "allreduce_sgd_kernel<default>", n)
    return re.sub(r"(conv_(gemm|wgrad)_kernelILi\d+ELi\d+E)Lb0E", r"\1", n)


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        csrc = os.path.join(tmp, "da_sac_b200", "csrc"); inc = os.path.join(tmp, "include")
        os.makedirs(csrc); os.makedirs(inc)
        for path in SRCS + ("da_sac_b200/csrc/sacb_common.cuh", "include/sacb.h"):
            with open(os.path.join(tmp, path), "wb") as f:
                f.write(subprocess.run(["git", "-C", ROOT, "show", "%s:%s" % (commit, path)], capture_output=True, check=True).stdout)
        ok = True
        for src in SRCS:
            old_o, new_o = os.path.join(tmp, "old.o"), os.path.join(tmp, "new.o")
            subprocess.run(NVCC + [os.path.join(tmp, src), "-o", old_o], check=True)
            subprocess.run(NVCC + [os.path.join(ROOT, src), "-o", new_o], check=True)
            old = {norm(k): v for k, v in funcs(old_o).items()}
            for k, v in funcs(new_o).items():
                if "Lb1" in k:
                    print("new       %5d instr  %s" % (len(v), k[:90]))
                    continue
                same = old.get(norm(k)) == v
                ok &= same
                print("%-9s %5d instr  %s" % ("IDENTICAL" if same else "DIFFERENT", len(v), norm(k)[:90]))
        print("all default kernels identical to %s" % commit if ok else "MISMATCH against %s" % commit)
        return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
