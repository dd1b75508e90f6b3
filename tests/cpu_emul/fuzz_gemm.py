"""TEST INFRASTRUCTURE.  Random convolution geometries through the tcgen05 / TMA kernels of csrc/sacb_gemm.cu running on the
primitive model (cuda_emul_tc.h), compared with the formula model of include/sacb.h -- or, with --variant, a kernel variant
selected by an environment switch compared BIT FOR BIT with the default kernels.

    python tests/cpu_emul/fuzz_gemm.py --seed 3 --cases 150
    python tests/cpu_emul/fuzz_gemm.py --seed 3 --cases 150 --variant SACB_EPI_STAGED     (or SACB_TAIL_SPLIT)
    SACB_EMUL_ASYNC=1 SACB_EMUL_SCHED_SEED=5 python tests/cpu_emul/fuzz_gemm.py ...        (late TMA / MMA, random warp order)
    SACB_EMUL_SMS=148 ...                                                                  (default 8: partial waves with small problems)
Round 1: 800 cases vs the formula model and 600 variant cases, no mismatch."""
import argparse
import ctypes as C
import os
import random
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--variant", default=None, help="environment switch of a kernel variant (compared bit for bit with the default)")
    a = ap.parse_args(argv)
    os.environ.setdefault("SACB_EMUL_SMS", "8")
    import torch
    import emul_harness as E
    import test_emul_tc_cpu as T
    E.emul_lib()

    def load(name, tag, **env):
        dst = os.path.join(tempfile.gettempdir(), "fz_%s_%d.so" % (tag, os.getpid()))
        shutil.copy(os.path.join(T.BUILD, name), dst)
        lib = C.CDLL(dst)
        lib.sacb_last_error.restype = C.c_char_p
        lib.sacb_emul_last_kernel.restype = C.c_char_p
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        T.run(lib, (1, 9, 9, 64, 64, 1, 1, 1, 0))              # the first call reads the switches
        for k, v in old.items():
            os.environ.pop(k) if v is None else os.environ.__setitem__(k, v)
        return lib

    ref = load("libsacb_emul_tc.so", "base") if a.variant else load("libsacb_emul.so", "model")
    tc = load("libsacb_emul_tc.so", "var", **{a.variant: "1"}) if a.variant else load("libsacb_emul_tc.so", "tc")
    rnd = random.Random(a.seed)
    bad = ran = hit = 0
    t0 = time.time()
    for it in range(a.cases):
        R = rnd.choice([1, 1, 3, 3, 3, 7])
        s = rnd.choice([1, 1, 1, 2])
        d = rnd.choice([1, 1, 2, 4, 6, 12]) if R == 3 else 1
        pad = rnd.choice([0, d * (R // 2), d * (R // 2), 1]) if R > 1 else 0
        lo = max(1, (R - 1) * d + 1 - 2 * pad)
        N = rnd.choice([1, 1, 2, 3]); H = rnd.randint(lo, max(lo, 24)); W = rnd.randint(lo, max(lo, 24))
        Cc = 64 * rnd.choice([1, 1, 2, 3, 4]); K = rnd.choice([32, 64, 64, 128, 192, 256, 256, 512, 768])
        if a.variant:                                           # the variants live in the CTA-pair kernel: K % 256 == 0, mostly 1x1 + residual
            K = rnd.choice([256, 256, 512, 768, 1024])
            if rnd.random() < 0.6:
                R, s, d, pad = 1, 1, 1, 0
            H = rnd.randint(lo, 40); W = rnd.randint(lo, 40)
        geom = (N, H, W, Cc, K, R, s, d, pad)
        P, Q = T.L.conv_out_hw(H, W, R, s, d, pad)
        if P < 1 or Q < 1:
            continue
        kind = "fprop" if a.variant else rnd.choice(["fprop", "fprop", "wgrad"])
        ran += 1
        try:
            if kind == "fprop":
                epi = rnd.choice(["res", "res", "dgrad", "plain"] if a.variant else ["plain", "res", "dgrad", "head"])
                kv = rnd.randint(1, K) if epi == "head" else None
                prec = rnd.choice([0, 0, 1])
                x = T.run(ref, geom, seed=it, epi=epi, k_valid=kv, precision=prec)
                y = T.run(tc, geom, seed=it, epi=epi, k_valid=kv, precision=prec)
                if a.variant:
                    ok = torch.equal(x["f32"], y["f32"]) and torch.equal(x["hi"], y["hi"]) and torch.equal(x["lo"], y["lo"]) and \
                        (y["colsum"] is None or T.close(y["colsum"], x["colsum"], 1e-5))
                    f = T.template_flags(tc)
                    hit += bool(f.get("STAGED") or f.get("TSPLIT"))
                else:
                    ok = T.close(y["f32"], x["f32"], 3e-5) and (y["nchw"] is None or T.close(y["nchw"], x["nchw"], 3e-5)) and \
                        (y["hi"] is None or T.close(y["hi"].float() + y["lo"].float(), x["hi"].float() + x["lo"].float(), 3e-5)) and \
                        (y["colsum"] is None or T.close(y["colsum"], x["colsum"], 2e-4))
                what = (kind, geom, epi, kv, prec)
            else:
                if K % 64:
                    continue
                kv = rnd.choice([None, None, rnd.randint(1, K)]); splits = rnd.choice([0, 0, 1, 2, 5]); prec = rnd.choice([0, 0, 1])
                x, _ = T.wgrad(ref, geom, seed=it, k_valid=kv, precision=prec)
                y, _ = T.wgrad(tc, geom, seed=it, k_valid=kv, splits=splits, precision=prec)
                ok = T.close(y, x, 3e-5) and not torch.isnan(y).any()
                what = (kind, geom, kv, splits, prec)
        except AssertionError as e:
            ok, what = False, (kind, geom, "assert", str(e)[:200])
        if not ok:
            bad += 1
            print("MISMATCH", what, tc.sacb_emul_last_kernel().decode(), flush=True)
    print("fuzz seed %d%s: %d cases, %d mismatches%s, %.0f s" % (a.seed, " " + a.variant if a.variant else "", ran, bad,
                                                               ", %d ran the variant's instantiation" % hit if a.variant else "", time.time() - t0))
    return bad, ran, hit


if __name__ == "__main__":
    sys.exit(1 if main()[0] else 0)
