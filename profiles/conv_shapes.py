"""Per-shape timing of the conv GEMM kernels at configs[1] sizes (24 crops of 512x512, ResNet-101).

    python profiles/conv_shapes.py                 # CUDA-event table of every distinct conv shape (fprop/dgrad/wgrad)
    python profiles/conv_shapes.py one <name> <mode>   # a few launches of one shape, for `ncu --set full -k regex:conv_`
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da_sac_b200 import engine as E, lib as L  # noqa: E402

N = int(os.environ.get("SACB_CROPS", "24"))
net = E.build_resnet101(512, 512)
dev = torch.device("cuda")
bf = torch.bfloat16


def rnd(n, dtype=bf):
    return (torch.randn(n, device=dev) * 0.5).to(dtype)


def make(s):
    d = dict(s=s)
    d["x"] = (rnd(N * s.hin * s.win * s.C), rnd(N * s.hin * s.win * s.C))
    M = N * s.hout * s.wout
    d["M"] = M
    d["wf"] = (rnd(s.R * s.R * s.Kf * s.C), rnd(s.R * s.R * s.Kf * s.C))
    d["wt"] = (rnd(s.R * s.R * s.C * s.Kt), rnd(s.R * s.R * s.C * s.Kt))
    d["g"] = (rnd(M * s.Kt), rnd(M * s.Kt))
    d["out"] = (torch.empty(M * s.Kf, device=dev, dtype=bf), torch.empty(M * s.Kf, device=dev, dtype=bf))
    d["gin"] = (torch.empty(N * s.hout * s.wout * s.C, device=dev, dtype=bf), torch.empty(N * s.hout * s.wout * s.C, device=dev, dtype=bf))
    d["mask"] = rnd(N * s.hout * s.wout * s.C)
    d["sc"] = torch.rand(s.Kf, device=dev) + 0.5
    d["sh"] = torch.randn(s.Kf, device=dev)
    return d


_ws = [torch.empty(1, device=dev)]


def workspace(n):
    if _ws[0].numel() < n:
        _ws[0] = torch.empty(n, device=dev)
    return _ws[0]


def run(d, mode):
    s = d["s"]
    if mode == "fprop":
        L.conv_gemm(d["x"][0], d["x"][1], d["wf"][0], d["wf"][1], s.geom(N), scale=d["sc"], shift=d["sh"], relu=True,
                    out_hi=d["out"][0], out_lo=d["out"][1])
    elif mode == "fprop_res":      # bottleneck conv3 as it runs in the step: BN affine + residual (split planes) + ReLU
        if "res" not in d:
            d["res"] = (rnd(d["M"] * s.Kf), rnd(d["M"] * s.Kf))
        L.conv_gemm(d["x"][0], d["x"][1], d["wf"][0], d["wf"][1], s.geom(N), scale=d["sc"], shift=d["sh"], relu=True,
                    add_hi=d["res"][0], add_lo=d["res"][1], out_hi=d["out"][0], out_lo=d["out"][1])
    elif mode == "fprop_res_unit":  # the same with the BN scale folded into the weights (what the engine runs since round 2): residual via the tensor core
        if "res" not in d:
            d["res"] = (rnd(d["M"] * s.Kf), rnd(d["M"] * s.Kf))
        if "ones" not in d:
            d["ones"] = torch.ones_like(d["sc"])
        L.conv_gemm(d["x"][0], d["x"][1], d["wf"][0], d["wf"][1], s.geom(N), scale=d["ones"], shift=d["sh"], relu=True,
                    add_hi=d["res"][0], add_lo=d["res"][1], out_hi=d["out"][0], out_lo=d["out"][1], unit_scale=True)
    elif mode in ("dgrad_cs", "dgrad_res"):   # data gradient as it runs in the step: ReLU mask + d(beta) column sums (+ residual gradient)
        if "cs" not in d:
            d["cs"] = torch.zeros(s.C, device=dev)
            d["gres"] = (rnd(N * s.hout * s.wout * s.C), rnd(N * s.hout * s.wout * s.C))
        extra = dict(add_hi=d["gres"][0], add_lo=d["gres"][1]) if mode == "dgrad_res" else {}
        L.conv_gemm(d["g"][0], d["g"][1], d["wt"][0], d["wt"][1], s.geom_dgrad(N), mask_hi=d["mask"], colsum=d["cs"],
                    out_hi=d["gin"][0], out_lo=d["gin"][1], **extra)
    elif mode == "dgrad":
        L.conv_gemm(d["g"][0], d["g"][1], d["wt"][0], d["wt"][1], s.geom_dgrad(N), mask_hi=d["mask"],
                    out_hi=d["gin"][0], out_lo=d["gin"][1])
    else:
        L.conv_wgrad(d["x"][0], d["x"][1], d["g"][0], d["g"][1], workspace, (N, s.hin, s.win, s.C, s.Kt, s.R, s.stride, s.dil, s.pad), k_valid=s.K)


def main():
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        s = net["specs"][sys.argv[2]]
        d = make(s)
        for _ in range(4):
            flush.zero_()
            run(d, sys.argv[3])
        torch.cuda.synchronize()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "epilogues":
        # the epilogue variants the real step uses, on the shapes that carry them
        print("%-28s %-10s %9s %8s" % ("shape", "mode", "us", "TF/s"))
        for nm, modes in (("model.layer3.1.conv3", ("fprop", "fprop_res", "fprop_res_unit", "dgrad", "dgrad_cs")),
                          ("model.layer3.1.conv1", ("fprop", "dgrad", "dgrad_cs", "dgrad_res")),
                          ("model.layer3.1.conv2", ("fprop", "dgrad", "dgrad_cs"))):
            s = net["specs"][nm]
            d = make(s)
            fl = 2.0 * d["M"] * s.K * s.C * s.R * s.R
            for mode in modes:
                for _ in range(2): run(d, mode)
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); run(d, mode); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t = sorted(ts)[len(ts) // 2]
                print("%-28s %-10s %9.1f %8.1f" % ("C%d K%d %dx%d" % (s.C, s.K, s.R, s.R), mode, t * 1e3, fl / (t * 1e-3) / 1e12))
            del d
            torch.cuda.empty_cache()
        return
    seen = {}
    for name in net["order"]:
        s = net["specs"][name]
        if s.C == 3: continue
        key = (s.C, s.K, s.R, s.stride, s.dil, s.hin)
        seen.setdefault(key, [s, 0])[1] += 1
    print("%-34s %3s %9s %9s | %8s %7s | %8s %7s | %8s %7s" % ("shape", "cnt", "M", "GFLOP", "fprop us", "TF/s", "dgrad us", "TF/s", "wgrad us", "TF/s"))
    tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    only = os.environ.get("SACB_ONLY")          # e.g. SACB_ONLY="C256 K1024,C1024 K256"
    for key, (s, cnt) in seen.items():
        if only and not any(("C%d K%d " % (s.C, s.K)) == o.strip() + " " for o in only.split(",")):
            continue
        d = make(s)
        fl = 2.0 * d["M"] * s.K * s.C * s.R * s.R
        res = []
        for mode in ("fprop", "dgrad", "wgrad"):
            for _ in range(2): run(d, mode)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); run(d, mode); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            res += [t * 1e3, fl / (t * 1e-3) / 1e12]
            tot[mode] += t * cnt * (2 if mode == "fprop" else 1)
        print("%-34s %3d %9d %9.1f | %8.1f %7.1f | %8.1f %7.1f | %8.1f %7.1f" %
              ("C%d K%d %dx%d s%d d%d @%d" % (s.C, s.K, s.R, s.R, s.stride, s.dil, s.hin), cnt, d["M"], fl / 1e9, *res))
        del d
        torch.cuda.empty_cache()
    print("per-step totals (ms): fprop x2 nets %.2f, dgrad %.2f, wgrad %.2f" % (tot["fprop"], tot["dgrad"], tot["wgrad"]))


if __name__ == "__main__":
    main()
