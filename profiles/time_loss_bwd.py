import sys, torch, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da_sac_b200 import lib as L, synth
BT, Cn, h, w, H, W = 24, 19, 65, 65, 512, 512
dev = torch.device("cuda")
torch.manual_seed(0)
logits = torch.randn(BT, Cn, h, w, device=dev) * 3
y = torch.full((BT, H, W), 255, dtype=torch.int64, device=dev)
labels = torch.randint(0, 19, (BT, H, W), device=dev).to(torch.uint8)
labels[torch.rand(BT, H, W, device=dev) < 0.4] = 255
conf = torch.rand(H, W, device=dev); rc = torch.rand(19, device=dev) * 0.1
losses = torch.zeros(2, device=dev); scratch = torch.zeros(2, dtype=torch.float64, device=dev)
gpx = torch.empty(BT * Cn * H * W, device=dev); grows = torch.empty(BT * Cn * H * w, device=dev)
outs = []
for ws in (False, True):
    dl = torch.zeros_like(logits)
    d = L.Loss(C.sizeof(L.Loss), BT, Cn, h, w, H, W, L.ptr(logits), L.ptr(y), L.ptr(labels), L.ptr(conf), L.ptr(rc), 3.0, L.ptr(losses),
               L.ptr(scratch), 5.0, L.ptr(dl), L.ptr(gpx) if ws else None, L.ptr(grows) if ws else None)
    for _ in range(3): L.check(L.lib().sacb_student_loss_bwd(C.byref(d), L.stream()), "bwd")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.check(L.lib().sacb_student_loss_bwd(C.byref(d), L.stream()), "bwd")
    e1.record(); torch.cuda.synchronize()
    print("two-stage" if ws else "gather   ", "%.3f ms" % (e0.elapsed_time(e1) / 10))
    outs.append(dl.clone())
print("rel diff two-stage vs gather: %.3e" % ((outs[0] - outs[1]).norm() / outs[0].norm()).item())
