"""One profiled SAC target step (configs[1]) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`; see profiles/README.md for the exact commands."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da_sac_b200 import synth  # noqa: E402
from da_sac_b200.models import get_model  # noqa: E402
from da_sac_b200.trainer import TargetStepper  # noqa: E402

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 8
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.ModelCfg()
net = get_model(cfg, 0, num_classes=19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255, reduction="none"))
net.backbone.load_state_dict(synth.make_backbone_params(seed=123))
net.cuda().train()
st = TargetStepper(net, cfg, 3, torch.device("cuda"))
batch = st.h2d(synth.make_target_batch(groups, 3, (512, 512), seed=0))
for _ in range(warm):
    st.step(tuple(t.clone() for t in batch))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
st.step(tuple(t.clone() for t in batch))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
