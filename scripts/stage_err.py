import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from oracle import tag_oracle as O
from helpers import build_model, rel_err
B, L = 64, 320000
batch = O.synth_batch(B, L, seed=21, ragged=True)
d = {"specaug": False}
for k, v in batch.items():
    d[k] = v.cuda() if isinstance(v, torch.Tensor) else v
for sharpen in (300.0, 100.0):
    sd = O.synth_state_dict(seed=3, sharpen=sharpen, perturb_bn=True)
    st = {}
    outs = {}
    for prec in ("fp32", "bf16"):
        m = build_model(sd, prec).eval()
        m.audio_encoder._stages = {}
        with torch.no_grad():
            outs[prec] = m(dict(d))["frame_sim"].float().cpu()
        st[prec] = {k: v.float().cpu() for k, v in m.audio_encoder._stages.items()}
    print("sharpen", sharpen, "frame_sim max diff", (outs["bf16"]-outs["fp32"]).abs().max().item(),
          "p99.9", (outs["bf16"]-outs["fp32"]).abs().flatten().kthvalue(int(0.999*outs["fp32"].numel())).values.item())
    for k in st["fp32"]:
        print("  ", k, "rel err", round(rel_err(st["bf16"][k], st["fp32"][k]), 5))
