"""e2e (host-resident cache) step time against the number of CTAs the host-pull kernels use.  python scratch/e2e_ctas_probe.py 8 16 24"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                                    # noqa: E402
from shiftedscalequantization_b200 import engine as E               # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
Q, qnn, cali = B.build_model(dev, 1024)
qnn.set_quant_state(True, False)
with torch.no_grad():
    qnn(cali[:64].to(dev))
engines, feats = B.make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=False)
out = {"device_resident_ms": B.timed_steps(engines, 30, 5, dev, 1)}
B.release(engines)
for ctas in [int(a) for a in sys.argv[1:]] or [8, 16, 24]:
    E.PULL_CTAS = ctas
    eng, _ = B.make_engines(Q, qnn, cali, dev, act_quant=False, multi_gpu=False, host_resident=True, feats=feats)
    out[f"pull_ctas_{ctas}_ms"] = B.timed_steps(eng, 20, 5, dev, 1, read_loss=True)
    B.release(eng)
    del eng
    torch.cuda.empty_cache()
B.emit(json.dumps(out))
