import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import golden
from oracle import ssq_oracle as O
from shiftedscalequantization_b200 import quant as Q, zoo
WQ = {'n_bits': 2, 'channel_wise': True, 'scale_method': 'max'}
AQ = {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}
g = golden("recon_loop")
torch.manual_seed(1005)
cnn = zoo.resnet18(num_classes=10).cuda().eval()
qnn = Q.QuantModel(cnn, dict(WQ), dict(AQ)).cuda().eval()
qnn.set_first_last_layer_to_8bit()
cali = torch.from_numpy(g["cali"])
qnn.set_quant_state(True, False)
with torch.no_grad():
    l0 = qnn(cali[:8].cuda()).cpu().numpy()
print("logits before any reconstruction: rel L2 vs golden-final", np.linalg.norm(l0 - g["final_logits"]) / np.linalg.norm(g["final_logits"]))
# per-layer check of quantised weights against the oracle
for name, m in qnn.named_modules():
    if isinstance(m, Q.QuantModule):
        q = m.weight_quantizer
        w = m.org_weight.cpu().numpy()
        ref, _ = O.uaq_forward(w, q.delta.detach().cpu().numpy(), q.zero_point.detach().cpu().numpy(), 0, q.n_levels - 1)
        got = q(m.weight).detach().cpu().numpy()
        d_ref, z_ref, _ = zip(*[O.max_init(r, q.n_bits) for r in w.reshape(w.shape[0], -1)])
        ok_d = np.array_equal(np.array(d_ref, np.float32), q.delta.detach().cpu().numpy().ravel())
        print(f"{name:28s} bits={q.n_bits} wq==oracle {np.array_equal(ref, got)} delta==oracle-max {ok_d}")
kw = dict(cali_data=cali, iters=12, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=False, opt_mode='mse', batch_size=16)
torch.manual_seed(77)
block = qnn.model.layer1[0]
Q.block_reconstruction(qnn, block, **kw)
qnn.set_quant_state(True, False)
with torch.no_grad():
    l1 = qnn(cali[:8].cuda()).cpu().numpy()
print("after block recon: rel L2 vs golden-final", np.linalg.norm(l1 - g["final_logits"]) / np.linalg.norm(g["final_logits"]), " vs before", np.linalg.norm(l1 - l0) / np.linalg.norm(l0))
torch.manual_seed(78)
Q.layer_reconstruction(qnn, qnn.model.fc, **kw)
qnn.set_quant_state(True, False)
with torch.no_grad():
    l2 = qnn(cali[:8].cuda()).cpu().numpy()
print("after fc recon: rel L2 vs golden-final", np.linalg.norm(l2 - g["final_logits"]) / np.linalg.norm(g["final_logits"]))
print(l2[0], g["final_logits"][0])
fc = qnn.model.fc
print("fc alpha sign flips vs golden", int((np.sign(fc.weight_quantizer.alpha.detach().cpu().numpy()) != np.sign(g["fc.alpha"])).sum()))
print("fc delta equal", np.array_equal(fc.weight_quantizer.delta.detach().cpu().numpy(), g["fc.delta"]), "weight equal", np.array_equal(fc.org_weight.cpu().numpy(), g["fc.weight"]))
