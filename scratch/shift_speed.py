import sys, time, torch
sys.path.insert(0, '.')
from shiftedscalequantization_b200 import quant as Q, zoo
from shiftedscalequantization_b200.quant.channelQuant import ChannelQuant
from shiftedscalequantization_b200.quant import layer_recon_shiftedScale as LS
from shiftedscalequantization_b200.quant.layer_recon_fused_shiftedScale import block_recon_fused_shiftedScale
import io, contextlib
def run(arch, path, captured, fused, iters=300, n=256):
    torch.manual_seed(1005)
    cnn = zoo.build(arch).cuda().eval()
    qnn = Q.QuantModel(cnn, {'n_bits': 4 if arch=='resnet50' else 2, 'channel_wise': True, 'scale_method': 'max'},
                       {'n_bits': 4, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True}).cuda().eval()
    qnn.set_first_last_layer_to_8bit()
    cali = torch.randn(n, 3, 224, 224)
    qnn.set_quant_state(True, False)
    with torch.no_grad(): qnn(cali[:32].cuda())
    block = qnn
    for p in path.split('.'): block = block[int(p)] if p.isdigit() else getattr(block, p)
    mods = [m for m in block.modules() if isinstance(m, Q.QuantModule)]
    for m in mods:
        m.weight_quantizer = ChannelQuant(1.0, uaq=m.weight_quantizer, weight_tensor=m.org_weight.data, shiftTarget=[0.96875, 1.03125, 1.0], name=m.pathName)
    qnn.set_quant_state(True, False); block.cache_features = 'if'
    with torch.no_grad():
        for i in range(0, n, 32): qnn(cali[i:i+32].cuda())
    block.cache_features = 'none'; qnn.set_quant_state(False, False); block.cache_features = 'of'
    with torch.no_grad():
        for i in range(0, n, 32): qnn(cali[i:i+32].cuda())
    block.cache_features = 'none'; block.set_quant_state(True, False)
    LS.USE_CAPTURED_LOOP = captured
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        if fused: block_recon_fused_shiftedScale(block, iters=iters, lmda=[0.01, 0.01], model=qnn)
        elif isinstance(block, Q.QuantModule): LS.layer_recon_shiftedScale(block, iters=iters, lmda=0.01, model=qnn)
        else: LS.block_recon_shiftedScale(block, iters=iters, lmda=0.01, model=qnn)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return iters / dt
for arch, path in (('resnet18', 'model.layer2.0'), ('resnet18', 'model.layer4.1'), ('resnet50', 'model.layer1.0'), ('resnet50', 'model.layer3.2.conv2')):
    for fused in ((False,) if path.endswith('conv2') else (False, True)):
        r = {c: run(arch, path, c, fused) for c in (False, True)}
        # second, longer run amortises capture
        r2 = run(arch, path, True, fused, iters=1500)
        print(f"{arch} {path} fused={fused}: eager {r[False]:.0f} it/s, captured(300 it incl. capture) {r[True]:.0f}, captured(1500 it) {r2:.0f}", flush=True)
