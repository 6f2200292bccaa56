"""Feature capture (quant/data_utils.py): the call-chain trace and the carried FP branch, on the CPU.
With every quantiser off a QuantModel is plain PyTorch, so the FP half of the capture logic runs without a GPU; the
quantised half is covered by tests/test_round2_gpu.py."""
import pytest
import torch

from shiftedscalequantization_b200 import quant as Q, zoo
from shiftedscalequantization_b200.quant import data_utils as DU


def _units(qnn):
    out = []

    def walk(m):
        for _n, c in m.named_children():
            if isinstance(c, (Q.QuantModule, Q.BaseQuantBlock)):
                out.append(c)
            else:
                walk(c)
    walk(qnn)
    return out


@pytest.mark.parametrize("arch,kw,res", [("resnet18", {"num_classes": 10}, 32), ("resnet50", {"num_classes": 10}, 32),
                                         ("mobilenetv2", {}, 32), ("regnetx_600m", {}, 32)])
def test_trace_finds_the_glue_between_consecutive_units(arch, kw, res):
    torch.manual_seed(0)
    qnn = Q.QuantModel(zoo.build(arch, **kw).eval(), {'n_bits': 4}, {'n_bits': 4}).eval()
    qnn.set_quant_state(False, False)
    x = torch.randn(2, 3, res, res)
    units = _units(qnn)
    trace = DU.ModelTrace(qnn, x[:1])
    taps = {}
    hs = [u.register_forward_hook(lambda m, a, o, u=u: taps.__setitem__(id(u), (a[0], o))) for u in units]
    with torch.no_grad():
        qnn(x)
    for h in hs:
        h.remove()
    found = 0
    for a, b in zip(units[:-1], units[1:]):
        steps = trace.glue(a, b)
        assert steps is not None, f"{arch}: no glue between {getattr(a, 'pathName', a)} and {getattr(b, 'pathName', b)}"
        with torch.no_grad():
            got = DU._run_glue(steps, taps[id(a)][1])
        assert torch.equal(got, taps[id(b)][0])
        found += 1
    assert found == len(units) - 1
    assert trace.glue(units[3], units[1]) is None                # not in execution order


def test_carried_fp_capture_equals_capture_from_the_image():
    torch.manual_seed(1)
    qnn = Q.QuantModel(zoo.resnet18(num_classes=10).eval(), {'n_bits': 4}, {'n_bits': 4}).eval()
    qnn.set_quant_state(False, False)
    cali = torch.randn(16, 3, 32, 32)
    units = _units(qnn)
    carried = []
    for u in units:
        # asym=False: inputs and outputs both come from the FP branch, so no quantiser kernel is needed
        real_set = u.set_quant_state
        u.set_quant_state = lambda *a, **k: None            # keep the unit FP on the CPU (the capture re-enables it at the end)
        carried.append(DU.save_inp_oup_data(qnn, u, cali, asym=False, act_quant=False, batch_size=8))
        u.set_quant_state = real_set
        qnn.set_quant_state(False, False)
    st = DU.capture_stats(qnn)
    assert st.from_image == 1 and st.carried == len(units) - 1, (st.from_image, st.carried)
    DU.reset_capture(qnn, enabled=False)
    for u, (ci, co) in zip(units, carried):
        real_set = u.set_quant_state
        u.set_quant_state = lambda *a, **k: None
        ri, ro = DU.save_inp_oup_data(qnn, u, cali, asym=False, act_quant=False, batch_size=8)
        u.set_quant_state = real_set
        qnn.set_quant_state(False, False)
        assert torch.equal(ci, ri) and torch.equal(co, ro)
    assert DU.capture_stats(qnn).carried == len(units) - 1     # nothing was carried with the switch off


def test_inplace_glue_is_caught_by_the_first_batch_check():
    """a model that changes the tensor in place between two units defeats the identity-based trace; the first-batch
    comparison must send the capture down the from-the-image path"""
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Conv2d(3, 4, 3, padding=1)
            self.b = torch.nn.Conv2d(4, 4, 3, padding=1)

        def forward(self, x):
            y = self.a(x)
            y += 1.0                                          # invisible to the trace
            return self.b(y)
    torch.manual_seed(2)
    qnn = Q.QuantModel(Net().eval(), {'n_bits': 4}, {'n_bits': 4}).eval()
    qnn.set_quant_state(False, False)
    cali = torch.randn(8, 3, 8, 8)
    ua, ub = qnn.model.a, qnn.model.b
    for u in (ua, ub):
        u.set_quant_state = lambda *a, **k: None
    DU.save_inp_oup_data(qnn, ua, cali, asym=False, batch_size=4)
    inp_b, out_b = DU.save_inp_oup_data(qnn, ub, cali, asym=False, batch_size=4)
    st = DU.capture_stats(qnn)
    assert st.carried == 0 and st.from_image == 2
    with torch.no_grad():
        assert torch.equal(inp_b, ua(cali) + 1.0)
