"""Calibration feature capture behind the reference's quant/data_utils.py names (save_inp_oup_data :8-37, save_grad_data
:40-71, GetLayerInpOut :102-139, GetLayerGrad :155-192, quantize_model_till :195-206).

What the reference computes for a unit U: the FP output of U and — with asym=True — the input U receives when everything
in front of it is quantised. It gets both by running the network from the image up to U twice per mini-batch, for every
unit (64 prefix forwards per unit for 1 024 images), and moves every batch to the host and back.

Here the two activations are CARRIED from unit to unit. A calibration run visits the units in
execution order, so when U is requested the previous request's results are still at hand:
    FP input of U        = glue(FP output of U_prev)                      (the FP output was U_prev's `cached_outs`)
    quantised input of U = glue(U_prev in its current quantised state applied to ITS quantised input = `cached_inps`)
where `glue` is whatever the model does between the two units (max-pool, average-pool + flatten, ...). One forward of
U_prev, one of U and the glue per mini-batch replace two prefix forwards. The glue is discovered once per model by
`ModelTrace` (a hook-recorded forward on one sample: which modules run between two units, each fed the previous one's
output; views and a few pure tensor functions are recognised by checking them against the traced tensors).

Safety: the first mini-batch of every carried capture is ALSO computed the reference's way (a truncated forward from
the image); unless both agree bit for bit the whole unit is captured the reference's way. A model whose glue cannot be
replayed (or any out-of-order request) therefore just runs the slow path — never a different result.
Captured tensors stay on the device (keep_gpu=True) or go to pinned host memory (keep_gpu=False, the reference's
host-resident mode). With a process group each rank captures its own shard of the calibration set (..dist).
"""
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F

from .quant_block import BaseQuantBlock
from .quant_layer import QuantModule
from .quant_model import QuantModel

Unit = Union[QuantModule, BaseQuantBlock]


# ------------------------------------------------------------------------------------------- truncated forward
class StopForwardException(Exception):
    """ends a forward pass at the unit of interest (same name as upstream's, data_utils.py:74-78)"""


class _Tap:
    """context manager: taps `unit` (first positional input, output) and optionally ends the forward there"""

    def __init__(self, unit: torch.nn.Module, stop: bool = True):
        self.unit, self.stop = unit, stop
        self.inp = self.out = None

    def _hook(self, _module, args, output):
        self.inp, self.out = args[0], output
        if self.stop:
            raise StopForwardException

    def __enter__(self):
        self._handle = self.unit.register_forward_hook(self._hook)
        return self

    def __exit__(self, exc_type, exc, tb):
        self._handle.remove()
        return exc_type is StopForwardException          # the truncated forward is the expected way out

    def run(self, model, x):
        with self:
            model(x)
        return self.inp, self.out


class DataSaverHook:
    """forward hook that stores a module's input / output (upstream data_utils.py:81-99); kept for callers that
    register it themselves — the capture code in this file uses _Tap"""

    def __init__(self, store_input=False, store_output=False, stop_forward=False):
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = self.output_store = None

    def __call__(self, module, input_batch, output_batch):
        self.input_store = input_batch if self.store_input else self.input_store
        self.output_store = output_batch if self.store_output else self.output_store
        if self.stop_forward:
            raise StopForwardException


class GetLayerInpOut:
    """(input, FP output) of `layer` for one mini-batch, from the image (upstream data_utils.py:102-139): the FP output from
    a forward with every quantiser off; with asym the input from a second forward with the prefix quantised."""

    def __init__(self, model: QuantModel, layer: Unit, device: torch.device, asym: bool = False, act_quant: bool = False):
        self.model, self.layer, self.device, self.asym, self.act_quant = model, layer, device, asym, act_quant

    @torch.no_grad()
    def __call__(self, model_input):
        model, layer = self.model, self.layer
        model.eval()
        x = model_input.to(self.device)
        model.set_quant_state(False, False)
        inp, out = _Tap(layer).run(model, x)
        if self.asym:
            model.set_quant_state(weight_quant=True, act_quant=self.act_quant)
            inp, _ = _Tap(layer).run(model, x)
        _leave_unit_ready(model, layer, self.act_quant)
        return inp.detach(), out.detach()


def _leave_unit_ready(model, layer, act_quant):
    """the state upstream leaves behind after a capture (data_utils.py:135-137, :187-189)"""
    model.set_quant_state(False, False)
    layer.set_quant_state(True, act_quant)
    model.train()


# ------------------------------------------------------------------------------------------- the model's call chain
class ModelTrace:
    """Which modules run between two units, and how the tensor gets from one to the next.

    One hook-recorded forward on a single sample: every module call is logged with its nesting depth, its first input and
    its output (the tensors are kept alive for the duration, so identity comparisons are meaningful). `glue(a, b)` returns
    the list of callables that turn unit a's output into unit b's input, or None when that cannot be established."""

    #: pure tensor functions models use between modules; a candidate is accepted only if it reproduces the traced tensor
    FUNCTIONAL_GLUE: List[Tuple[str, Callable]] = [
        ("flatten(1)", lambda t: torch.flatten(t, 1)),
        ("mean([2,3])", lambda t: t.mean([2, 3])),
        ("relu", torch.relu),
    ]

    def __init__(self, model: torch.nn.Module, sample: torch.Tensor):
        self.calls = []                       # dicts: module, depth, inp, out, done (index in completion order)
        stack, handles = [], []

        def pre(mod, args):
            rec = {"module": mod, "depth": len(stack), "inp": args[0] if args and torch.is_tensor(args[0]) else None, "out": None}
            stack.append(rec)
            self.calls.append(rec)

        def post(mod, args, output):
            rec = stack.pop()
            rec["out"] = output if torch.is_tensor(output) else None

        for m in model.modules():
            handles.append(m.register_forward_pre_hook(pre))
            handles.append(m.register_forward_hook(post))
        was_training = model.training
        try:
            with torch.no_grad():
                model.eval()
                model(sample)
        finally:
            for h in handles:
                h.remove()
            model.train(was_training)
        self.index = {}
        for i, rec in enumerate(self.calls):
            self.index.setdefault(id(rec["module"]), i)          # first call of every module
        self._glue_cache = {}

    def release(self):
        """drop the traced tensors (the glue programs found so far stay)"""
        for rec in self.calls:
            rec["inp"] = rec["out"] = None

    def _subtree_end(self, i: int) -> int:
        d = self.calls[i]["depth"]
        j = i + 1
        while j < len(self.calls) and self.calls[j]["depth"] > d:
            j += 1
        return j

    def _bridge(self, cur: torch.Tensor, want: torch.Tensor):
        """a pure function f with f(cur) == want, or None"""
        if want is cur:
            return []
        if want.data_ptr() == cur.data_ptr() and want.numel() == cur.numel():         # a view (flatten, reshape, squeeze)
            tail = tuple(want.shape[1:])
            return [lambda t, tail=tail: t.reshape((t.shape[0],) + tail)]
        for _name, fn in self.FUNCTIONAL_GLUE:
            try:
                got = fn(cur)
            except Exception:
                continue
            if got.shape == want.shape and torch.equal(got, want):
                return [fn]
        return None

    def glue(self, a: torch.nn.Module, b: torch.nn.Module):
        key = (id(a), id(b))
        if key not in self._glue_cache:
            self._glue_cache[key] = self._find_glue(a, b)
        return self._glue_cache[key]

    def _find_glue(self, a, b):
        ia, ib = self.index.get(id(a)), self.index.get(id(b))
        if ia is None or ib is None or self.calls[ia]["out"] is None or self.calls[ib]["inp"] is None:
            return None
        start = self._subtree_end(ia)
        if ib < start:
            return None                                   # b does not run after a
        cur, steps, i = self.calls[ia]["out"], [], start
        while i < ib:
            rec = self.calls[i]
            if self._subtree_end(i) > ib:                 # an ancestor of b (a container): it hands its input down
                i += 1
                continue
            if rec["inp"] is None or rec["out"] is None:
                return None
            pre = self._bridge(cur, rec["inp"])
            if pre is None:
                return None                               # a side branch or something this trace cannot express
            steps += pre + [rec["module"]]
            cur = rec["out"]
            i = self._subtree_end(i)
        last = self._bridge(cur, self.calls[ib]["inp"])
        return None if last is None else steps + last


# ------------------------------------------------------------------------------------------- carried capture
class _Frontier:
    """what the previous capture left behind: the unit, its cached inputs / FP outputs and the flags they were captured under"""

    def __init__(self, unit, inps, outs, key):
        self.unit, self.inps, self.outs, self.key = unit, inps, outs, key


class CaptureStats:
    def __init__(self):
        self.carried = self.from_image = 0
        self.last_mode = None

    def note(self, mode):
        self.last_mode = mode
        if mode == 'carried':
            self.carried += 1
        else:
            self.from_image += 1


def _plan(model) -> dict:
    plan = model.__dict__.get('_ssq_capture')
    if plan is None:
        plan = {"trace": None, "frontier": None, "stats": CaptureStats(), "enabled": True}
        model.__dict__['_ssq_capture'] = plan             # plain attribute: not a submodule, not in the state_dict
    return plan


def capture_stats(model) -> CaptureStats:
    return _plan(model)["stats"]


def reset_capture(model, enabled: Optional[bool] = None):
    """forget the carried activations (call after changing anything in front of the next unit by hand); `enabled=False`
    makes every capture run from the image, as upstream"""
    plan = _plan(model)
    plan["frontier"] = None
    if enabled is not None:
        plan["enabled"] = bool(enabled)


def _cali_key(cali_data, batch_size, asym, act_quant):
    return (cali_data.data_ptr(), tuple(cali_data.shape), cali_data._version, int(batch_size), bool(asym), bool(act_quant))


def _run_glue(steps, x):
    for f in steps:
        x = f(x)
    return x


class _Sink:
    """the cached tensor of one capture, filled mini-batch by mini-batch: one [N, ...] allocation on the device
    (keep_gpu=True) or in pinned host memory (keep_gpu=False, upstream's host-resident mode) — never both"""

    def __init__(self, n_rows: int, keep_gpu: bool):
        self.n_rows, self.keep_gpu, self.buf, self.fill = n_rows, keep_gpu, None, 0

    def add(self, batch: torch.Tensor):
        if self.buf is None:
            shape = (self.n_rows,) + tuple(batch.shape[1:])
            self.buf = torch.empty(shape, dtype=batch.dtype, device=batch.device) if self.keep_gpu else \
                torch.empty(shape, dtype=batch.dtype, pin_memory=True)
        n = batch.shape[0]
        self.buf[self.fill:self.fill + n].copy_(batch, non_blocking=True)
        self.fill += n

    def result(self):
        if self.buf is None:
            return torch.empty(0)
        if not self.keep_gpu:
            torch.cuda.current_stream().synchronize()      # the asynchronous device-to-host copies have landed
        return self.buf[:self.fill]


@torch.no_grad()
def _carried_capture(model, layer, cali_data, asym, act_quant, batch_size, keep_gpu, device, frontier, steps):
    """inputs / FP outputs of `layer` from the previous unit's cached tensors; None if the first mini-batch does not
    reproduce the from-the-image capture bit for bit"""
    prev = frontier.unit
    n_batches = int(cali_data.size(0) / batch_size)
    inps, outs = _Sink(n_batches * batch_size, keep_gpu), _Sink(n_batches * batch_size, keep_gpu)
    model.eval()
    for i in range(n_batches):
        sl = slice(i * batch_size, (i + 1) * batch_size)
        model.set_quant_state(False, False)
        fp_in = _run_glue(steps, frontier.outs[sl].to(device, non_blocking=True))
        out = layer(fp_in)
        if asym:
            model.set_quant_state(weight_quant=True, act_quant=act_quant)
            inp = _run_glue(steps, prev(frontier.inps[sl].to(device, non_blocking=True)))
        else:
            inp = fp_in
        if i == 0:
            ref_inp, ref_out = GetLayerInpOut(model, layer, device=device, asym=asym, act_quant=act_quant)(cali_data[sl])
            model.eval()
            if not (torch.equal(ref_inp, inp) and torch.equal(ref_out, out)):
                return None
        inps.add(inp)
        outs.add(out)
    _leave_unit_ready(model, layer, act_quant)
    return inps.result(), outs.result()


def save_inp_oup_data(model: QuantModel, layer: Unit, cali_data: torch.Tensor, asym: bool = False, act_quant: bool = False,
                      batch_size: int = 32, keep_gpu: bool = True):
    """(inputs, FP outputs) of `layer` over the calibration set (signature and results of upstream data_utils.py:8-37);
    asym=True takes the inputs from the already-quantised prefix of the network."""
    device = next(model.parameters()).device
    plan = _plan(model)
    key = _cali_key(cali_data, batch_size, asym, act_quant)
    frontier = plan["frontier"]
    result = None
    if plan["enabled"] and frontier is not None and frontier.key == key and frontier.unit is not layer:
        if plan["trace"] is None:
            model.set_quant_state(False, False)          # the traced forward must not initialise any quantiser
            plan["trace"] = ModelTrace(model, cali_data[:1].to(device))
        steps = plan["trace"].glue(frontier.unit, layer)
        if steps is not None:
            result = _carried_capture(model, layer, cali_data, asym, act_quant, batch_size, keep_gpu, device, frontier, steps)
    if result is None:
        get_inp_out = GetLayerInpOut(model, layer, device=device, asym=asym, act_quant=act_quant)
        n_batches = int(cali_data.size(0) / batch_size)
        inps, outs = _Sink(n_batches * batch_size, keep_gpu), _Sink(n_batches * batch_size, keep_gpu)
        for i in range(n_batches):
            cur_inp, cur_out = get_inp_out(cali_data[i * batch_size:(i + 1) * batch_size])
            inps.add(cur_inp)
            outs.add(cur_out)
        result = (inps.result(), outs.result())
        plan["stats"].note('from_image')
    else:
        plan["stats"].note('carried')
    plan["frontier"] = _Frontier(layer, result[0], result[1], key) if plan["enabled"] else None
    return result


# ------------------------------------------------------------------------------------------- Fisher gradients
class GradSaverHook:
    """backward hook that keeps the gradient w.r.t. a module's output (upstream data_utils.py:142-152)"""

    def __init__(self, store_grad=True):
        self.store_grad, self.stop_backward, self.grad_out = store_grad, False, None

    def __call__(self, module, grad_input, grad_output):
        self.grad_out = grad_output[0] if self.store_grad else self.grad_out
        if self.stop_backward:
            raise StopForwardException


class GetLayerGrad:
    """d KL(FP logits || logits with everything up to `layer` quantised) / d(layer output) for one mini-batch
    (upstream data_utils.py:155-192)"""

    def __init__(self, model: QuantModel, layer: Unit, device: torch.device, act_quant: bool = False):
        self.model, self.layer, self.device, self.act_quant = model, layer, device, act_quant
        self.data_saver = GradSaverHook(True)

    def __call__(self, model_input):
        model = self.model
        model.eval()
        inputs = model_input.to(self.device)
        handle = self.layer.register_full_backward_hook(self.data_saver)
        try:
            with torch.enable_grad():
                model.zero_grad()
                model.set_quant_state(False, False)
                target = F.softmax(model(inputs), dim=1)
                quantize_model_till(model, self.layer, self.act_quant)
                log_q = F.log_softmax(model(inputs), dim=1)
                F.kl_div(log_q, target, reduction='batchmean').backward()
        except StopForwardException:
            pass
        finally:
            handle.remove()
        _leave_unit_ready(model, self.layer, self.act_quant)
        return self.data_saver.grad_out.data


def save_grad_data(model: QuantModel, layer: Unit, cali_data: torch.Tensor, damping: float = 1., act_quant: bool = False,
                   batch_size: int = 32, keep_gpu: bool = True):
    """|dKL/d(layer output)| + 1 over the calibration set (Fisher weights, upstream data_utils.py:40-71)"""
    device = next(model.parameters()).device
    get_grad = GetLayerGrad(model, layer, device, act_quant=act_quant)
    n_batches = int(cali_data.size(0) / batch_size)
    sink = _Sink(n_batches * batch_size, keep_gpu)
    for i in range(n_batches):
        sink.add(get_grad(cali_data[i * batch_size:(i + 1) * batch_size]).abs() + 1.0)
    return sink.result()


def quantize_model_till(model: QuantModule, layer: Unit, act_quant: bool = False):
    """quantise every unit up to and including `layer` (module order == execution order for all zoo models)"""
    model.set_quant_state(False, False)
    for module in model.modules():
        if isinstance(module, (QuantModule, BaseQuantBlock)):
            module.set_quant_state(True, act_quant)
        if module is layer:
            return
